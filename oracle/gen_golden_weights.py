"""TEST INFRASTRUCTURE ONLY — golden vectors of the reference's sigma builders (models/weights.py:10-76, 148-240).

Run in the build container:  python oracle/gen_golden_weights.py  ->  tests/golden/weights.npz
The UNMODIFIED reference file is loaded through the config stubs of oracle/ref_shim.py (USE_CUSTOM_WEIGHTS is an
import-time constant of the reference: the file is loaded twice, once per value)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
out = {"t": T}
for custom in (True, False):
    ref_shim.install_stubs()
    sys.modules["config.constants"].USE_CUSTOM_WEIGHTS = custom
    w = ref_shim.load("models/weights.py", modname=f"_pkref_weights_{int(custom)}")
    for ns in (1, 3, 5):
        rng = np.random.default_rng(100 + ns)
        pr = rng.uniform(0.2, 2.0, (1, 14))
        pd_ = rng.uniform(0.0, 2.0, (ns, 14))
        pd_[0, 3] = 0.0                                    # an exact zero in the data
        r = rng.uniform(0.5, 1.5, 9)
        target = np.concatenate([r, pr.ravel(), pd_.ravel()])
        ms = rng.uniform(0.05, 0.5, 14 * (ns + 1))
        P = 4 + 2 * ns
        early = w.early_emphasis(pr, pd_, T, ns)
        opts = w.get_weight_options(target, T, ns, True, P, early, ms)
        tag = f"ns{ns}_c{int(custom)}"
        out[f"{tag}_pr"], out[f"{tag}_p"], out[f"{tag}_target"], out[f"{tag}_ms"], out[f"{tag}_early"] = pr, pd_, target, ms, early
        out[f"{tag}_keys"] = np.array(list(opts.keys()))
        for k, v in opts.items():
            out[f"{tag}_opt_{k}"] = np.asarray(v)
        out[f"{tag}_full_noreg"] = w.full_weight(ms, False, P)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "weights.npz"), **out)
print("wrote tests/golden/weights.npz with", len(out), "arrays")
