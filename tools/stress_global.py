"""Dev script (GPU box): global kernel over the reference's whole parameter box (config.toml:364-398): status codes,
agreement with a tight-tolerance run, step statistics."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.global_model import simulate_batch, synthetic_system

eng = pk.get_engine(0)
BOX = {"c_k": (1e-3, 5.0), "A_i": (1e-3, 5.0), "B_i": (1e-3, 1.0), "C_i": (1e-3, 2.0), "D_i": (0.1, 0.5), "Dp_i": (0.05, 5.0),
       "E_i": (1e-4, 10.0), "tf_scale": (2.0, 10.0)}
t = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for model, N, K, B in ((0, 120, 40, 2048), (1, 60, 20, 1024), (4, 60, 20, 1024)):
    s = synthetic_system(seed=5, N=N, K=K, max_sites=4, model=model)
    sl = s.param_slices()
    rng = np.random.default_rng(7)
    P = np.empty((B, s.n_params))
    for k, (lo, hi) in BOX.items():
        n = sl[k].stop - sl[k].start
        P[:, sl[k]] = np.exp(rng.uniform(np.log(lo), np.log(hi), (B, n)))        # log-uniform over the box
    b = simulate_batch(s, P[:256], t, ("Y",), rtol=1e-9, atol=1e-13, engine=eng)
    for rtol, atol in ((None, None), (2e-6, 2e-10), (2e-6, 2e-11), (1e-6, 1e-9), (1e-6, 1e-10), (5e-7, 5e-10)):
        a = simulate_batch(s, P, t, ("Y",), rtol=rtol, atol=atol, engine=eng)
        ms = eng.last_launch_info()[1]
        ok = a["status"] == 0
        both = ok[:256] & (b["status"] == 0)
        ratio = np.abs(a["Y"][:256][both] - b["Y"][both]) / (1e-6 * np.abs(b["Y"][both]) + 1e-9)
        worst = np.unravel_index(np.argmax(ratio), ratio.shape)
        st = a["nsteps"] + a["nrej"]
        print(f"model {model} N={N} rtol {rtol} atol {atol}: status {np.bincount(a['status'], minlength=4)} steps mean {st.mean():.0f} max {st.max()} "
              f"| vs tight: max {ratio.max():.3g} (value there {b['Y'][both][worst]:.2e}) p99 {np.percentile(ratio.max(axis=(1, 2)), 99):.3g} median {np.median(ratio.max(axis=(1, 2))):.3g} | "
              f"{B / ms * 1e3:.0f} solves/s", flush=True)
