"""Dev script (GPU box): error of the local kernels against the tight reference goldens vs tolerance."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
eng = pk.get_engine(0)
method = sys.argv[1] if len(sys.argv) > 1 else None      # None = library default (ROS6L on the small models, ROS5L dense)
tols = ((None, None), (1e-7, 1e-10), (1e-6, 1e-9), (2e-6, 2e-9), (5e-6, 5e-9), (1e-5, 2e-9), (2e-5, 2e-9), (4e-5, 2e-9))
for rtol, atol in tols:
    worst_t, worst_s, steps = 0.0, 0.0, []
    per = []
    for f in sorted(glob.glob("tests/golden/local_*.npz")):
        g = np.load(f)
        base = os.path.basename(f)[6:-4]
        model, ns = base.split("_ns"); ns = int(ns)
        r = eng.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("sol",), rtol=rtol, atol=atol,
                                  method=(method if model != "randmod" or method != "ros6l" else None))
        tight, stock = g["sol_tight"], g["sol"]
        e_t = (np.abs(r["sol"] - tight) / (1e-6 * np.abs(tight) + 1e-9)).max()
        e_s = (np.abs(r["sol"] - stock) / (1e-6 * np.abs(stock) + 1e-7)).max()
        worst_t, worst_s = max(worst_t, e_t), max(worst_s, e_s)
        steps.append(r["nsteps"].mean())
        per.append(f"{base}:{e_t:.3g}")
    print(f"method {method} rtol {rtol} atol {atol}: worst vs tight {worst_t:.3g} of bound | vs stock {worst_s:.3g} | mean steps {np.mean(steps):.0f} | " + " ".join(per), flush=True)
