"""Dev script (GPU box): parity of the CUDA path against the committed goldens + quick timings."""
import glob, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition

METHODS = (("ros5l", None, None), ("rodas4", 1e-8, 1e-11))
eng = pk.get_engine(0)
print("device", eng.device_name, eng.sm_count, "SMs", flush=True)
print("fp64 peak TF", eng.measure_fp64_peak(), flush=True)
for f in sorted(glob.glob("tests/golden/local_*.npz")):
    g = np.load(f)
    base = os.path.basename(f)[6:-4]
    model, ns = base.split("_ns"); ns = int(ns)
    for method, rtol, atol in METHODS:
        t0 = time.time()
        r = eng.solve_local_batch(model, g["params"], g["y0"], ns, g["t"], want=("sol", "flat", "Y", "score", "ssr"),
                                  target=g["target"], method=method, rtol=rtol, atol=atol)
        dt = time.time() - t0
        tight, stock = g["sol_tight"], g["sol"]
        e_t = np.abs(r["sol"] - tight) / (1e-6 * np.abs(tight) + 1e-9)
        e_s = np.abs(r["sol"] - stock) / (1e-6 * np.abs(stock) + 1e-7)
        e_st = np.abs(stock - tight) / (1e-6 * np.abs(tight) + 1e-9)
        print(f"{base:12s} {method:6s} status {np.bincount(r['status'], minlength=4)} steps mean {r['nsteps'].mean():.0f} max {r['nsteps'].max()} rej {r['nrej'].mean():.1f} "
              f"| vs tight: max {e_t.max():.3g} | vs stock: max {e_s.max():.3g} | stock vs tight: {e_st.max():.3g} "
              f"| score {np.abs(r['score']-g['score']).max():.2e} [{dt*1e3:.1f} ms]", flush=True)

T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for model, ns, B in (("distmod", 3, 1 << 20), ("succmod", 5, 1 << 20), ("distmod", 4, 256000), ("randmod", 6, 4096),
                     ("randmod", 4, 65536), ("distmod", 8, 65536)):
    n, P, L = pk.local_dims(model, ns, 14)
    rng = np.random.default_rng(2)
    params = torch.from_numpy(rng.uniform(0.05, 3.0, (B, P))).cuda()
    y0 = torch.tensor(initial_condition(ns, model)).cuda()
    tt = torch.from_numpy(T).cuda()
    target = torch.rand(L, dtype=torch.float64).cuda()
    for method, rtol, atol in METHODS:
        for rep in range(3):
            r = eng.solve_local_batch(model, params, y0, ns, tt, want=("score",), target=target, method=method,
                                      rtol=rtol, atol=atol)
            nl, ms = eng.last_launch_info()
        st = r["nsteps"].double()
        print(f"{model}-{ns} {method:6s} B={B}: kernel {ms:.2f} ms -> {B/ms*1e3:.3g} solves/s; steps mean {st.mean():.1f} "
              f"max {st.max():.0f} rej {r['nrej'].double().mean():.2f} bad {(r['status']!=0).sum().item()}", flush=True)
