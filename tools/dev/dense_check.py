"""Dev (GPU box): rand-6 dense kernel — goldens + rate."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for ns in (6,):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"local_randmod_ns{ns}.npz"))
    r = eng.solve_local_batch("randmod", g["params"], g["y0"], ns, g["t"], want=("sol", "flat"))
    e = (np.abs(r["sol"] - g["sol_tight"]) / (1e-6 * np.abs(g["sol_tight"]) + 1e-9)).max()
    print("golden ns", ns, "status", r["status"], "err/bound", e, "steps", r["nsteps"].mean(), flush=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n, P, L = pk.local_dims("randmod", 6, 14)
p = torch.from_numpy(np.random.default_rng(3).uniform(0.05, 3.0, (B, P))).cuda()
y0 = torch.tensor(initial_condition(6, "randmod")).cuda()
tt = torch.from_numpy(T).cuda()
for rep in range(3):
    r = eng.solve_local_batch("randmod", p, y0, 6, tt, want=("flat",))
    ms = eng.last_launch_info()[1]
    print(f"rand-6 B={B}: {ms:.2f} ms -> {B/ms*1e3:.4g} solves/s, steps {r['nsteps'].double().mean().item():.1f}, failed {int((r['status']!=0).sum())}", flush=True)
