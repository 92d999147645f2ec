"""Dev script (GPU box): run one local config a few times (for ncu captures)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition

model, ns, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
want = tuple(sys.argv[5].split(",")) if len(sys.argv) > 5 else ("score",)
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
n, P, L = pk.local_dims(model, ns, 14)
rng = np.random.default_rng(2)
params = torch.from_numpy(rng.uniform(0.05, 3.0, (B, P))).cuda()
y0 = torch.tensor(initial_condition(ns, model)).cuda()
tt = torch.from_numpy(T).cuda()
target = torch.rand(L, dtype=torch.float64).cuda()
for rep in range(reps):
    r = eng.solve_local_batch(model, params, y0, ns, tt, want=want, target=target)
    nl, ms = eng.last_launch_info()
    print(f"{model}-{ns} B={B} {want}: kernel {ms:.3f} ms -> {B/ms*1e3:.4g} solves/s steps {r["nsteps"].double().mean().item():.1f} nrej {r["nrej"].double().mean().item():.1f}", flush=True)
