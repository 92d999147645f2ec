"""Dev (GPU box): step counts of the succ-5 bench workload on coarser output grids (how much does landing cost?)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
B = 200000
for model, ns in (("succmod", 5), ("distmod", 3)):
    n, P, L = pk.local_dims(model, ns, 14)
    params = torch.from_numpy(np.random.default_rng(2).uniform(0.05, 3.0, (B, P))).cuda()
    y0 = torch.tensor(initial_condition(ns, model)).cuda()
    for grid in (T, T[[0, 3, 6, 9, 13]], T[[0, 13]]):
        tt = torch.from_numpy(np.ascontiguousarray(grid)).cuda()
        r = eng.solve_local_batch(model, params, y0, ns, tt, want=("sol",))
        print(model, ns, len(grid), "outputs: steps", r["nsteps"].double().mean().item(), "rej", r["nrej"].double().mean().item(), flush=True)
