"""Dev (GPU box): pure relative error of the bench workload vs the absolute tolerance."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for model, ns, seed in (("succmod", 5, 2), ("distmod", 3, 11)):
    B = 1_000_000
    n, P, L = pk.local_dims(model, ns, 14)
    p = torch.from_numpy(np.random.default_rng(seed).uniform(0.05, 3.0, (B, P))).cuda()
    y0 = torch.tensor(initial_condition(ns, model)).cuda()
    tt = torch.from_numpy(T).cuda()
    tight = eng.solve_local_batch(model, p, y0, ns, tt, want=("sol",), method="rodas4", rtol=1e-10, atol=1e-14)["sol"]
    print(model, ns, "min state", float(tight[:, 1:].min()))
    for rtol, atol in ((2e-5, 2e-9), (2e-5, 5e-10), (2e-5, 1e-10), (2e-5, 2e-11), (1e-5, 1e-10), (1e-5, 1e-11), (5e-6, 1e-11)):
        for rep in range(2):
            r = eng.solve_local_batch(model, p, y0, ns, tt, want=("sol",), rtol=rtol, atol=atol)
        ms = eng.last_launch_info()[1]
        r2 = eng.solve_local_batch(model, p, y0, ns, tt, want=("score",), target=torch.rand(L, dtype=torch.float64).cuda(), rtol=rtol, atol=atol)
        r2 = eng.solve_local_batch(model, p, y0, ns, tt, want=("score",), target=torch.rand(L, dtype=torch.float64).cuda(), rtol=rtol, atol=atol)
        ms2 = eng.last_launch_info()[1]
        d = (r["sol"] - tight).abs()
        print(f"  rtol {rtol:g} atol {atol:g}: steps {r['nsteps'].double().mean().item():.1f} rel {float((d / tight.abs().clamp_min(1e-12)).max()):.3g} "
              f"bound {float((d / (1e-6 * tight.abs() + 1e-9)).max()):.3g} scalar-kernel {ms2:.3f} ms", flush=True)
