"""Dev script (GPU box): end-to-end (host buffers) time of the bench workload for the library's host pipeline (the chunk
count / growth scan recorded in pk_api.cu was run with temporary environment knobs that have since been removed)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import phoskintime_b200 as pk
from phoskintime_b200.models import succmod
from phoskintime_b200.steady import initial_condition
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
B = 1000000
ph = torch.empty((B, 14), dtype=torch.float64).pin_memory(); ph.numpy()[:] = np.random.default_rng(2).uniform(0.05, 3.0, (B, 14))
y0 = np.asarray(initial_condition(5, "succmod"))
target = np.random.default_rng(1).random(93)
out = {k: torch.empty(B, dtype=torch.float64).pin_memory().numpy() for k in ("ssr", "score")}
out["status"] = torch.empty(B, dtype=torch.int32).pin_memory().numpy()
f = lambda: succmod.solve_ode_batch(ph.numpy(), y0, 5, T, want=("ssr", "score"), target=target, out=out, counters=False)
for _ in range(3): f()
ts = []
for _ in range(10):
    t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
print(f"e2e {np.mean(ts) * 1e3:.3f} ms (min {np.min(ts) * 1e3:.3f}) -> {B / np.mean(ts):.4g} solves/s")
