"""Dev script (GPU box): a short bounded fit (for ncu captures of nlls_step_kernel)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
B, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 48000, int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
model, ns = "distmod", 4
n, P, L = pk.local_dims(model, ns, 14)
rng = np.random.default_rng(1)
y0 = np.asarray(initial_condition(ns, model))
tgt = eng.solve_local_batch(model, rng.uniform(0.3, 2.0, (1, P)), y0, ns, T, want=("flat",))["flat"][0]
r = eng.nlls_local_batch(model, rng.uniform(0.05, 3.0, (B, P)), y0, ns, T, tgt, np.full(P, 1e-2), np.full(P, 20.0), max_iter=iters)
print("launches, device ms:", eng.last_launch_info(), "status", np.unique(r["status"], return_counts=True))
