"""Dev script (GPU box): throughput of the batched multistart fit at the BASELINE configs[3] shape
(distributive, 4 sites, `proteins` x `starts` problems) -> fits/s, solves/s, iterations."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import phoskintime_b200 as pk
from phoskintime_b200 import paramest
from phoskintime_b200.steady import initial_condition
G, starts = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 48
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
model, ns = "distmod", 4
n, P, L = pk.local_dims(model, ns, 14)
rng = np.random.default_rng(4)
y0 = np.asarray(initial_condition(ns, model))
truth = rng.uniform(0.05, 3.0, (G, P))
targets = eng.solve_local_batch(model, truth, y0, ns, T, want=("flat",))["flat"] * (1.0 + 0.05 * rng.standard_normal((G, L)))
lb, ub = np.full(P, 1e-2), np.full(P, 20.0)
out = []
for max_iter in (50, 100):
    t0 = time.perf_counter()
    fit = paramest.fit_multistart(model, np.ones(P), lb, ub, y0, ns, T, targets, genes=[f"G{p}" for p in range(G)],
                                  n_starts=starts, engine=eng, max_iter=max_iter)
    wall = time.perf_counter() - t0
    nl, ms = eng.last_launch_info()
    B = fit["theta"].shape[0]
    st, cnt = np.unique(fit["status"], return_counts=True)
    row = {"proteins": G, "starts": B // G, "problems": B, "max_iter": max_iter, "wall_s": wall, "device_ms": ms, "launches": nl,
           "solves": int(fit["nfev"].sum()), "solves_per_s": float(fit["nfev"].sum() / (ms * 1e-3)), "fits_per_s": B / (ms * 1e-3),
           "iters_mean": float(fit["iters"].mean()), "iters_max": int(fit["iters"].max()),
           "status_counts": {int(a): int(b) for a, b in zip(st, cnt)}, "median_best_score": float(np.median(fit["best_score"]))}
    print(json.dumps(row), flush=True)
    out.append(row)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/nlls_bench.json", "w"), indent=1)
