"""Dev script (GPU box): dense (random-model) kernel throughput and error vs the tight run for the current step grid."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for model, ns, B in (("randmod", 6, 16384), ("randmod", 5, 32768), ("randmod", 4, 65536)):
    n, P, L = pk.local_dims(model, ns, 14)
    rng = np.random.default_rng(3)
    y0 = np.asarray(initial_condition(ns, model))
    for name, p in (("U", rng.uniform(0.05, 3.0, (B, P))), ("logU", np.exp(rng.uniform(np.log(0.01), np.log(20.0), (B // 4, P))))):
        ref = eng.solve_local_batch(model, p[:2048], y0, ns, T, want=("sol",), rtol=1e-10, atol=1e-14)
        a = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",))
        a = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",))
        ms = eng.last_launch_info()[1]
        ratio = (np.abs(a["sol"][:2048] - ref["sol"]) / (1e-6 * np.abs(ref["sol"]) + 1e-9)).max()
        print(f"{model}-{ns} {name:5s} B={p.shape[0]}: {ms:8.2f} ms {p.shape[0] / ms * 1e3:10.4g} solves/s steps {a['nsteps'].mean():6.1f} rej {a['nrej'].mean():5.2f} fail {int((a['status'] != 0).sum())} err/bound {ratio:.3f}", flush=True)
