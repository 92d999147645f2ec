import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import local_models as om
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
eng = pk.get_engine(0)
T = om.TIME_POINTS
model, ns, B = "randmod", 6, 4000
n, P, L = pk.local_dims(model, ns, 14)
rng = np.random.default_rng(3)
y0 = np.asarray(initial_condition(ns, model))
p = rng.uniform(0.05, 3.0, (B, P))
ref = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",), rtol=1e-10, atol=1e-14)
a = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",))
ratio = (np.abs(a["sol"] - ref["sol"]) / (1e-6 * np.abs(ref["sol"]) + 1e-9)).max(axis=(1, 2))
bad = np.argsort(-ratio)[:4]
print("worst systems", bad, ratio[bad], "steps default", a["nsteps"][bad], "tight", ref["nsteps"][bad], ref["nrej"][bad], "status", ref["status"][bad], a["status"][bad])
for b in list(bad) + [0, 1]:
    ex = om.exact_linear(model, p[b], y0, ns, T)
    ea = np.abs(a["sol"][b] - ex) / (1e-6 * np.abs(ex) + 1e-9)
    er = np.abs(ref["sol"][b] - ex) / (1e-6 * np.abs(ex) + 1e-9)
    k, i = np.unravel_index(np.argmax(er), er.shape)
    print(f"sys {b}: default vs exact {ea.max():.3g} | tight vs exact {er.max():.3g} at t[{k}] state {i}: exact {ex[k, i]:.6e} tight {ref['sol'][b][k, i]:.6e} default {a['sol'][b][k, i]:.6e}")
print("---- method comparison on the worst systems")
for method, rt, at in (("ros5l", None, None), ("rodas4", 1e-7, 1e-10), ("ros5l", 1e-7, 1e-10)):
    r = eng.solve_local_batch(model, p[bad], y0, ns, T, want=("sol",), method=method, rtol=rt, atol=at)
    for j, b in enumerate(bad):
        ex = om.exact_linear(model, p[b], y0, ns, T)
        e = np.abs(r["sol"][j] - ex) / (1e-6 * np.abs(ex) + 1e-9)
        print(method, rt, "sys", b, f"err {e.max():.3g} steps {r['nsteps'][j]} rej {r['nrej'][j]}; per-time max:", np.array2string(e.max(axis=1), precision=2, max_line_width=200))
