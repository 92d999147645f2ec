"""Dev script (GPU box): per-phase cycle shares of the global-network step loop (needs `make trace`).
   PHOSKIN_LIB=phoskintime_b200/libphoskin_b200_trace.so python tools/trace_global.py [N K B [model]]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.global_model import simulate_batch, synthetic_system

N, K, B = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (120, 40, 148)
MODEL = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng = pk.get_engine(0)
t = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
s = synthetic_system(seed=5, N=N, K=K, max_sites=4, model=MODEL)
rng = np.random.default_rng(0)
base = s.pack_params()
P = base[None, :] * np.exp(0.05 * rng.standard_normal((B, base.size)))
names = ["setup", "rhs+factor blocks", "schur assemble", "schur invert", "scale U1", "stage combine", "stage rhs", "stage scale",
         "-", "error+control", "solve: blocks", "solve: schur apply", "solve: correction"]
buf = (C.c_ulonglong * 16)()
for rep in range(2):
    r = simulate_batch(s, P, t, ("Y",), engine=eng)
    eng.lib.pk_global_trace_read(buf)
    steps = int(r["nsteps"][0] + r["nrej"][0])
    ms = eng.last_launch_info()[1]
tot = sum(buf)
print(f"N={N} B={B}: kernel {ms:.1f} ms, CTA 0 first system {steps} steps; traced cycles {tot} ({tot / max(steps, 1):.0f} per step)")
for i, nme in enumerate(names):
    if buf[i]:
        print(f"  {nme:22s} {100.0 * buf[i] / tot:5.1f} %   {buf[i] / max(steps, 1):9.0f} cycles/step")
