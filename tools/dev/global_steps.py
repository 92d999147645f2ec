"""Dev script (GPU box): accepted / rejected step counts of the global kernel at N = 120 and where they fall (per stop interval)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import phoskintime_b200 as pk
from phoskintime_b200.global_model import simulate_batch, synthetic_system
eng = pk.get_engine(0)
s = synthetic_system(seed=5, N=120, K=40, max_sites=4, model=0)
rng = np.random.default_rng(0)
base = s.pack_params()
P = base[None, :] * np.exp(0.05 * rng.standard_normal((296, base.size)))
T15 = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for rtol, atol in ((2e-6, 2e-9), (1e-5, 1e-8)):
    r = simulate_batch(s, P, T15, ("Y",), engine=eng, rtol=rtol, atol=atol)
    print(f"rtol {rtol}: accepted {r['nsteps'].mean():.1f} rejected {r['nrej'].mean():.1f} kernel {eng.last_launch_info()[1]:.1f} ms")
# steps per interval: solve to successive end times and difference the counters
prev = 0.0
for k in range(2, len(T15) + 1):
    r = simulate_batch(s, P[:32], T15[:k], ("Y",), engine=eng)
    tot = (r["nsteps"] + r["nrej"]).mean()
    print(f"  up to t = {T15[k - 1]:7.2f}: {tot:7.1f} attempts (+{tot - prev:6.1f}), rejected so far {r['nrej'].mean():.1f}")
    prev = tot
