"""Dev script (GPU box): one global-network launch at the BASELINE configs[4] shape (for ncu captures)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.global_model import simulate_batch, synthetic_system

N, K, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
model = int(sys.argv[5]) if len(sys.argv) > 5 else 0
eng = pk.get_engine(0)
t = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
s = synthetic_system(seed=5, N=N, K=K, max_sites=4, model=model)
rng = np.random.default_rng(0)
base = s.pack_params()
P = base[None, :] * np.exp(0.05 * rng.standard_normal((B, base.size)))
mt = {"t_prot": np.arange(15), "t_rna": np.arange(5, 15), "t_pho": np.arange(15), "prot_b": 0, "rna_b": 5, "pho_b": 0}
for _ in range(reps):
    r = simulate_batch(s, P, t, ("metric",), engine=eng, metric_times=mt)
    ms = eng.last_launch_info()[1]
    print(f"model {model} N={N} B={B}: kernel {ms:.1f} ms -> {B / ms * 1e3:.0f} solves/s steps {r['nsteps'].mean():.0f} rejected {r['nrej'].mean():.1f}", flush=True)
