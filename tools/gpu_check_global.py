"""Dev script (GPU box): global-network kernel vs the committed goldens + timings at the BASELINE configs[4] shape."""
import glob, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.global_model import simulate_batch, synthetic_system

eng = pk.get_engine(0)
print("device", eng.device_name, eng.sm_count, "SMs", flush=True)
small = len(sys.argv) > 1 and sys.argv[1] == "small"
for f in sorted(glob.glob("tests/golden/global_*.npz")):
    g = np.load(f)
    if small and int(g["N"]) > 12:
        continue
    s = synthetic_system(seed=int(g["seed"]), N=int(g["N"]), K=int(g["K"]), max_sites=int(g["max_sites"]), model=int(g["model"]))
    for rtol, atol in ((1e-5, 1e-8), (3e-6, 3e-9), (2e-6, 2e-9), (1e-6, 1e-9), (1e-7, 1e-10)):
        t0 = time.time()
        r = simulate_batch(s, g["params"], g["t"], ("Y",), y0=g["y0"], rtol=rtol, atol=atol, engine=eng)
        dt = time.time() - t0
        Y, tight, stock = r["Y"], g["Y_tight"], g["Y"]
        e_t = np.abs(Y - tight) / (1e-6 * np.abs(tight) + 1e-9)
        e_s = np.abs(Y - stock) / (1e-6 * np.abs(stock) + 1e-7)
        e_st = np.abs(stock - tight) / (1e-6 * np.abs(tight) + 1e-9)
        print(f"{os.path.basename(f)[7:-4]:10s} rtol {rtol:g} status {np.bincount(r['status'], minlength=4)} steps {r['nsteps'].mean():.0f} rej {r['nrej'].mean():.1f} "
              f"| vs tight {np.nanmax(e_t):.3g} | vs stock {np.nanmax(e_s):.3g} | stock vs tight {e_st.max():.3g} | dims {eng.global_dims(s._topo_id[eng.token])} [{dt*1e3:.1f} ms, kernel {eng.last_launch_info()[1]:.2f} ms]", flush=True)
if small:
    sys.exit(0)
t = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for N, K, B in ((36, 12, 1184), (120, 40, 148), (120, 40, 1184)):
    s = synthetic_system(seed=5, N=N, K=K, max_sites=4, model=0)
    rng = np.random.default_rng(0)
    base = s.pack_params()
    P = base[None, :] * np.exp(0.05 * rng.standard_normal((B, base.size)))
    for rtol, atol in ((1e-5, 1e-8), (1e-6, 1e-9)):
        r = simulate_batch(s, P, t, ("metric",), rtol=rtol, atol=atol, engine=eng,
                           metric_times={"t_prot": np.arange(15), "t_rna": np.arange(5, 15), "t_pho": np.arange(15), "prot_b": 0, "rna_b": 5, "pho_b": 0})
        ms = eng.last_launch_info()[1]
        print(f"N={N} B={B} rtol {rtol:g}: dims {eng.global_dims(s._topo_id[eng.token])} status {np.bincount(r['status'], minlength=4)} steps {r['nsteps'].mean():.0f} rej {r['nrej'].mean():.1f} "
              f"kernel {ms:.1f} ms -> {B / ms * 1e3:.0f} solves/s, {ms * 1e3 / (r['nsteps'] + r['nrej']).sum() * min(B, 148):.1f} us/step/CTA", flush=True)
