"""Dev script (GPU box): tiny batches through every kernel family, meant to run under
`compute-sanitizer --tool racecheck|memcheck` (a few systems each so that the instrumented run ends in minutes)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
from phoskintime_b200.global_model import simulate_batch, synthetic_system, synthetic_loss_data, metric_time_indices
which = set(sys.argv[1:]) or {"tps", "dense", "global", "comb", "nlls", "morris"}
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
rng = np.random.default_rng(0)
def local(model, ns, B, want):
    n, P, L = pk.local_dims(model, ns, 14)
    p = rng.uniform(0.05, 3.0, (B, P))
    y0 = np.asarray(initial_condition(ns, model))
    tgt = eng.solve_local_batch(model, p[:1], y0, ns, T, want=("flat",))["flat"][0]
    r = eng.solve_local_batch(model, p, y0, ns, T, want=want, target=tgt)
    print(model, ns, "status", np.unique(r["status"]), flush=True)
if "tps" in which:
    local("distmod", 3, 300, ("sol", "flat", "Y", "ssr", "score"))
    local("succmod", 5, 300, ("ssr", "score"))
if "dense" in which:
    local("randmod", 3, 12, ("sol", "flat"))       # 1 warp per system
    local("randmod", 5, 8, ("flat",))              # 2 warps
    local("randmod", 6, 6, ("flat", "score"))      # 4 warps
    local("distmod", 10, 8, ("flat",))
T15 = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
def glob(model, N, K, B, force=False):
    s = synthetic_system(seed=3, N=N, K=K, max_sites=3, model=model)
    P = s.pack_params()[None, :] * np.exp(0.05 * rng.standard_normal((B, s.n_params)))
    ld = synthetic_loss_data(s, T15, seed=1)
    mt = metric_time_indices(T15, T, [4.0, 8.0, 15.0, 30.0, 60.0], T)
    sub = {k: ld[k] for k in ("prot_base_idx", "rna_base_idx", "pho_base_idx")}
    for mod, keys in (("prot", ("p_prot", "t_prot", "obs_prot", "w_prot")), ("rna", ("p_rna", "t_rna", "obs_rna", "w_rna")),
                      ("pho", ("p_pho", "s_pho", "t_pho", "obs_pho", "w_pho"))):
        keep = ld["t_" + mod] < 8
        for k in keys:
            sub[k] = ld[k][keep]
    mts = {**mt, **{k: mt[k][mt[k] < 8] for k in ("t_prot", "t_rna", "t_pho")}}
    if force:
        topo = eng.global_upload(s, force_generic=True)
        r = eng.global_solve_batch(topo, P, s.y0(), T15[:6], ("Y",))
    else:
        r = simulate_batch(s, P, T15[:8], ("Y", "loss", "metric"), loss_data=sub, metric_times=mts, engine=eng)
    print("global model", model, "N", N, "status", np.unique(r["status"]), flush=True)
if "global" in which:
    glob(0, 12, 5, 3)
    glob(1, 40, 10, 2)
    glob(4, 12, 5, 2, force=True)
if "comb" in which:
    glob(2, 12, 5, 3)
if "nlls" in which:
    model, ns = "distmod", 2
    n, P, L = pk.local_dims(model, ns, 14)
    y0 = np.asarray(initial_condition(ns, model))
    th = rng.uniform(0.3, 2.0, P)
    tgt = eng.solve_local_batch(model, th[None], y0, ns, T, want=("flat",))["flat"][0]
    r = eng.nlls_local_batch(model, th * np.exp(0.1 * rng.standard_normal((5, P))), y0, ns, T, tgt, np.full(P, 1e-2), np.full(P, 20.0), max_iter=4)
    print("nlls status", r["status"], flush=True)
if "morris" in which:
    X = rng.uniform(0, 1, (33, 10)); Y = rng.uniform(0, 1, 33)
    print("morris", eng.morris_ee(X, Y, 4)["mu_star"][:2], flush=True)
