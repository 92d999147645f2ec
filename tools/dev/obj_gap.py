"""Dev (GPU box): gap between the kernel's fused objectives / fold-change tables / Morris scalar and the reference goldens."""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle_objectives import load_case, FILES, KEYS
import phoskintime_b200 as pk
from phoskintime_b200.global_model import GlobalODE_MOO, fold_change_tables, simulate_batch, metric_time_indices
eng = pk.get_engine(0)
for path in FILES:
    g, s, ld, defaults, slices = load_case(path)
    lam = dict(zip(("protein", "rna", "phospho", "prior"), g["lambdas"]))
    prob = GlobalODE_MOO(s, slices, ld, defaults, lam, g["t_grid"], engine=eng)
    F = prob.evaluate_batch(g["theta"])
    print(os.path.basename(path), "F rel", np.max(np.abs(F - g["F"]) / np.abs(g["F"])))
    tab = fold_change_tables(s, g["phys"], g["t_prot"], g["t_rna"], g["t_prot"], engine=eng)
    for k in ("fc_prot", "fc_rna", "fc_pho"):
        a = tab[k].reshape(g[k].shape)
        print("   ", k, "rel", np.max(np.abs(a - g[k]) / np.abs(g[k])))
    mt = metric_time_indices(g["t_grid"], g["t_prot"], g["t_rna"], g["t_prot"])
    for m, name in enumerate(g["metric_names"]):
        r = simulate_batch(s, g["phys"], g["t_grid"], ("metric",), rtol=1e-5, atol=1e-7, mxstep=5000, metric=str(name), metric_times=mt, engine=eng)
        print("   ", name, "rel", np.max(np.abs(r["metric"] - g["metrics"][:, m]) / np.abs(g["metrics"][:, m])))
