"""TEST INFRASTRUCTURE ONLY — pins the objective / observable layer to the UNMODIFIED reference.

Run in the build container:  python oracle/gen_golden_objectives.py
Writes tests/golden/globalobj_m{model}_N{N}.npz.  One subprocess per kinetic model (`MODEL` is an import-time constant
of the reference).  In each, around the same synthetic topology as oracle/gen_golden_global.py, the reference's own

  * `global_model.params.init_raw_params` / `unpack_params`            (params.py:25-132)
  * `global_model.optproblem.GlobalODE_MOO._evaluate`                    (optproblem.py:87-160)
  * `global_model.simulate.simulate_and_measure`                         (simulate.py:83-202)
  * `global_model.sensitivity._compute_scalar_metric`, `compute_bounds`  (sensitivity.py:39-140)

are executed unmodified.  pymoo, SALib, matplotlib and seaborn are absent from this image; they are only imported by the
reference files at module level (base class `ElementwiseProblem`, sampler, plotting), never by the functions above, so
three-line stand-ins are registered in `sys.modules` for the import to succeed.  The trajectories the reference
integrates inside these calls are captured (by wrapping the module attribute `simulate_odeint`, not by editing the
file) so that oracle/global_models.py can be held to the reference's numbers on the reference's own trajectories.
"""
import os
import subprocess
import sys
import tomllib
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
T_PROT = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
T_RNA = np.array([4.0, 8.0, 15.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
CASES = [(0, 10, 5, 3, 11), (1, 14, 6, 4, 13), (2, 10, 5, 3, 21), (4, 14, 6, 3, 14)]     # model, N, K, max_sites, seed
LAMBDAS = {"protein": 1.0, "rna": 0.5, "phospho": 2.0, "prior": 0.1}
METRICS = ("total_signal", "mean", "variance", "l2_norm")


def stub_everything(model):
    import gen_golden_global as gg
    import ref_shim
    gg.stub_modules(model, 0)
    gc = sys.modules["global_model.config"]
    cfg = tomllib.load(open(os.path.join(ref_shim.REF_ROOT, "config.toml"), "rb"))["global_model"]
    gc.BOUNDS_CONFIG = {k: (float(v[0]), float(v[1])) for k, v in cfg["bounds"].items()}
    gc.ODE_ABS_TOL = float(cfg["solver"]["absolute_tolerance"])
    gc.ODE_REL_TOL = float(cfg["solver"]["relative_tolerance"])
    gc.ODE_MAX_STEPS = int(cfg["solver"]["max_timesteps"])
    gc.TIME_POINTS_RNA, gc.TIME_POINTS_PHOSPHO = T_RNA.copy(), T_PROT.copy()
    gc.SENSITIVITY_TRAJECTORIES, gc.SENSITIVITY_LEVELS, gc.SENSITIVITY_PERTURBATION = 100, 40, 0.05
    gc.SENSITIVITY_TOP_CURVES, gc.SEED = 20, 42

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class ElementwiseProblem:                       # stand-in for pymoo.core.problem.ElementwiseProblem (base class only)
        def __init__(self, **kw):
            self.__dict__.update(kw)
    mod("pymoo"); mod("pymoo.core"); mod("pymoo.core.problem", ElementwiseProblem=ElementwiseProblem)
    mod("SALib"); mod("SALib.sample", morris=None); mod("SALib.analyze"); mod("SALib.analyze.morris", analyze=None)
    mod("matplotlib"); mod("matplotlib.pyplot"); mod("seaborn")
    return gc


def run_case(model, N, K, max_sites, seed, B=4):
    from scipy import sparse
    from phoskintime_b200.global_model import synthetic_loss_data, synthetic_system
    gc = stub_everything(model)
    import importlib
    network = importlib.import_module("global_model.network")
    simulate = importlib.import_module("global_model.simulate")
    params = importlib.import_module("global_model.params")
    optproblem = importlib.import_module("global_model.optproblem")
    sens = importlib.import_module("global_model.sensitivity")

    s = synthetic_system(seed=seed, N=N, K=K, max_sites=max_sites, model=model)
    names = [f"P{i:03d}" for i in range(N)]
    kin_names = [f"K{k:03d}" for k in range(K)]
    for i in range(N):
        if s.driver_map[i] >= 0:
            kin_names[s.driver_map[i]] = names[i]
    sites = [[f"S{j + 1}" for j in range(int(s.idx.n_sites[i]))] for i in range(N)]
    idx = types.SimpleNamespace(N=N, proteins=names, kinases=kin_names, p2i={n: i for i, n in enumerate(names)},
                                k2i={k: i for i, k in enumerate(kin_names)}, proxy_map={}, sites=sites,
                                offset_y=s.idx.offset_y.copy(), offset_s=s.idx.offset_s.copy(),
                                n_sites=s.idx.n_sites.copy(), state_dim=s.idx.state_dim, total_sites=s.idx.total_sites)
    if model == 2:
        idx.n_states = s.idx.n_states.copy()
    W = sparse.csr_matrix((s.W_data, s.W_indices, s.W_indptr), shape=(s.idx.total_sites, K))
    TF = sparse.csr_matrix((s.TF_data, s.TF_indices, s.TF_indptr), shape=(N, N))
    kin = types.SimpleNamespace(grid=s.kin_grid.copy(), Kmat=s.kin_Kmat.copy())
    ref = network.System(idx, W, TF, kin, {**s.defaults}, s.tf_deg.copy())

    # capture the trajectories the reference integrates inside _evaluate / simulate_and_measure
    captured = []
    real_sim = simulate.simulate_odeint

    def spy(*a, **k):
        Y = real_sim(*a, **k)
        captured.append(np.array(Y, copy=True))
        return Y

    # defaults clipped into the reference's bound box, as its runner does before init_raw_params
    defaults = {k: (np.clip(v, *gc.BOUNDS_CONFIG[k]) if k != "tf_scale" else float(np.clip(v, *gc.BOUNDS_CONFIG[k])))
                for k, v in s.defaults.items()}
    theta0, slices, xl, xu = params.init_raw_params(defaults)
    t_grid = np.unique(np.concatenate([T_PROT, T_RNA]))
    ld = synthetic_loss_data(s, t_grid, seed=seed + 7)
    prob = optproblem.GlobalODE_MOO(ref, slices, ld, defaults, LAMBDAS, t_grid, xl, xu)
    rng = np.random.default_rng(seed + 300)
    X = theta0[None, :] + 0.3 * rng.standard_normal((B, theta0.size))
    X[0] = theta0
    optproblem.simulate_odeint = spy
    F, Yobj, phys = [], [], []
    for b in range(B):
        out = {}
        prob._evaluate(X[b], out)
        F.append(out["F"])
        Yobj.append(captured.pop())
        p = params.unpack_params(X[b], slices)
        phys.append(np.concatenate([np.ravel(p[k]) for k in ("c_k", "A_i", "B_i", "C_i", "D_i", "Dp_i", "E_i")] + [[p["tf_scale"]]]))
    optproblem.simulate_odeint = real_sim

    # observables and the Morris scalar (the system carries the parameters of the last _evaluate)
    simulate.simulate_odeint = spy
    fcs, metrics, Ymeas = [], [], []
    for b in range(B):
        ref.update(**params.unpack_params(X[b], slices))
        dfp, dfr, dfph = simulate.simulate_and_measure(ref, idx, T_PROT, T_RNA, T_PROT)
        Ymeas.append(captured.pop())
        fcs.append((dfp["pred_fc"].to_numpy().copy(), dfr["pred_fc"].to_numpy().copy(), dfph["pred_fc"].to_numpy().copy()))
        metrics.append([float(sens._compute_scalar_metric(dfp, dfr, dfph, m)) for m in METRICS])
        if b == 0:
            layout = dict(fc_prot_protein=dfp["protein"].to_numpy().astype(str), fc_prot_time=dfp["time"].to_numpy(),
                          fc_rna_time=dfr["time"].to_numpy(), fc_pho_protein=dfph["protein"].to_numpy().astype(str),
                          fc_pho_psite=dfph["psite"].to_numpy().astype(str), fc_pho_time=dfph["time"].to_numpy())
    simulate.simulate_odeint = real_sim
    prob_def = sens.compute_bounds({k: (np.asarray(v) if k != "tf_scale" else float(v)) for k, v in defaults.items()})

    np.savez_compressed(
        os.path.join(OUT, f"globalobj_m{model}_N{N}.npz"), model=model, N=N, K=K, max_sites=max_sites, seed=seed,
        t_grid=t_grid, t_prot=T_PROT, t_rna=T_RNA, theta=X, theta0=theta0, xl=xl, xu=xu,
        slices=np.array([[slices[k].start, slices[k].stop] for k in ("c_k", "A_i", "B_i", "C_i", "D_i", "Dp_i", "E_i", "tf_scale")]),
        phys=np.array(phys), F=np.array(F), Y_obj=np.array(Yobj), Y_meas=np.array(Ymeas),
        lambdas=np.array([LAMBDAS[k] for k in ("protein", "rna", "phospho", "prior")]),
        ode_tol=np.array([gc.ODE_REL_TOL, gc.ODE_ABS_TOL]),
        fc_prot=np.array([f[0] for f in fcs]), fc_rna=np.array([f[1] for f in fcs]), fc_pho=np.array([f[2] for f in fcs]),
        metrics=np.array(metrics), metric_names=np.array(METRICS),
        bounds=np.array(prob_def["bounds"]), bound_names=np.array(prob_def["names"]),
        **{f"def_{k}": np.asarray(v) for k, v in defaults.items()},
        **{f"ld_{k}": np.asarray(v) for k, v in ld.items()}, **layout)
    print(f"model {model} N={N}: F[0] = {F[0]}, metrics[0] = {metrics[0]}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "case":
        run_case(*[int(x) for x in sys.argv[2:7]])
    else:
        os.makedirs(OUT, exist_ok=True)
        for c in CASES:
            subprocess.run([sys.executable, __file__, "case"] + [str(x) for x in c], check=True)
        print("done")
