"""TEST INFRASTRUCTURE ONLY — RHS / Jacobian golden vectors of the global path from the UNMODIFIED reference.

Run in the build container:  python oracle/gen_golden_rhs.py  ->  tests/golden/globalrhs_m{model}_N{N}.npz
For each kinetic model (own subprocess: `MODEL` is an import-time constant of the reference):
  * `rhs_odeint(y, t, *args)` (global_model/jacspeedup.py:175-375) and the finite-difference Jacobian
    `fd_jacobian_odeint(y, t, *args)` (:397-588) at states taken from the stored trajectories, at times in different
    kinase buckets, for two parameter vectors;
  * the `fun(t, y)` closures of global_model/model_ivp.py:49-277 (`make_solve_ivp_fun_*`) with caller-supplied TF
    inputs and phosphorylation rates.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
CASES = [(0, 10, 5, 3, 11), (1, 14, 6, 4, 13), (2, 10, 5, 3, 21), (4, 14, 6, 3, 14)]     # model, N, K, max_sites, seed


def run_case(model, N, K, max_sites, seed):
    import importlib
    import gen_golden_global as gg
    s, ref, simulate, jac = gg.build_reference_system(model, N, K, max_sites, seed)
    ivp = importlib.import_module("global_model.model_ivp")
    g = np.load(os.path.join(OUT, f"global_m{model}_N{N}.npz"))
    rng = np.random.default_rng(seed + 500)
    P, Y, T, F, J = [], [], [], [], []
    for b in (0, 2):
        ref.update(**s.unpack_params(g["params"][b]))
        if model == 2:
            jac.build_S_cache_into(ref.S_cache, ref.W_indptr, ref.W_indices, ref.W_data, ref.kin_Kmat, ref.c_k)
        args = ref.odeint_args(ref.S_cache) if model == 2 else ref.odeint_args()
        for k, t in ((0, 0.0), (3, 0.9), (6, 8.0), (9, 45.0), (12, 300.0), (14, 2000.0)):
            y = np.ascontiguousarray(g["Y"][b][k] * np.exp(0.1 * rng.standard_normal(g["Y"].shape[2])))
            P.append(g["params"][b]); Y.append(y); T.append(t)
            F.append(np.array(jac.rhs_odeint(y, t, *args), copy=True))
            J.append(np.array(jac.fd_jacobian_odeint(y, t, *args), copy=True))
    # model_ivp closures: TF inputs and S_all supplied by the caller
    p = s.unpack_params(g["params"][1])
    S_all = rng.uniform(0.05, 2.0, s.idx.total_sites)
    tf_in = rng.uniform(-0.8, 1.5, N)
    kw = dict(A_i=p["A_i"], B_i=p["B_i"], C_i=p["C_i"], D_i=p["D_i"], Dp_i=p["Dp_i"], E_i=p["E_i"], tf_scale=p["tf_scale"],
              tf_input=tf_in, offset_y=s.idx.offset_y, offset_s=s.idx.offset_s, n_sites=s.idx.n_sites)
    if model == 2:
        S_cache = np.repeat(S_all[:, None], 3, axis=1) * np.array([1.0, 0.5, 2.0])[None, :]
        fun = ivp.make_solve_ivp_fun_combinatorial(S_cache=S_cache, jb=1, n_states=s.idx.n_states, trans_from=s.trans_from,
                                                   trans_to=s.trans_to, trans_site=s.trans_site, trans_off=s.trans_off,
                                                   trans_n=s.trans_n, **kw)
        extra = {"ivp_S_cache": S_cache, "ivp_jb": 1}
    else:
        maker = {0: ivp.make_solve_ivp_fun_distributive, 1: ivp.make_solve_ivp_fun_sequential,
                 4: ivp.make_solve_ivp_fun_saturating}[model]
        fun = maker(S_all=S_all, **kw)
        extra = {"ivp_S_all": S_all}
    # the reference's custom DOPRI5 solver (jacspeedup.py:31-67 -> solvers.py:292-758) at two tolerance pairs
    custom = {}
    for tag, (rt, at) in (("a", (1e-5, 1e-7)), ("b", (1e-4, 1e-6))):
        outs = []
        for b in (0, 1, 3):
            ref.update(**s.unpack_params(g["params"][b]))
            outs.append(np.array(jac.solve_custom(ref, ref.y0(), g["t"], rt, at), copy=True))
        custom[f"custom_Y_{tag}"] = np.array(outs)
        custom[f"custom_tol_{tag}"] = np.array([rt, at])
    ivp_Y = np.array([g["Y"][1][k] for k in (0, 5, 10)])
    ivp_F = np.array([fun(1.0, np.ascontiguousarray(y)) for y in ivp_Y])
    np.savez_compressed(os.path.join(OUT, f"globalrhs_m{model}_N{N}.npz"), model=model, N=N, K=K, max_sites=max_sites, seed=seed,
                        params=np.array(P), Y=np.array(Y), t=np.array(T), f=np.array(F), J_fd=np.array(J),
                        ivp_params=g["params"][1], ivp_tf=tf_in, ivp_Y=ivp_Y, ivp_f=ivp_F, custom_rows=np.array([0, 1, 3]), custom_t=g["t"], **custom, **extra)
    print(f"model {model} N={N}: |f| max {np.abs(np.array(F)).max():.3g}, J nnz fraction {np.mean(np.abs(np.array(J)) > 0):.3f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "case":
        run_case(*[int(x) for x in sys.argv[2:7]])
    else:
        for c in CASES:
            subprocess.run([sys.executable, __file__, "case"] + [str(x) for x in c], check=True)
        print("done")
