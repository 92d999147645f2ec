"""TEST INFRASTRUCTURE ONLY — writes tests/golden/global_*.npz from the UNMODIFIED reference.

Run in the build container:  python oracle/gen_golden_global.py   (or `... only 2` for one kinetic model)
For each kinetic model (0 distributive, 1 sequential, 2 combinatorial, 4 saturating — `MODEL` is an import-time
constant of the reference, so each runs in its own subprocess) a REAL `global_model.network.System`
is constructed around a small synthetic topology (the same arrays `phoskintime_b200.global_model.
synthetic_system` builds), and the reference's own `simulate_odeint` (LSODA + finite-difference
Jacobian, global_model/simulate.py:34-80), `LOSS_FN` (lossfn.py) for all 8 loss modes and
`System.odeint_args` are executed for several parameter vectors.  A tight solution (reference RHS,
bucket-by-bucket restart at 1e-12) is stored next to the stock one.
"""
import os
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
T_EVAL = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 15.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
CASES = [(0, 10, 5, 3, 11), (0, 36, 12, 4, 12), (1, 14, 6, 4, 13), (4, 14, 6, 3, 14),   # model, N, K, max_sites, seed
         (0, 120, 40, 4, 5, 2),     # BASELINE configs[4] shape (state_dim ~ 500), 2 parameter vectors
         (2, 10, 5, 3, 21), (2, 24, 8, 4, 22),   # combinatorial: one state per phosphorylation pattern
         (2, 7, 4, 6, 27),          # combinatorial blocks of 32 and 64 patterns (sites [1 5 0 2 6 1 4]): the warp-per-block inverse
         (0, 300, 60, 4, 6, 2)]     # capacity case: 1371 states, 254 regulators -> Schur block and stage vectors leave shared memory


def stub_modules(model, loss_mode):
    import logging
    import ref_shim
    os.environ.setdefault("NUMBA_CACHE_DIR", f"/tmp/phoskin_numba_cache_global_{model}_{loss_mode}")
    import numba
    numba.config.CACHE_DIR = os.environ["NUMBA_CACHE_DIR"]
    lg = lambda *a, **k: logging.getLogger("phoskin_ref")
    cfg = types.ModuleType("config"); cfg.__path__ = []
    sys.modules["config"] = cfg
    cc = types.ModuleType("config.config"); cc.setup_logger = lg
    sys.modules["config.config"] = cc
    gm = types.ModuleType("global_model"); gm.__path__ = [os.path.join(ref_shim.REF_ROOT, "global_model")]
    sys.modules["global_model"] = gm
    gc = types.ModuleType("global_model.config")
    gc.TIME_POINTS_PROTEIN = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
    gc.MODEL, gc.LOSS_MODE, gc.RESULTS_DIR, gc.USE_CUSTOM_SOLVER = model, loss_mode, "/tmp", False
    sys.modules["global_model.config"] = gc


def build_reference_system(model, N, K, max_sites, seed):
    """(host mirror system, UNMODIFIED reference System, reference modules) around one synthetic topology; checks that the
    reference's own odeint_args() equals the host mirror's, array by array."""
    from scipy import sparse
    from phoskintime_b200.global_model import synthetic_system
    stub_modules(model, 0)
    import importlib
    network = importlib.import_module("global_model.network")
    simulate = importlib.import_module("global_model.simulate")
    jac = importlib.import_module("global_model.jacspeedup")
    s = synthetic_system(seed=seed, N=N, K=K, max_sites=max_sites, model=model)
    ref, args = _reference_system(model, N, K, s, network, jac, sparse)
    return s, ref, simulate, jac


def _reference_system(model, N, K, s, network, jac, sparse):
    names = [f"P{i:03d}" for i in range(N)]
    kin_names = [f"K{k:03d}" for k in range(K)]
    p2i = {n: i for i, n in enumerate(names)}
    for i in range(N):                       # driven proteins ARE kinases: same name in both maps
        if s.driver_map[i] >= 0:
            kin_names[s.driver_map[i]] = names[i]
    idx = types.SimpleNamespace(N=N, proteins=names, kinases=kin_names, p2i=p2i,
                                k2i={k: i for i, k in enumerate(kin_names)}, proxy_map={},
                                offset_y=s.idx.offset_y.copy(), offset_s=s.idx.offset_s.copy(),
                                n_sites=s.idx.n_sites.copy(), state_dim=s.idx.state_dim,
                                total_sites=s.idx.total_sites)
    if model == 2:
        idx.n_states = s.idx.n_states.copy()
    W = sparse.csr_matrix((s.W_data, s.W_indices, s.W_indptr), shape=(s.idx.total_sites, K))
    TF = sparse.csr_matrix((s.TF_data, s.TF_indices, s.TF_indptr), shape=(N, N))
    kin = types.SimpleNamespace(grid=s.kin_grid.copy(), Kmat=s.kin_Kmat.copy())
    ref = network.System(idx, W, TF, kin, {**s.defaults}, s.tf_deg.copy())
    if model == 2:
        jac.build_S_cache_into(ref.S_cache, ref.W_indptr, ref.W_indices, ref.W_data, ref.kin_Kmat, ref.c_k)
        args, mine = ref.odeint_args(ref.S_cache), s.odeint_args(s.build_S_cache())
        assert len(args) == len(mine) == 27 and np.array_equal(args[24], s.driver_map)
        assert np.allclose(args[9], mine[9], rtol=1e-14, atol=0.0), "S_cache mismatch"
        for k, (a, b) in enumerate(zip(args, mine)):
            assert k == 9 or np.array_equal(np.asarray(a), np.asarray(b)), f"odeint_args wire format mismatch at {k}"
    else:
        args = ref.odeint_args()
        assert np.array_equal(args[22], s.driver_map), "driver_map built by the reference differs"
        for a, b in zip(args, s.odeint_args()):
            assert np.array_equal(np.asarray(a), np.asarray(b)), "odeint_args wire format mismatch"

    return ref, args


def run_case(model, N, K, max_sites, seed, B=4):
    from scipy.integrate import odeint
    from phoskintime_b200.global_model import synthetic_loss_data
    s, ref, simulate, jac = build_reference_system(model, N, K, max_sites, seed)
    rng = np.random.default_rng(seed + 100)
    P = np.empty((B, s.n_params))
    Ys, Yt = [], []
    for b in range(B):
        p = {k: v * np.exp(0.25 * rng.standard_normal(v.shape)) for k, v in s.defaults.items() if k != "tf_scale"}
        p["tf_scale"] = s.defaults["tf_scale"] * float(np.exp(0.2 * rng.standard_normal()))
        if b == 0:
            p = {**s.defaults}
        P[b] = s.pack_params(p)
        ref.update(**p)
        Ys.append(simulate.simulate_odeint(ref, T_EVAL, rtol=1e-8, atol=1e-8, mxstep=200000))
        a = ref.odeint_args(ref.S_cache) if model == 2 else ref.odeint_args()    # S_cache: refreshed by simulate_odeint
        stops = np.unique(np.concatenate([T_EVAL, s.kin_grid]))
        y = ref.y0()
        rows = {0.0: y.copy()}
        for lo, hi in zip(stops[:-1], stops[1:]):
            mid = 0.5 * (lo + hi)
            y = odeint(lambda yy, tt, *aa: jac.rhs_odeint(yy, mid, *aa), y, [lo, hi], args=a, rtol=1e-12,
                       atol=1e-13, mxstep=500000)[-1]
            rows[float(hi)] = y.copy()
        Yt.append(np.array([rows[float(t)] for t in T_EVAL]))
    Ys, Yt = np.array(Ys), np.array(Yt)
    ld = synthetic_loss_data(s, T_EVAL, seed=seed + 7)
    np.savez_compressed(os.path.join(OUT, f"global_m{model}_N{N}.npz"), model=model, N=N, K=K, max_sites=max_sites,
                        seed=seed, t=T_EVAL, params=P, Y=Ys, Y_tight=Yt, y0=ref.y0(),
                        **{f"ld_{k}": np.asarray(v) for k, v in ld.items()})
    print(f"model {model} N={N} state_dim={s.idx.state_dim}: stock vs tight max "
          f"{float((np.abs(Ys - Yt) / (1e-6 * np.abs(Yt) + 1e-9)).max()):.3g} of (1e-6 rel + 1e-9)")


def run_losses(model, N, loss_mode):
    """LOSS_FN closes over LOSS_MODE at jit time -> one process per mode."""
    stub_modules(model, loss_mode)
    import importlib
    lossfn = importlib.import_module("global_model.lossfn")
    g = np.load(os.path.join(OUT, f"global_m{model}_N{N}.npz"))
    ld = {k[3:]: g[k] for k in g.files if k.startswith("ld_")}
    out = []
    for Y in g["Y"]:
        out.append(lossfn.LOSS_FN(np.ascontiguousarray(Y), ld["p_prot"], ld["t_prot"], ld["obs_prot"], ld["w_prot"],
                                  ld["p_rna"], ld["t_rna"], ld["obs_rna"], ld["w_rna"], ld["p_pho"], ld["s_pho"],
                                  ld["t_pho"], ld["obs_pho"], ld["w_pho"], ld["prot_map"], int(ld["prot_base_idx"]),
                                  int(ld["rna_base_idx"]), int(ld["pho_base_idx"])))
    np.save(os.path.join(OUT, f"_tmp_loss_m{model}_N{N}_mode{loss_mode}.npy"), np.array(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "case":
        run_case(*[int(x) for x in sys.argv[2:8]])
    elif len(sys.argv) > 1 and sys.argv[1] == "loss":
        run_losses(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    else:
        # `python oracle/gen_golden_global.py only 2` regenerates the cases of one kinetic model
        keep = (lambda c: c[0] == int(sys.argv[2])) if len(sys.argv) > 2 and sys.argv[1] == "only" else (lambda c: True)
        os.makedirs(OUT, exist_ok=True)
        for c in filter(keep, CASES):
            subprocess.run([sys.executable, __file__, "case"] + [str(x) for x in c], check=True)
        for model, N, *_ in filter(keep, CASES[:1] + CASES[2:3] + CASES[5:6]):
            losses = {}
            for mode in (0, 1, 2, 3, 4, 5, 6, -1):
                subprocess.run([sys.executable, __file__, "loss", str(model), str(N), str(mode)], check=True)
                f = os.path.join(OUT, f"_tmp_loss_m{model}_N{N}_mode{mode}.npy")
                losses[f"loss_mode{mode}"] = np.load(f)
                os.remove(f)
            f = os.path.join(OUT, f"global_m{model}_N{N}.npz")
            g = dict(np.load(f))
            g.update(losses)
            np.savez_compressed(f, **g)
        print("done")
