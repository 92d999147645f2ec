"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):  python oracle/gen_golden.py
The reference has no golden vectors or tests for this path, so its own
functions are executed here on seeded synthetic inputs and the outputs are committed:

  local_<model>_ns<k>.npz : params[B,P], y0[n], t[T], sol[B,T,n], flat[B,L]   models/*.py solve_ode
                            (stock LSODA = "O1"), sol_tight[B,T,n] (reference RHS through odeint
                            at rtol=atol=1e-12 = "O2"), score[B] (config/config.py score_fit vs
                            target), Y_<metric>[B] (sensitivity/analysis.py _compute_Y)
  steady.npz              : steady/init*.py initial_condition values
Also records scipy/numba/numpy versions (LSODA step sequences can differ at 1e-8 level
between SciPy versions; the reference pins scipy 1.15.2, this image has 1.18.x).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
CASES = [("distmod", 1), ("distmod", 3), ("distmod", 4), ("succmod", 1), ("succmod", 2),
         ("succmod", 5), ("randmod", 1), ("randmod", 3), ("randmod", 4), ("randmod", 6)]
METRICS = ("total_signal", "mean_activity", "variance", "dynamics", "l2_norm")


def load_compute_Y(metric):
    """sensitivity/analysis.py closes over Y_METRIC at jit time: one module per metric,
    with SALib / plotting (absent here, not on the path being pinned) stubbed out."""
    ref_shim.install_stubs(y_metric=metric)
    # numba's on-disk cache is keyed by source file, not by the global it closed over: give every
    # metric its own cache directory or the first compiled metric would be returned for all.
    import numba
    numba.config.CACHE_DIR = f"/tmp/phoskin_numba_cache_{metric}"
    const = sys.modules["config.constants"]
    const.NUM_TRAJECTORIES, const.PARAMETER_SPACE = 1000, 400
    const.TIME_POINTS_RNA = np.array([4.0, 8.0, 15.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
    const.PERTURBATIONS_VALUE, const.OUT_DIR = 0.5, "/tmp"
    for name, attrs in (("SALib", {}), ("SALib.sample", {"morris": None}), ("SALib.analyze", {}),
                        ("SALib.analyze.morris", {"analyze": None}), ("plotting", {}),
                        ("plotting.plotting", {"Plotter": None}),
                        ("config.helpers", {"get_number_of_params_rand": None, "get_param_names_rand": None}),
                        ("models", {"solve_ode": None})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        if name in ("SALib", "plotting", "models"):
            m.__path__ = []
        sys.modules.setdefault(name, m)
    return ref_shim.load("sensitivity/analysis.py", f"_pkref_sens_{metric}")._compute_Y


def load_score_fit():
    ref_shim.install_stubs()
    const = sys.modules["config.constants"]
    for k in ("INPUT_EXCEL_PROTEIN", "INPUT_EXCEL_PSITE", "INPUT_EXCEL_RNA"):
        setattr(const, k, "")
    const.DEV_TEST, const.TIME_POINTS, const.BOOTSTRAPS = False, T, 0
    for k in ("UB_mRNA_prod", "UB_mRNA_deg", "UB_Protein_prod", "UB_Protein_deg", "UB_Phospho_prod"):
        setattr(const, k, 20.0)
    return ref_shim.load("config/config.py", "_pkref_config_config").score_fit


def main():
    import numba
    import scipy
    from scipy.integrate import odeint
    os.makedirs(OUT, exist_ok=True)
    models = ref_shim.load_local_models()
    steady = ref_shim.load_steady()
    score_fit = load_score_fit()
    compY = {m: load_compute_Y(m) for m in METRICS}
    versions = np.array([f"scipy={scipy.__version__}", f"numba={numba.__version__}",
                         f"numpy={np.__version__}"])
    st = {}
    for name, ns in CASES:
        mod = models[name]
        y0 = np.array(steady[name].initial_condition(ns))
        st[f"{name}_ns{ns}"] = y0
        n = y0.size
        P = 4 + 2 * ns if name != "randmod" else 4 + ns + (1 << ns) - 1
        B = 12 if n > 20 else 24
        rng = np.random.default_rng(1000 + 17 * ns + len(name))
        params = rng.uniform(0.05, 3.0, (B, P))
        params[B // 2:] = rng.uniform(0.01, 20.0, (B - B // 2, P))   # stiff half (fit bounds)
        params[-1, 4:] *= (rng.random(P - 4) > 0.3)                  # knockout-like zeros
        y0s = np.tile(y0, (B, 1))
        y0s[1::3] = rng.uniform(0.05, 2.0, (len(y0s[1::3]), n))      # non-steady starts
        sol = np.empty((B, T.size, n))
        tight = np.empty_like(sol)
        flat = []
        for b in range(B):
            s, f = mod.solve_ode(params[b], y0s[b], ns, T)
            sol[b] = s
            flat.append(f)
            if name == "randmod":
                A, Bm, C, D, S, Dd = mod.unpack_params(params[b], ns)
                args = (A, Bm, C, D, ns, S, Dd) + tuple(mod._precompute_indices(ns))
                fun = mod.ode_system
            else:
                args = tuple(mod.unpack_params(params[b], ns))
                fun = mod.ode_core
            tight[b] = np.clip(odeint(fun, y0s[b], T, args=args, rtol=1e-12, atol=1e-12,
                                      mxstep=100000), 0, None)
        flat = np.array(flat)
        target = flat[0] * (1.0 + 0.05 * rng.standard_normal(flat.shape[1]))
        score = np.array([score_fit(params[b], target, flat[b]) for b in range(B)])
        Ys = {f"Y_{m}": np.array([compY[m](np.ascontiguousarray(sol[b]), ns) for b in range(B)])
              for m in METRICS}
        np.savez_compressed(os.path.join(OUT, f"local_{name}_ns{ns}.npz"), params=params, y0=y0s,
                            t=T, sol=sol, sol_tight=tight, flat=flat, target=target, score=score,
                            versions=versions, **Ys)
        print(name, ns, "B", B, "max|stock-tight|/max(|tight|,1e-12):",
              float((np.abs(sol - tight) / np.maximum(np.abs(tight), 1e-12)).max()))
    np.savez_compressed(os.path.join(OUT, "steady.npz"), versions=versions, **st)


if __name__ == "__main__":
    main()
