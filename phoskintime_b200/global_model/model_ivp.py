"""`fun(t, y)` factories for solve_ivp-style host solvers — device mirror of the reference's
`global_model/model_ivp.py` (:49-277).

The reference closes the block-kinetics kernels (`distributive_rhs`, `sequential_rhs`, `saturating_rhs`,
`combinatorial_rhs`, global_model/models.py:71-432) over the parameter arrays and returns `fun(t, y) -> dy`; the caller
supplies the transcription-factor inputs (array or callable) and the phosphorylation rates `S_all` itself.  Here the same
keyword signatures build a small device topology once and every call of `fun` is one `pk_global_rhs_batch` launch in its
direct mode (TF inputs and S_all as given, ONE squash of the TF input as in models.py:52).  `fun.batch(t, Y)` evaluates
many states at once and `fun.jac(t, y)` returns the analytic block Jacobian — the batched surface the per-vector
closures were missing.  `make_solve_ivp_fun(sys)` is the form the integrator itself uses: the full `rhs_odeint`
(jacspeedup.py:175-375: kinase buckets, W and TF matrices, both squashes) with its analytic Jacobian.
"""
import numpy as np

from ..engine import get_engine
from .network import GlobalSystem


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


def _wrap_tf_input(tf_input):
    """model_ivp.py:31-46: None -> zeros, ndarray -> constant, callable(t) / callable(t, y) -> as is."""
    if tf_input is None:
        return None
    if callable(tf_input):
        return tf_input
    const = _c(tf_input)
    return lambda t, y=None: const


def _call_tf(tfp, t, y, N):
    if tfp is None:
        return np.zeros(N)
    try:
        return _c(tfp(t, y))
    except TypeError:
        return _c(tfp(t))


def _factory(model, *, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, tf_input, S_provider, offset_y, offset_s, n_sites, engine):
    A_i, B_i, C_i, D_i, Dp_i, E_i = (_c(v) for v in (A_i, B_i, C_i, D_i, Dp_i, E_i))
    n_sites = _c(n_sites, np.int32)
    N, S = A_i.shape[0], int(n_sites.sum())
    eng = engine or get_engine()
    # a shell network: the block layout is all the direct mode reads (no kinases, no TF edges)
    shell = GlobalSystem.shell(model, n_sites, offset_y=_c(offset_y, np.int32), offset_s=_c(offset_s, np.int32))
    topo = eng.global_upload(shell)
    params = np.concatenate([np.ones(shell.K), A_i, B_i, C_i, D_i, Dp_i, E_i, [float(tf_scale)]])
    tfp = _wrap_tf_input(tf_input)

    def batch(t, Y, want_jac=False):
        Y = np.atleast_2d(_c(Y))
        B = Y.shape[0]
        tf = np.stack([_call_tf(tfp, t, Y[b], N) for b in range(B)]) if callable(tf_input) else \
            np.broadcast_to(_call_tf(tfp, t, Y[0], N), (B, N))
        return eng.global_rhs_batch(topo, params, Y, tf_inputs=np.ascontiguousarray(tf),
                                    S_all=np.broadcast_to(S_provider(), (B, S)).copy(), want_jac=want_jac)

    def fun(t, y):
        return batch(t, np.asarray(y, dtype=np.float64)[None, :])[0]

    fun.batch = batch
    fun.jac = lambda t, y: batch(t, np.asarray(y, dtype=np.float64)[None, :], want_jac=True)[1][0]
    fun.topology = topo
    return fun


def make_solve_ivp_fun_distributive(*, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, tf_input, S_all, offset_y, offset_s, n_sites,
                                    engine=None):
    """model_ivp.py:108-156."""
    S = _c(S_all)
    return _factory(0, A_i=A_i, B_i=B_i, C_i=C_i, D_i=D_i, Dp_i=Dp_i, E_i=E_i, tf_scale=tf_scale, tf_input=tf_input,
                    S_provider=lambda: S, offset_y=offset_y, offset_s=offset_s, n_sites=n_sites, engine=engine)


def make_solve_ivp_fun_sequential(*, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, tf_input, S_all, offset_y, offset_s, n_sites,
                                  engine=None):
    """model_ivp.py:159-207."""
    S = _c(S_all)
    return _factory(1, A_i=A_i, B_i=B_i, C_i=C_i, D_i=D_i, Dp_i=Dp_i, E_i=E_i, tf_scale=tf_scale, tf_input=tf_input,
                    S_provider=lambda: S, offset_y=offset_y, offset_s=offset_s, n_sites=n_sites, engine=engine)


def make_solve_ivp_fun_saturating(*, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, tf_input, S_all, offset_y, offset_s, n_sites,
                                  engine=None):
    """model_ivp.py:49-105."""
    S = _c(S_all)
    return _factory(4, A_i=A_i, B_i=B_i, C_i=C_i, D_i=D_i, Dp_i=Dp_i, E_i=E_i, tf_scale=tf_scale, tf_input=tf_input,
                    S_provider=lambda: S, offset_y=offset_y, offset_s=offset_s, n_sites=n_sites, engine=engine)


def make_solve_ivp_fun_combinatorial(*, A_i, B_i, C_i, D_i, Dp_i, E_i, tf_scale, tf_input, S_cache, jb, offset_y, offset_s,
                                     n_sites, n_states=None, trans_from=None, trans_to=None, trans_site=None, trans_off=None,
                                     trans_n=None, engine=None):
    """model_ivp.py:210-277.  The transition tables of the reference (`build_random_transitions`, models.py:435-485)
    enumerate the hypercube of phosphorylation patterns; the device kernel walks the same hypercube by bit arithmetic,
    so `trans_*` / `n_states` are accepted for signature compatibility and not consulted."""
    S = _c(S_cache)[:, int(jb)].copy()
    return _factory(2, A_i=A_i, B_i=B_i, C_i=C_i, D_i=D_i, Dp_i=Dp_i, E_i=E_i, tf_scale=tf_scale, tf_input=tf_input,
                    S_provider=lambda: S, offset_y=offset_y, offset_s=offset_s, n_sites=n_sites, engine=engine)


def make_solve_ivp_fun(sys, engine=None):
    """`fun(t, y)` of the CURRENT parameters of a `GlobalSystem` with the semantics of `rhs_odeint`
    (jacspeedup.py:175-375) — what `simulate_odeint` integrates — plus `fun.jac` (analytic Jacobian, the quantity
    `fd_jacobian_odeint` approximates, jacspeedup.py:397-588) and `fun.batch(t, Y, params=None, want_jac=False)`."""
    from .simulate import _topology
    eng = engine or get_engine()
    topo = _topology(sys, eng)

    def batch(t, Y, params=None, want_jac=False):
        return eng.global_rhs_batch(topo, sys.pack_params() if params is None else params, Y, t, want_jac=want_jac)

    def fun(t, y):
        return batch(t, np.asarray(y, dtype=np.float64)[None, :])[0]

    fun.batch = batch
    fun.jac = lambda t, y: batch(t, np.asarray(y, dtype=np.float64)[None, :], want_jac=True)[1][0]
    return fun
