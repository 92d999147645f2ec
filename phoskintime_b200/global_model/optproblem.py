"""Population-batched form of the reference's `GlobalODE_MOO` (global_model/optproblem.py:31-160).

The reference is a pymoo `ElementwiseProblem`: one `simulate_odeint` + `LOSS_FN` per decision vector.
Here a whole population goes to the GPU in ONE launch (`pk_global_solve_batch` with `theta_mode=1`):
softplus unpacking (global_model/params.py:106-132), the prior penalty (optproblem.py:105-114), the
integration, the three-modality loss and the normalised objectives (optproblem.py:146-160) are all
computed per system on the device; only F[B,3] comes back.  pymoo itself (optimiser policy) is out of
scope; `_evaluate(x, out)` keeps the reference's per-vector signature for drop-in use.
"""
import numpy as np

from ..engine import get_engine
from .network import PARAM_KEYS
from .simulate import simulate_batch

ODE_REL_TOL, ODE_ABS_TOL, ODE_MAX_STEPS = 2e-6, 2e-9, 200000      # library defaults (reference: config.toml:403-406)


def softplus(x):
    """global_model/utils.py:228-253"""
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def inv_softplus(y):
    """global_model/utils.py:245-253"""
    y = np.maximum(np.asarray(y, dtype=np.float64), 1e-12)
    return np.log(np.expm1(y))


def init_raw_params(defaults, bounds_config):
    """theta0, slices, xl, xu in the reference's order c_k|A_i|B_i|C_i|D_i|Dp_i|E_i|tf_scale
    (global_model/params.py:25-103).  `bounds_config[key] = (phys_min, phys_max)`."""
    vecs, slices, lo, hi, curr = [], {}, [], [], 0
    for k in PARAM_KEYS + ("tf_scale",):
        raw = inv_softplus(np.atleast_1d(np.asarray(defaults[k], dtype=np.float64)))
        vecs.append(raw)
        slices[k] = slice(curr, curr + raw.size)
        curr += raw.size
        pmin, pmax = bounds_config[k]
        lo += [float(inv_softplus(np.array([pmin]))[0])] * raw.size
        hi += [float(inv_softplus(np.array([pmax]))[0])] * raw.size
    return np.concatenate(vecs), slices, np.asarray(lo), np.asarray(hi)


def unpack_params(theta, slices):
    """global_model/params.py:106-132"""
    out = {k: softplus(theta[slices[k]]) for k in PARAM_KEYS}
    out["tf_scale"] = float(softplus(theta[slices["tf_scale"]])[0])
    return out


class GlobalODE_MOO:
    """Three objectives (protein, RNA, phospho) of a population of raw decision vectors.

    Same constructor data as the reference (optproblem.py:38-85): `sys`, `slices`, `loss_data`,
    `defaults`, `lambdas` (keys protein/rna/phospho/prior), `time_grid`, `fail_value`.  The decision
    vector layout must be the packed order of `init_raw_params` (checked)."""

    n_obj = 3

    def __init__(self, sys, slices, loss_data, defaults, lambdas, time_grid, xl=None, xu=None, fail_value=1e12,
                 loss_mode=0, engine=None):
        self.sys, self.slices, self.loss_data, self.defaults = sys, slices, loss_data, defaults
        self.lambdas, self.time_grid, self.fail_value = lambdas, np.asarray(time_grid, dtype=np.float64), float(fail_value)
        self.xl, self.xu, self.loss_mode = xl, xu, int(loss_mode)
        self.engine = engine
        expect = sys.param_slices()
        for k in PARAM_KEYS + ("tf_scale",):
            if (slices[k].start, slices[k].stop) != (expect[k].start, expect[k].stop):
                raise ValueError(f"slices['{k}'] does not match the packed parameter order c_k|A|B|C|D|Dp|E|tf_scale")
        self.n_var = sys.n_params
        # optproblem.py:83-85 (also computed by the library for the F output; kept for inspection)
        self.norm_p = 1.0 / max(1e-6, float(np.sum(loss_data["w_prot"])))
        self.norm_r = 1.0 / max(1e-6, float(np.sum(loss_data["w_rna"])))
        self.norm_ph = 1.0 / max(1e-6, float(np.sum(loss_data["w_pho"])))
        sys.defaults = {**{k: np.asarray(defaults[k], dtype=np.float64).copy() for k in PARAM_KEYS},
                        "tf_scale": float(defaults["tf_scale"])}

    def evaluate_batch(self, X, return_status=False):
        """X[B, n_var] raw thetas -> F[B,3]; failed or non-finite systems get `fail_value`
        (optproblem.py:125-137)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[None, :]
        lam = (self.lambdas["protein"], self.lambdas["rna"], self.lambdas["phospho"])
        res = simulate_batch(self.sys, X, self.time_grid, ("F",), rtol=ODE_REL_TOL, atol=ODE_ABS_TOL, mxstep=ODE_MAX_STEPS,
                             theta_mode=True, loss_data=self.loss_data, loss_mode=self.loss_mode, lambdas=lam,
                             lambda_prior=self.lambdas["prior"], engine=self.engine or get_engine())
        F = np.array(res["F"], dtype=np.float64)
        bad = (np.asarray(res["status"]) != 0) | ~np.isfinite(F).all(axis=1)
        F[bad] = self.fail_value
        return (F, res["status"]) if return_status else F

    def _evaluate(self, x, out, *args, **kwargs):
        """Reference signature (optproblem.py:87): one decision vector, result in out['F'].
        Like the reference it also writes the unpacked parameters through to `sys`."""
        self.sys.update(**unpack_params(np.asarray(x, dtype=np.float64), self.slices))
        out["F"] = self.evaluate_batch(x)[0]
