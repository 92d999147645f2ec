"""CPU: the objective / observable layer of the oracle (oracle/global_models.py: softplus unpacking, `objectives`,
`fc_tables`, `scalar_metric`) and of the host mirror (`init_raw_params`, `unpack_params`, `compute_bounds`) against
golden vectors produced by the UNMODIFIED reference (oracle/gen_golden_objectives.py ran the reference's own
`params.init_raw_params` / `unpack_params`, `GlobalODE_MOO._evaluate`, `simulate_and_measure`,
`_compute_scalar_metric` and `compute_bounds`, and captured the trajectories those calls integrated).

Every comparison here is on the REFERENCE's trajectories, so the bound is rounding only (1e-12 relative)."""
import glob
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import global_models as og  # noqa: E402
from phoskintime_b200.global_model import (init_raw_params, metric_time_indices, synthetic_system,  # noqa: E402
                                           unpack_params)
from phoskintime_b200.global_model.sensitivity import compute_bounds  # noqa: E402

FILES = sorted(glob.glob(os.path.join(GOLDEN, "globalobj_*.npz")))
IDS = [os.path.basename(f)[10:-4] for f in FILES]
KEYS = ("c_k", "A_i", "B_i", "C_i", "D_i", "Dp_i", "E_i", "tf_scale")
# [global_model.bounds] of the reference's config.toml:367-395 (what its init_raw_params read when the goldens were made)
BOUNDS_CONFIG = {"c_k": (1e-3, 4.0), "A_i": (1e-6, 10.0), "B_i": (1e-3, 1.0), "C_i": (1e-3, 2.0), "D_i": (0.1, 0.5),
                 "Dp_i": (0.05, 5.0), "E_i": (1e-4, 10.0), "tf_scale": (2.0, 10.0)}


def load_case(path):
    g = np.load(path)
    s = synthetic_system(seed=int(g["seed"]), N=int(g["N"]), K=int(g["K"]), max_sites=int(g["max_sites"]),
                         model=int(g["model"]))
    ld = {k[3:]: g[k] for k in g.files if k.startswith("ld_")}
    for k in ("prot_base_idx", "rna_base_idx", "pho_base_idx"):
        ld[k] = int(ld[k])
    defaults = {k: (g[f"def_{k}"] if k != "tf_scale" else float(g[f"def_{k}"])) for k in KEYS}
    slices = {k: slice(int(a), int(b)) for k, (a, b) in zip(KEYS, g["slices"])}
    return g, s, ld, defaults, slices


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300)))


def test_goldens_present():
    assert len(FILES) == 4          # kinetic models 0, 1, 2 (combinatorial), 4


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_raw_parameter_transform_matches_reference(path):
    """params.py:25-132: theta0 / bounds of init_raw_params and the softplus unpacking."""
    g, s, _, defaults, slices = load_case(path)
    theta0, sl, xl, xu = init_raw_params(defaults, BOUNDS_CONFIG)
    assert all((sl[k].start, sl[k].stop) == (slices[k].start, slices[k].stop) for k in KEYS)
    assert np.allclose(theta0, g["theta0"], rtol=1e-14, atol=0) and np.allclose(xl, g["xl"], rtol=1e-14, atol=0)
    assert np.allclose(xu, g["xu"], rtol=1e-14, atol=0)
    for b in range(g["theta"].shape[0]):
        p = unpack_params(g["theta"][b], slices)
        mine = np.concatenate([np.ravel(p[k]) for k in KEYS[:-1]] + [[p["tf_scale"]]])
        assert _rel(mine, g["phys"][b]) < 1e-14
        assert _rel(og.softplus(g["theta"][b]), g["phys"][b]) < 1e-14          # the oracle's own softplus


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_objectives_match_reference_evaluate(path):
    """GlobalODE_MOO._evaluate (optproblem.py:87-160): F[3] from the reference's own trajectory."""
    g, s, ld, defaults, slices = load_case(path)
    net = s.as_dict()
    lam = g["lambdas"]
    for b in range(g["theta"].shape[0]):
        p = og.unpack_params(g["phys"][b], net)
        losses = og.loss(int(g["model"]), g["Y_obj"][b], ld, 0)
        F = og.objectives(losses, ld, p, defaults, tuple(lam[:3]), float(lam[3]))
        assert _rel(F, g["F"][b]) < 1e-12, (b, F, g["F"][b])


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_fold_change_tables_and_metrics_match_reference(path):
    """simulate_and_measure (simulate.py:83-202) and _compute_scalar_metric (sensitivity.py:106-140)."""
    g, s, _, _, _ = load_case(path)
    net = s.as_dict()
    mt = metric_time_indices(g["t_grid"], g["t_prot"], g["t_rna"], g["t_prot"])
    # row order of the reference's DataFrames: protein-major, sites in block order, times ascending
    assert np.array_equal(g["fc_prot_time"][:len(g["t_prot"])], g["t_prot"]) and np.array_equal(g["fc_rna_time"][:len(g["t_rna"])], g["t_rna"])
    assert g["fc_prot_protein"][0] == "P000" and g["fc_prot_protein"][len(g["t_prot"])] == "P001"
    for b in range(g["theta"].shape[0]):
        P, R, PH = og.fc_tables(g["Y_meas"][b], net, mt)
        assert _rel(P.reshape(-1), g["fc_prot"][b]) < 1e-12
        assert _rel(R.reshape(-1), g["fc_rna"][b]) < 1e-12
        assert _rel(PH.reshape(-1), g["fc_pho"][b]) < 1e-12
        for m, name in enumerate(g["metric_names"]):
            assert abs(og.scalar_metric(g["Y_meas"][b], net, mt, str(name)) - g["metrics"][b, m]) <= 1e-12 * abs(g["metrics"][b, m])


@pytest.mark.parametrize("path", FILES[:1], ids=IDS[:1])
def test_compute_bounds_matches_reference(path):
    """sensitivity.py:39-79 with the reference's SENSITIVITY_PERTURBATION = 0.05 (config.toml:351)."""
    g, s, _, defaults, _ = load_case(path)
    prob = compute_bounds(defaults, perturbation=0.05)
    assert prob["names"] == [str(n) for n in g["bound_names"]]
    assert np.allclose(np.array(prob["bounds"]), g["bounds"], rtol=1e-15, atol=0)
