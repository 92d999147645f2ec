"""CPU: the C-ABI library loads, exports every symbol the header declares, and the host-side
mirror of the reference interface behaves (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

import phoskintime_b200 as pk
from phoskintime_b200 import _lib, parallel, paramest, sensitivity
from phoskintime_b200.steady import initial_condition


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "phoskin_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(pk.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/phoskin_b200.h but not exported"
    assert set(syms) == set(_lib.SYMBOLS), "ctypes prototypes out of sync with the header"


def test_abi_version_and_struct_layout():
    lib = _lib.load()
    assert lib.pk_abi_version() == 1
    assert lib.pk_sizeof_local_job() == ctypes.sizeof(_lib.PkLocalJob)
    job = _lib.PkLocalJob()
    lib.pk_local_job_init(ctypes.byref(job))
    assert job.y_metric == -1 and job.n_groups == 1 and list(job.score_w) == [1.0] * 5


def test_nlls_job_defaults_and_layout():
    """pk_nlls_job_init: curve_fit's default tolerances (ftol = xtol = gtol = 1e-8), the struct mirrors the header."""
    import ctypes as C
    lib = _lib.load()
    assert lib.pk_sizeof_nlls_job() == C.sizeof(_lib.PkNllsJob)
    job = _lib.PkNllsJob()
    lib.pk_nlls_job_init(C.byref(job))
    assert (job.ftol, job.xtol, job.gtol) == (1e-8, 1e-8, 1e-8) and job.max_iter == 100 and job.n_groups == 1
    assert job.fd_rel == 1e-4 and job.rtol == 2e-7 and job.atol == 2e-10 and list(job.score_w) == [1.0] * 5
    # (the defaults written by the C side land in the right ctypes fields: offsets agree, not just the size)
    assert job.mu0 == 0.0 and job.lam == 0.0 and job.max_steps == 0 and job.method == 0 and job.log_params == 0


def test_local_dims_follow_reference_layout():
    assert pk.local_dims("distmod", 3, 14) == (5, 10, 65)      # L = 9 + 14 + 3*14
    assert pk.local_dims("distmod", 4, 14) == (6, 12, 79)
    assert pk.local_dims("succmod", 5, 14) == (7, 14, 93)
    assert pk.local_dims("randmod", 6, 14) == (65, 73, 107)
    assert pk.local_dims("randmod", 3, 4) == (9, 14, 4 + 3 * 4)  # T<=5: empty RNA block
    with pytest.raises(pk.PhoskinError):
        pk.local_dims("randmod", 9, 14)
    with pytest.raises(pk.PhoskinError):
        pk.local_dims("distmod", 0, 14)


def test_compute_call_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pk.PhoskinError):
        pk.Engine(0)


def test_product_does_not_import_oracle():
    """The product path must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "phoskintime_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # (`odeint_args` / `simulate_odeint` are reference API names and allowed; importing
                # SciPy's integrators or anything under oracle/ is not)
                assert "import local_models" not in text and "from oracle" not in text and \
                       "import oracle" not in text and "import global_models" not in text and \
                       "scipy.integrate" not in text and "import scipy" not in text and \
                       "from scipy" not in text, f
    code = "import sys; import phoskintime_b200, phoskintime_b200.models, phoskintime_b200.sensitivity, phoskintime_b200.global_model; " \
           "assert not any(m.startswith('scipy') or 'local_models' in m or 'global_models' in m for m in sys.modules), 'leak'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_models_plugin_selects_module():
    from phoskintime_b200 import models
    for name in ("distmod", "succmod", "randmod"):
        mod = models.set_model(name)
        assert models.ODE_MODEL == name and models.solve_ode is mod.solve_ode and mod.MODEL == name
    with pytest.raises(ImportError):
        models.set_model("testmod")
    models.set_model("randmod")


def test_initial_condition_closed_forms():
    assert np.allclose(initial_condition(3, "distmod"), [1, 0.4, 0.2, 0.2, 0.2])
    assert initial_condition(5, "succmod") == initial_condition(5, "distmod")   # initsucc.py:38-41 quirk
    r = initial_condition(3, "randmod")
    assert len(r) == 9 and abs(r[0] - 1) < 1e-15 and np.allclose(r[2:5], r[2]) and np.allclose(r[5:8], r[5])
    with pytest.raises(ValueError):
        initial_condition(2, "testmod")


def test_initial_condition_matches_reference_goldens():
    """phoskintime_b200.steady.initial_condition (the PRODUCT's closed form / linear solve) against the values the
    unmodified reference's SLSQP formulation converged to (tests/golden/steady.npz, written by oracle/gen_golden.py from
    steady/initdist.py:9-50, initsucc.py:9-55, initrand.py:10-77)."""
    from conftest import GOLDEN
    from phoskintime_b200.steady import initial_condition
    g = np.load(os.path.join(GOLDEN, "steady.npz"))
    checked = 0
    for key in g.files:
        if key == "versions":
            continue
        model, ns = key.split("_ns")
        got = np.asarray(initial_condition(int(ns), model))
        assert got.shape == g[key].shape and np.abs(got - g[key]).max() < 1e-12, key
        checked += 1
    assert checked >= 9


def test_morris_sample_and_problem_definitions():
    prob = sensitivity.define_sensitivity_problem_ds(3, np.linspace(0.5, 2, 10))
    assert prob["names"] == ["A", "B", "C", "D", "S1", "S2", "S3", "D1", "D2", "D3"]
    pr = sensitivity.define_sensitivity_problem_rand(2, np.ones(9))
    assert pr["names"] == ["A", "B", "C", "D", "S1", "S2", "D1", "D2", "D12"]
    assert sensitivity.compute_bound(0.0) == [0.0, 0.1] and sensitivity.compute_bound(2.0) == [1.0, 3.0]
    X = sensitivity.morris_sample(prob, 40, 400, seed=1)
    assert X.shape == (440, 10)
    b = np.asarray(prob["bounds"])
    assert (X >= b[:, 0] - 1e-12).all() and (X <= b[:, 1] + 1e-12).all()
    d = np.diff(X.reshape(40, 11, 10), axis=1)
    assert ((np.abs(d) > 0).sum(axis=2) == 1).all() and ((np.abs(d) > 0).sum(axis=1) == 1).all()
    assert np.array_equal(X, sensitivity.morris_sample(prob, 40, 400, seed=1))


def test_multistart_points_follow_normest():
    lb, ub = np.zeros(12), np.full(12, 20.0)
    pts = paramest.multistart_points(np.ones(12), lb, ub, n_starts=48, seed=42, gene="ABL1")
    assert pts.shape == (48, 12) and np.array_equal(pts[0], np.ones(12))
    assert (pts >= lb).all() and (pts <= ub).all()
    strat = pts[17:]                      # 1 base + 16 jitters, then 31 stratified rows
    for j in range(12):                   # one sample per stratum in every dimension
        assert sorted(np.floor(strat[:, j] / 20.0 * 31).astype(int)) == list(range(31))
    assert np.array_equal(pts, paramest.multistart_points(np.ones(12), lb, ub, 48, seed=42, gene="ABL1"))
    best = paramest.best_per_group(np.array([3.0, 1, 2, 5, 4, 0.5]), np.array([0, 0, 1, 1, 2, 2]), 3)
    assert list(best) == [1, 2, 5]


def test_knockout_combinations_and_application():
    """knockout/helper.py:5-62 — order and content of the 4*(ns+2) settings, parameter zeroing."""
    from phoskintime_b200 import knockout
    combos = knockout.generate_knockout_combinations(3)
    assert len(combos) == 4 * (3 + 2) and combos[0] == {"transcription": False, "translation": False, "phosphorylation": False}
    assert combos[1]["phosphorylation"] is True and combos[2]["phosphorylation"] == [0] and combos[5]["translation"] is True
    base = np.arange(1.0, 11.0)                                     # A,B,C,D,S1..3,D1..3
    assert np.array_equal(knockout.apply_knockout(base, combos[0], 3), base)
    ko = knockout.apply_knockout(base, {"transcription": True, "translation": True, "phosphorylation": [1, 7]}, 3)
    assert list(ko) == [0, 2, 0, 4, 5, 0, 7, 8, 9, 10] and base[0] == 1.0            # copy, out-of-range site ignored
    assert list(knockout.apply_knockout(base, {"phosphorylation": True}, 3)[4:7]) == [0, 0, 0]
    assert knockout.knockout_name(combos[0]) == "WT"
    assert knockout.knockout_name({"transcription": True, "translation": False, "phosphorylation": [2]}, ["S1", "T5", "Y9"]) \
        == "Transcription KO_PhosphoSite KO Y9"


def test_shard_bounds_cover_and_align():
    for total, world, align in ((11000, 8, 11), (1000, 3, 1), (256000, 8, 256), (7, 8, 1)):
        spans = [parallel.shard_bounds(total, world, r, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all((hi - lo) % align == 0 for lo, hi in spans)
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= align
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 0, 3)


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["PK_ROOT"])
import numpy as np, torch
from phoskintime_b200 import parallel
run = parallel.ShardedRun(engine=None, backend="gloo")
total, align = 55, 11
lo, hi = run.bounds(total, align)
full_ref = np.arange(total, dtype=np.float64) ** 2
local = full_ref[lo:hi].copy()                       # stands in for this rank's per-sample losses
got = run.allgather(local, total, align)
assert got.shape == (total,) and np.array_equal(got, full_ref), got
pair = np.stack([full_ref[lo:hi], -full_ref[lo:hi]], axis=1)
got2 = run.allgather(torch.from_numpy(pair), total, align)
assert got2.shape == (total, 2) and np.array_equal(got2[:, 1].numpy(), -full_ref)
m = run.max_over_ranks(float(run.rank + 1))
assert m == float(run.world)
run.barrier()
print("rank", run.rank, "ok")
'''


def test_sharded_gather_world2_gloo(tmp_path):
    """N>1 host logic (shard -> per-sample results -> all-gather with ragged shards) on CPU/gloo."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PK_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2, out.stdout          # both ranks finished (lines may interleave)


def test_sigma_builders_match_reference_goldens():
    """phoskintime_b200.models.weights (early_emphasis, full_weight, get_weight_options) against the unmodified
    reference's models/weights.py:10-76, 148-240 (tests/golden/weights.npz, oracle/gen_golden_weights.py): every one
    of the 17 sigma options, with and without USE_CUSTOM_WEIGHTS, 1 / 3 / 5 sites."""
    from conftest import GOLDEN
    from phoskintime_b200.models import weights as W
    g = np.load(os.path.join(GOLDEN, "weights.npz"))
    t = g["t"]
    for custom in (1, 0):
        for ns in (1, 3, 5):
            tag = f"ns{ns}_c{custom}"
            early = W.early_emphasis(g[f"{tag}_pr"], g[f"{tag}_p"], t, ns)
            assert np.allclose(early, g[f"{tag}_early"], rtol=1e-15, atol=0)
            opts = W.get_weight_options(g[f"{tag}_target"], t, ns, True, 4 + 2 * ns, early, g[f"{tag}_ms"],
                                        use_custom_weights=bool(custom))
            assert list(opts.keys()) == [str(k) for k in g[f"{tag}_keys"]]
            assert len(opts) == (17 if custom else 1)
            for k, v in opts.items():
                # (inverse_moving_avg divides by x - mean3(x): SciPy's running-sum filter and a direct 3-term sum round
                #  differently, amplified by the cancellation -> 1e-12 instead of 1e-14)
                assert v.shape == g[f"{tag}_opt_{k}"].shape and np.allclose(v, g[f"{tag}_opt_{k}"], rtol=1e-12, atol=0), (tag, k)
            assert np.array_equal(W.full_weight(g[f"{tag}_ms"], False, 4 + 2 * ns), g[f"{tag}_full_noreg"])
