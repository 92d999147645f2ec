"""In-silico knockout sweeps as ONE batched solve — mirrors `knockout/helper.py:5-62` and the loop in
`paramest/core.py:144-187`.

The reference builds 4*(num_psites+2) knockout settings per protein (transcription x translation x
{none, all sites, each single site}), zeroes the corresponding parameters and calls `solve_ode` once per
setting.  Here the modified parameter vectors form a batch: one kernel launch returns every knockout
trajectory (and the wild type)."""
import itertools

import numpy as np

from . import models


def apply_knockout(base_params, knockout_targets, num_psites):
    """knockout/helper.py:5-36 — A (index 0) = transcription, C (index 2) = translation, S_i (4..4+ns) = sites."""
    params = np.array(base_params, dtype=np.float64, copy=True)
    if knockout_targets.get("transcription", False):
        params[0] = 0.0
    if knockout_targets.get("translation", False):
        params[2] = 0.0
    if "phosphorylation" in knockout_targets:
        k = knockout_targets["phosphorylation"]
        start, end = 4, 4 + num_psites
        if isinstance(k, bool) and k:
            params[start:end] = 0.0
        elif isinstance(k, (list, tuple)):
            for idx in k:
                if 0 <= idx < num_psites:
                    params[start + idx] = 0.0
    return params


def generate_knockout_combinations(num_psites):
    """knockout/helper.py:39-62 (same order)."""
    phospho = [False, True] + [[i] for i in range(num_psites)]
    return [{"transcription": a, "translation": b, "phosphorylation": c}
            for a, b, c in itertools.product([False, True], [False, True], phospho)]


def knockout_name(setting, psite_labels=None):
    """paramest/core.py:154-166"""
    parts = []
    if setting["transcription"]:
        parts.append("Transcription KO")
    if setting["translation"]:
        parts.append("Translation KO")
    ph = setting["phosphorylation"]
    if ph is True:
        parts.append("Phospho KO")
    elif isinstance(ph, list) and ph:
        parts.append("PhosphoSite KO " + ",".join(str(psite_labels[p]) if psite_labels is not None else str(p) for p in ph))
    return "_".join(parts) if parts else "WT"


def simulate_knockouts(final_params, init_cond, num_psites, time_points, model=None, psite_labels=None, **kw):
    """{name: {knockout_setting, sol_ko, p_fit_ko}} for every combination, from ONE batched launch
    (`models.<model>.solve_ode_batch`); the first combination is the wild type."""
    mod = models.set_model(model) if model is not None else models.model_module
    combos = generate_knockout_combinations(num_psites)
    P = np.stack([apply_knockout(final_params, c, num_psites) for c in combos])
    res = mod.solve_ode_batch(P, init_cond, num_psites, time_points, want=("sol", "flat"), **kw)
    out = {}
    for b, c in enumerate(combos):
        out[knockout_name(c, psite_labels)] = {"knockout_setting": c, "sol_ko": res["sol"][b], "p_fit_ko": res["flat"][b],
                                               "status": int(res["status"][b])}
    return out
