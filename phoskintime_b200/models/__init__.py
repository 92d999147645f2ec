"""Plugin point mirroring the reference's `models/__init__.py:6-12`: the module named by
ODE_MODEL provides `solve_ode`.  The reference reads ODE_MODEL from config.toml `[ode].model`
(config/constants.py:27, default "randmod"); here it comes from the environment variable
PHOSKIN_ODE_MODEL (same default) or `set_model()`.
"""
import importlib
import os

ODE_MODEL = os.environ.get("PHOSKIN_ODE_MODEL", "randmod")
_VALID = ("distmod", "succmod", "randmod")


def _load(name):
    if name not in _VALID:
        raise ImportError(f"Cannot import model module 'models.{name}'")
    return importlib.import_module(f"{__name__}.{name}")


model_module = _load(ODE_MODEL)
solve_ode = model_module.solve_ode
solve_ode_batch = model_module.solve_ode_batch


def set_model(name):
    """Re-select the plugin at run time (the reference needs a config edit + re-import)."""
    global ODE_MODEL, model_module, solve_ode, solve_ode_batch
    model_module = _load(name)
    ODE_MODEL = name
    solve_ode = model_module.solve_ode
    solve_ode_batch = model_module.solve_ode_batch
    return model_module
