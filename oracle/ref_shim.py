"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference modules.

Imports the reference's hot-path modules from /root/reference through config
stubs: `config.constants` pulls in matplotlib and mkdirs in
the read-only tree, so the few constants the hot path reads are stubbed.  Used
solely by `oracle/gen_golden.py` (in the build container, where /root/reference
exists) to produce the committed fixtures in tests/golden/.  Nothing on the
product path, and nothing that runs on the GPU box, imports this file.
"""
import importlib.util
import logging
import os
import sys
import types

REF_ROOT = os.environ.get("PHOSKIN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs(ode_model="distmod", y_metric="total_signal"):
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/phoskin_numba_cache")
    cfg = _stub("config")
    cfg.__path__ = []
    _stub("config.constants",
          NORMALIZE_MODEL_OUTPUT=False, ODE_MODEL=ode_model, Y_METRIC=y_metric,
          ALPHA_WEIGHT=1.0, BETA_WEIGHT=1.0, GAMMA_WEIGHT=1.0, DELTA_WEIGHT=1.0,
          MU_WEIGHT=1.0)
    lg = lambda *a, **k: logging.getLogger("phoskin_ref")
    _stub("config.logconf", setup_logger=lg)


def load(relpath, modname=None):
    """Load /root/reference/<relpath> as module `modname` (registered before exec,
    otherwise numba's cache=True cannot locate it)."""
    modname = modname or ("_pkref_" + relpath.replace("/", "_").removesuffix(".py"))
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_local_models():
    install_stubs()
    return {name: load(f"models/{name}.py") for name in ("distmod", "succmod", "randmod")}


def load_steady():
    install_stubs()
    return {"distmod": load("steady/initdist.py"),
            "succmod": load("steady/initsucc.py"),
            "randmod": load("steady/initrand.py")}
