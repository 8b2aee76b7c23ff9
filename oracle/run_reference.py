"""Import and execute the UNMODIFIED reference (`/root/reference/kon/...`) over the eager
`tensorflow` shim in `oracle/_ref_shim/`.  TEST INFRASTRUCTURE ONLY.

Used by `tests/golden/make_ref_golden.py` (which writes the committed fixtures
`tests/golden/ref_*.npz`) and by the CPU test that regenerates those fixtures when the reference
tree is present.  `/root/reference` does not exist on the GPU box: nothing that runs there
imports this module.

What is replaced (and why it does not touch the hot path):
  * `tensorflow`                      -> `oracle/_ref_shim/tensorflow` (see its README.md)
  * `kon.model.feature_eng.feature_transform`, `kon.model.feature_eng.base_model`
        pandas/LightGBM/gensim feature engineering; every hot-path module imports them only to
        instantiate `feature_tool(path)` / `base_model(path)` objects it never uses (IL:14-15,30-31)
  * `kon.model.ctr_model.layer.behavior_layer.rnn_demo`
        a fork of TensorFlow's own recurrent_v2.py (AUGRU for DIEN), imported by BL:13
Everything else -- interactive_layer.py, core_layer.py, behavior_layer.py, models.py,
data_prepare.py -- is loaded from the reference tree as it lies.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KON_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref_shim")

_STUBBED = (
    "kon.model.feature_eng.feature_transform",
    "kon.model.feature_eng.base_model",
    "kon.model.ctr_model.layer.behavior_layer.rnn_demo",
)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "kon", "model", "ctr_model", "model", "models.py"))


class _Unused:
    """Stands in for `feature_tool` / `base_model` / `AUGRU`: constructible, nothing else."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        raise AttributeError("shim stub: %s.%s is outside the hot path" % (type(self).__name__, name))


def load():
    """Returns a namespace with the reference modules: `.tf` (the shim), `.IL`, `.CL`, `.BL`,
    `.MD`, `.DP`.  Idempotent."""
    if not available():
        raise FileNotFoundError("reference tree not found at %s" % REFERENCE_ROOT)
    real_tf = sys.modules.get("tensorflow")
    if real_tf is not None and not str(getattr(real_tf, "__version__", "")).endswith("-shim"):
        raise RuntimeError("a real tensorflow is already imported; the shim is only for images without it")
    for p in (REFERENCE_ROOT, SHIM_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in _STUBBED:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(feature_tool=type("feature_tool", (_Unused,), {}),
                              base_model=type("base_model", (_Unused,), {}),
                              AUGRU=type("AUGRU", (_Unused,), {}))
            sys.modules[name] = m
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(io.StringIO()):        # every reference module prints os.getcwd()
        ns.tf = importlib.import_module("tensorflow")
        ns.IL = importlib.import_module("kon.model.ctr_model.layer.interactive_layer.interactive_layer")
        ns.CL = importlib.import_module("kon.model.ctr_model.layer.core_layer.core_layer")
        ns.BL = importlib.import_module("kon.model.ctr_model.layer.behavior_layer.behavior_layer")
        ns.DP = importlib.import_module("kon.utils.data_prepare")
        ns.MD = importlib.import_module("kon.model.ctr_model.model.models")
    for mod in (ns.IL, ns.CL, ns.BL, ns.DP, ns.MD):
        assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), mod.__file__
    return ns
