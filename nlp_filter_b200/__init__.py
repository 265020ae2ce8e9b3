"""nlp_filter_b200 -- B200-native estimator hot path of kingdwd/nlp-filter.

Layout
  csrc/ + libnlpfilter.so   CUDA kernels (sm_100a) behind the C ABI of include/nlpfilter.h
  _capi.py, engine.py       ctypes binding and batched device entry points
  desc.py, shapes.py        flat problem-class descriptor and builders of the example shapes
  nlp/, utils/              host-side mirror of the reference's Python API (same names)

The reference's scripts import top-level packages called `nlp` and `utils`.
`install_reference_names()` publishes this package's sub-packages under those names so the
scripts run unmodified (see run_script.py).
"""
import importlib
import sys

__version__ = "0.1.0"

_NLP_MODULES = ("collocation", "constraints", "cost_functions", "dynamics", "measurements", "nlp")
_UTILS_MODULES = ("utils", "gnss", "ekf", "leastsquares", "data", "simulate", "vehicle_sim")


def install_reference_names(force=False):
    """Make `import nlp.nlp`, `import utils.ekf`, ... resolve to this package's modules."""
    for top, names in (("nlp", _NLP_MODULES), ("utils", _UTILS_MODULES)):
        if top in sys.modules and not force and getattr(sys.modules[top], "__nlpf__", False):
            continue
        pkg = importlib.import_module("nlp_filter_b200." + top)
        pkg.__nlpf__ = True
        sys.modules[top] = pkg
        for n in names:
            mod = importlib.import_module("nlp_filter_b200.%s.%s" % (top, n))
            sys.modules["%s.%s" % (top, n)] = mod
            setattr(pkg, n, mod)
    # user-defined model functions written against the reference do `from casadi import vertcat, sin, ...`
    # (nlp/dynamics.py:2): when CasADi itself is not installed, publish the stand-in names (NumPy on numbers, traced
    # expressions under the code generator)
    if "casadi" not in sys.modules:
        try:
            importlib.import_module("casadi")
        except Exception:
            sys.modules["casadi"] = importlib.import_module("nlp_filter_b200.nlp._sym")
