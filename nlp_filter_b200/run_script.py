"""Runs one of the reference's example scripts, unmodified, against this package:

    python -m nlp_filter_b200.run_script /path/to/reference/gnss_stationary.py

What the harness supplies: the reference package names (`nlp`, `utils`), a
no-op `matplotlib` when matplotlib is not installed, and the working directory the script
expects (rc-car.py uses ./data/... relative paths).  Returns the script's globals.
"""
import os
import runpy
import sys
import types


class _Anything(object):
    """Absorbs any attribute access / call (figure(), plot(), legend(), show(), ...)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


def install_matplotlib_stub():
    try:
        import matplotlib.pyplot  # noqa: F401
        return False
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    any_ = _Anything()

    def _attr(name):
        if name.startswith("__"):          # keep inspect / importlib introspection of sys.modules sane
            raise AttributeError(name)
        return any_

    plt.__getattr__ = _attr
    mpl.__getattr__ = _attr
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    return True


def run(path, quiet=False):
    import nlp_filter_b200
    nlp_filter_b200.install_reference_names()
    install_matplotlib_stub()
    path = os.path.abspath(path)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(path))
    old_path0 = sys.path[0]
    try:
        if quiet:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                return runpy.run_path(path, run_name="__main__")
        return runpy.run_path(path, run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.path[0] = old_path0


if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    run(sys.argv[1])
