"""Drop-in switch for the UNMODIFIED reference scripts (evalOC.py, timeDeployment/timeOC.py, compare*.py).

Put this directory first on PYTHONPATH:

    PYTHONPATH=/path/to/repo/neuraloc_b200/dropin:/path/to/repo:/path/to/NeuralOC python evalOC.py --nt 50 ...

Python imports `sitecustomize` at start-up; it installs an import hook that lets the reference's own
`src/OCflow.py` load and then rebinds the hot-path names (`OCflow`, `stepRK4`, `stepRK1`, `ocOdefun`) to the
B200 implementation, so `from src.OCflow import OCflow` in the drivers and in src/plotter.py picks it up.
Phi, the problem classes and initProb stay the reference's own objects (the rollout duck-types them).
"""
import importlib.abc
import importlib.machinery
import os
import sys

if os.environ.get("NOC_DROPIN", "1") != "0":

    class _Loader(importlib.abc.Loader):
        def __init__(self, inner):
            self.inner = inner

        def create_module(self, spec):
            return self.inner.create_module(spec)

        def exec_module(self, module):
            self.inner.exec_module(module)
            import neuraloc_b200 as nb
            for name in ("OCflow", "stepRK4", "stepRK1", "ocOdefun"):
                setattr(module, "_reference_" + name, getattr(module, name, None))
                setattr(module, name, getattr(nb, name))

    class _Finder(importlib.abc.MetaPathFinder):
        def find_spec(self, fullname, path, target=None):
            if fullname != "src.OCflow":
                return None
            spec = importlib.machinery.PathFinder.find_spec(fullname, path)
            if spec is not None and spec.loader is not None:
                spec.loader = _Loader(spec.loader)
            return spec

    sys.meta_path.insert(0, _Finder())
