"""Python-3 import shim for the UNMODIFIED reference at /root/reference (build container only).

Used solely by tests/golden/make_golden.py to generate the committed fixtures; the reference
tree does not exist on the GPU box and nothing at test/bench run time imports this file.
Shims: xrange/reduce builtins, numpy.float, stub `ujson` and
`schematics` modules, a py3 restatement of Parallel's sequential dispatcher (the reference's
uses dict.iteritems), and a replayable stand-in for scipy.stats.gamma inside
sbo/lib/expectations.py so that W samples are explicit.
"""
import builtins
import functools
import json
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def install():
    builtins.xrange = range
    builtins.reduce = functools.reduce
    if not hasattr(np, "float"):
        np.float = float
    uj = types.ModuleType("ujson")
    uj.load, uj.dump, uj.loads, uj.dumps = json.load, json.dump, json.loads, json.dumps
    sys.modules.setdefault("ujson", uj)

    class _Any(object):
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, n):
            return _Any()

    class Model(object):
        def __init__(self, raw=None, **k):
            for a, b in (raw or {}).items():
                setattr(self, a, b)

        def validate(self):
            pass

    class ModelValidationError(Exception):
        pass

    for name in ["schematics", "schematics.models", "schematics.types",
                 "schematics.types.compound", "schematics.exceptions"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["schematics.models"].Model = Model
    for t in ["IntType", "StringType", "BooleanType", "FloatType", "BaseType"]:
        setattr(sys.modules["schematics.types"], t, _Any)
    for t in ["ModelType", "ListType", "DictType"]:
        setattr(sys.modules["schematics.types.compound"], t, _Any)
    sys.modules["schematics.exceptions"].ModelValidationError = ModelValidationError
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    from stratified_bayesian_optimization.lib.parallel import Parallel

    def run_sequential(function, arguments, *args, **kwargs):
        return {key: function(*((argument,) + args), **kwargs)
                for key, argument in arguments.items()}

    def run_parallel(function, arguments, all_success=False, signal=None, parallel=True,
                     threads=0, *args, **kwargs):
        return run_sequential(function, arguments, *args, **kwargs)

    Parallel.run_function_different_arguments_sequentially = staticmethod(run_sequential)
    Parallel.run_function_different_arguments_parallel = staticmethod(run_parallel)


class ReplayGamma(object):
    """Stand-in for scipy.stats.gamma: rvs() replays pre-drawn arrays cyclically."""

    def __init__(self, sets):
        self.sets = [np.asarray(s, dtype=float) for s in sets]
        self.calls = 0

    def rvs(self, a, scale=1.0, size=None):
        s = self.sets[self.calls % len(self.sets)]
        self.calls += 1
        assert tuple(size) == s.shape, (size, s.shape)
        return s.copy()


def patch_gamma(sets):
    import stratified_bayesian_optimization.lib.expectations as ex
    rg = ReplayGamma(sets)
    ex.gamma = rg
    return rg
