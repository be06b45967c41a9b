"""Loads the UNMODIFIED reference ``processing/pipeline_torch.py`` under an alias -- TEST INFRASTRUCTURE.

Used only by ``oracle/make_golden.py`` (fixture generation, in the build container) and by the optional
live-reference checks in ``tests/`` (skipped when ``/root/reference`` is absent, e.g. on the GPU box).
Nothing in the product path, ``smoke()`` or ``bench.py`` reads the reference tree.

The reference file does not import as shipped: ``numpy.lib.function_base`` is gone in NumPy 2 and the
import chain pulls packages that are not installed.  Four stub modules are pre-seeded; the reference's classes then
run unmodified on CPU.  For an fp64 run set the default dtype to float64 BEFORE loading (module-level constants and
``torch.zeros`` in ``raw2rgb`` follow the default dtype) -- hence fp64 needs its own process.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("R2L_REF", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF_ROOT, "processing", "pipeline_torch.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
        for k, v in attrs.items():
            if not hasattr(mod, k):
                setattr(mod, k, v)
        return
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod


def load_reference(fp64=False):
    import torch
    if fp64:
        torch.set_default_dtype(torch.float64)
    _stub("numpy.lib.function_base", interp=None)        # pipeline_torch.py:2
    _stub("processing")
    _stub("processing.pipeline_numpy", processing=None)  # :8
    _stub("utils")
    _stub("utils.base", np2torch=None, torch2np=None)    # :9
    _stub("segmentation_models_pytorch")                 # :11
    cwd = os.getcwd()
    os.chdir(REF_ROOT)                                   # :5-6 chdir('..') unless README.md is in cwd
    try:
        spec = importlib.util.spec_from_file_location(
            "_ref_pipeline_torch", os.path.join(REF_ROOT, "processing", "pipeline_torch.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod
