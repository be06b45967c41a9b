"""Loads the UNMODIFIED reference ``processing/pipeline_torch.py`` under an alias -- TEST INFRASTRUCTURE.

Used by ``oracle/make_golden.py`` (fixture generation, in the build container), by the optional live-reference
checks in ``tests/`` (skipped when no copy is present) and by ``bench.py``'s reference arm / ``cpu_baseline`` leg and
``scripts/ref_on_gpu.py`` (timing only).  Nothing in the product path or ``smoke()`` reads it.

Where the file comes from: ``/root/reference`` in the build container; on the GPU box (which has no ``/root/reference``)
the single file ``processing/pipeline_torch.py`` staged by ``stage()`` -- called from ``__graft_entry__.build()`` --
under the git-ignored ``baseline/_ref/``, which travels with the snapshot.  The reference has no
packaging metadata, so ``pip install /root/reference`` has nothing to install; staging the file is the install.

The reference file does not import as shipped: ``numpy.lib.function_base`` is gone in NumPy 2 and the
import chain pulls packages that are not installed.  Four stub modules are pre-seeded; the reference's classes then
run unmodified on CPU.  For an fp64 run set the default dtype to float64 BEFORE loading (module-level constants and
``torch.zeros`` in ``raw2rgb`` follow the default dtype) -- hence fp64 needs its own process.
"""
import importlib.util
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _pick_root():
    for cand in (os.environ.get("R2L_REF"), "/root/reference", STAGED_ROOT):
        if cand and os.path.exists(os.path.join(cand, "processing", "pipeline_torch.py")):
            return cand
    return os.environ.get("R2L_REF", "/root/reference")


REF_ROOT = _pick_root()


def available():
    return os.path.exists(os.path.join(REF_ROOT, "processing", "pipeline_torch.py"))


def stage(src_root="/root/reference"):
    """Copies the unmodified reference file (and a README.md marker its ``chdir`` logic looks for, :5-6) into the
    git-ignored ``baseline/_ref/`` so that timing runs on the GPU box can execute it.  No-op without the source."""
    import shutil
    src = os.path.join(src_root, "processing", "pipeline_torch.py")
    if not os.path.exists(src):
        return None
    os.makedirs(os.path.join(STAGED_ROOT, "processing"), exist_ok=True)
    shutil.copyfile(src, os.path.join(STAGED_ROOT, "processing", "pipeline_torch.py"))
    with open(os.path.join(STAGED_ROOT, "README.md"), "w") as f:
        f.write("staged copy of the reference's processing/pipeline_torch.py (unmodified; git-ignored; timing only)\n")
    return STAGED_ROOT


def _stub(name, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
        for k, v in attrs.items():
            if not hasattr(mod, k):
                setattr(mod, k, v)
        return
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    # a stub that shadows one of this repo's own drop-in packages (processing/, utils/) stays a package: later
    # `import utils.ssim` / `import processing.pipeline_torch` in the same process must still find the real files
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), *name.split("."))
    if os.path.isdir(here):
        mod.__path__ = [here]
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
