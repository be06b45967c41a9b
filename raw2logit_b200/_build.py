"""Builds the CUDA library in-tree: raw2logit_b200/libr2l_isp.so (sm_100a only, no torch headers involved).

Every csrc/*.cu is its own translation unit (forward / backward per raw element type, generic kernels, host ABI);
they are compiled in parallel to objects under csrc/_obj/ and linked with `nvcc -shared`.
"""
import glob
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB_PATH = os.path.join(HERE, "libr2l_isp.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  [os.path.join(ROOT, "include", "r2l_isp.h")])


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the raw2logit_b200 CUDA library cannot be built")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in sources() + headers())


# ---- torch operator shim (TORCH_LIBRARY + C++ autograd node): raw2logit_b200/libr2l_torch.so -------------------------
TORCH_SRC = os.path.join(HERE, "csrc_torch", "r2l_torch.cpp")
TORCH_LIB_PATH = os.path.join(HERE, "libr2l_torch.so")


def torch_shim_is_stale():
    if not os.path.exists(TORCH_LIB_PATH):
        return True
    built = os.path.getmtime(TORCH_LIB_PATH)
    deps = [TORCH_SRC, os.path.join(ROOT, "include", "r2l_isp.h")]
    return any(os.path.getmtime(p) > built for p in deps)


def build_torch_shim(force=False):
    """g++ on csrc_torch/r2l_torch.cpp against the installed torch headers; links libr2l_isp.so by $ORIGIN rpath."""
    if not force and not torch_shim_is_stale():
        return TORCH_LIB_PATH
    import torch
    from torch.utils import cpp_extension as ce
    cuda_home = os.path.dirname(os.path.dirname(nvcc_path()))
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-o", TORCH_LIB_PATH + ".tmp", TORCH_SRC,
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(cuda_home, "include")]
    for inc in ce.include_paths():
        cmd += ["-isystem", inc]
    for lib in ce.library_paths():
        cmd += ["-L", lib, f"-Wl,-rpath,{lib}"]
    cmd += ["-L", HERE, "-l:libr2l_isp.so", "-Wl,-rpath,$ORIGIN", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda",
            "-ltorch_cuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed on the torch shim:\n" + res.stdout + res.stderr)
    os.replace(TORCH_LIB_PATH + ".tmp", TORCH_LIB_PATH)
    return TORCH_LIB_PATH


def _obj_path(src):
    return os.path.join(OBJ, os.path.splitext(os.path.basename(src))[0] + ".o")


def _compile_one(src, verbose):
    obj = _obj_path(src)
    newest_dep = max(os.path.getmtime(p) for p in [src] + headers())
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest_dep:
        return src, ""
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
    return src, res.stderr


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libr2l_isp.so next to this file.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for o in glob.glob(os.path.join(OBJ, "*.o")):
            os.remove(o)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        logs = list(pool.map(lambda s: _compile_one(s, verbose), srcs))
    cmd = [nvcc_path()] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH + ".tmp"] + [_obj_path(s) for s in srcs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        for src, log in logs:
            print(f"== {os.path.basename(src)}\n{log}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
