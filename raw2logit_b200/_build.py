"""Builds the CUDA library in-tree: raw2logit_b200/libr2l_isp.so (sm_100a only, no torch headers involved)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libr2l_isp.so")
SOURCES = [os.path.join(CSRC, "isp_kernels.cu")]
HEADERS = [os.path.join(CSRC, "isp_core.cuh"), os.path.join(CSRC, "isp_fwd2.cuh"), os.path.join(CSRC, "isp_bwd2.cuh"), os.path.join(CSRC, "isp_config.h"),
           os.path.join(ROOT, "include", "r2l_isp.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the raw2logit_b200 CUDA library cannot be built")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libr2l_isp.so next to this file.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH + ".tmp"] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
