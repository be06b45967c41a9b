"""The data-parallel exchange fused into the backward kernel (r2l_isp_backward_dp, include/r2l_isp.h)."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exchange_abi_without_gpu():
    """Buffer size rule and argument validation of the fused exchange: no GPU needed, nothing is launched."""
    from raw2logit_b200 import _lib
    lib = _lib.load()
    assert lib.r2l_isp_exchange_bytes(0) == 0 and lib.r2l_isp_exchange_bytes(17) == 0
    for world in (1, 2, 4, 8, 16):
        assert lib.r2l_isp_exchange_bytes(world) == 2 * world * 136 * 8 + 16      # + the device-side epoch word
    null = ctypes.c_void_p(None)
    args = [null, _lib.F32, 65535.0, 2, 8, 8, None, null, null, null, null, null, null, null, null, 0]
    assert lib.r2l_isp_backward_dp(*args, None, null) == -3                       # R2L_ERR_NULL_POINTER: no descriptor
    for bad in (_lib.IspAllreduce(0, 0, None, 1, 1.0), _lib.IspAllreduce(17, 0, None, 1, 1.0),
                _lib.IspAllreduce(2, 2, None, 1, 1.0), _lib.IspAllreduce(2, 0, None, 0, 1.0)):
        assert lib.r2l_isp_backward_dp(*args, ctypes.byref(bad), null) == -7      # R2L_ERR_BAD_ARGUMENT
    assert lib.r2l_isp_backward_dp(*args, ctypes.byref(_lib.IspAllreduce(2, 0, None, 1, 1.0)), null) == -3   # no peers


@pytest.mark.gpu
def test_fused_exchange_matches_nccl_on_two_gpus():
    """scripts/dp_check.py under torchrun: fused exchange == r2l_isp_backward + NCCL all-reduce, ranks bit-identical,
    both epoch parities, float and uint16 raw.  Needs two GPUs (skipped on a single-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "scripts", "dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "dp_check ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_module_path_switch_is_explicit():
    """The module path only uses the fused exchange after parallel.enable_fused_gradient_exchange(); the switch is a
    plain process-wide setting that can be cleared again (no GPU needed to check the plumbing)."""
    from raw2logit_b200 import ops, parallel
    assert ops._exchange is None
    marker = object()
    ops.set_gradient_exchange(marker, average=False)
    assert ops._exchange is marker and ops._exchange_average is False
    parallel.disable_fused_gradient_exchange()
    assert ops._exchange is None
