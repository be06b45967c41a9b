"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (no GPU needed).

The CUDA ISP cannot run here, so the processor stand-in is the CPU oracle wrapped as an nn.Module with the same
parameter names -- the checker standing in for the product only inside this test.  What is under test is
raw2logit_b200.parallel / raw2logit_b200.model: batch sharding, the single flat gradient bucket, averaging, the
freeze / adv_parameters rules of LitModel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from oracle import isp_oracle
from raw2logit_b200 import model as r2l_model
from raw2logit_b200 import parallel
from raw2logit_b200 import synthetic as syn


class OracleProcessor(nn.Module):
    """CPU stand-in with ParametrizedProcessing's parameter names (test only)."""

    def __init__(self, preset="drone"):
        super().__init__()
        st = isp_oracle.default_state(syn.CAMERA_PRESETS[preset])
        self.black_level = nn.Parameter(st["black_level"].clone())
        self.white_balance = nn.Parameter(st["white_balance"].clone())
        self.colour_correction = nn.Parameter(st["colour_correction"].clone())
        self.gamma_correct = nn.Parameter(st["gamma_correct"].clone())
        self.debayer = nn.Conv2d(3, 3, 3, bias=False)
        self.debayer.weight.data = st["debayer.weight"].clone()
        self.sharpening_filter = nn.Conv2d(1, 1, 3, bias=False)
        self.sharpening_filter.weight.data = st["sharpening_filter.weight"].clone()
        self.gaussian_blur = nn.Conv2d(1, 1, 5, bias=False)
        self.gaussian_blur.weight.data = st["gaussian_blur.weight"].clone()
        self.register_buffer("M_RGB_2_YUV", st["M_RGB_2_YUV"].clone())
        self.register_buffer("M_YUV_2_RGB", st["M_YUV_2_RGB"].clone())

    def forward(self, raw):
        state = {k: v for k, v in list(self.named_parameters()) + list(self.named_buffers())}
        out, _ = isp_oracle.forward(raw, state)
        return out


def _tiny_classifier():
    torch.manual_seed(3)
    return nn.Sequential(nn.Conv2d(3, 4, 3, stride=2), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(4, 3))


def _build(**kw):
    return r2l_model.LitModel(_tiny_classifier(), nn.CrossEntropyLoss(reduction="mean"), processor=OracleProcessor(), **kw)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _build()
        raw = syn.smooth_scene(4, 16, 16, "drone", seed=5)
        labels = torch.tensor([0, 1, 2, 1])
        parallel.data_parallel_step(model, (raw, labels), rank=rank, world=world)
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        torch.save(grads, os.path.join(out_dir, f"grads{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_step_equals_single_process_full_batch(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(os.path.join(tmp_path, "grads0.pt"))
    g1 = torch.load(os.path.join(tmp_path, "grads1.pt"))
    model = _build()
    raw = syn.smooth_scene(4, 16, 16, "drone", seed=5)
    labels = torch.tensor([0, 1, 2, 1])
    parallel.data_parallel_step(model, (raw, labels), rank=0, world=1)
    for n, p in model.named_parameters():
        assert torch.equal(g0[n], g1[n]), n                                       # every rank holds the same average
        assert torch.allclose(g0[n], p.grad, rtol=1e-4, atol=1e-7), n             # = the full-batch gradient
    assert sum(v.numel() for k, v in g0.items() if k.startswith("processor.")) == 132


def test_sharding_and_flat_bucket_layout():
    x = torch.arange(10)
    assert torch.equal(parallel.shard_batch(x, 1, 4), torch.tensor([1, 5, 9]))
    assert sum(len(parallel.shard_batch(x, r, 4)) for r in range(4)) == 10
    model = _build()
    raw = syn.smooth_scene(2, 16, 16, "drone", seed=1)
    model.update_step((raw, torch.tensor([0, 1]))).backward()
    flat, layout = parallel.flat_gradients(model.parameters())
    assert flat.numel() == sum(n for _, _, n in layout) == sum(p.numel() for p in model.parameters())
    assert parallel.allreduce_gradients(model.parameters(), world=1) == flat.numel()


def test_litmodel_freeze_and_adversarial_parameter_selection():
    m = _build(freeze_processor=True)
    assert all(not p.requires_grad for p in m.processor.parameters()) and not m.processor.training
    m.train(True)
    assert not m.processor.training and m.classifier.training                    # model.py:136-142
    m = _build(adv_training=True, adv_parameters="gamma")
    on = [n for n, p in m.processor.named_parameters() if p.requires_grad]
    assert on == ["gamma_correct"]
    m.train(True)
    assert not m.processor.training                                               # no BN updates in adversarial mode
    m = _build(freeze_classifier=True)
    assert all(not p.requires_grad for p in m.classifier.parameters())
    opt = _build().configure_optimizers()
    assert isinstance(opt, torch.optim.Adam)
