"""`raw2logit_b200.model.LitModel` (the reference's caller of the path, model.py:64-83,136-142) driving the CUDA processor:
freeze_processor, adv_parameters substring selection (only that group's gradient is produced) and the eval-BN rule of
adversarial training, each against the CPU oracle on the same weights."""
import pytest
import torch

from oracle import isp_oracle
from raw2logit_b200 import model as r2l_model, synthetic as syn
from processing.pipeline_torch import ParametrizedProcessing

pytestmark = pytest.mark.gpu

CAM = syn.CAMERA_PRESETS["drone"]


def _clf():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1),
                               torch.nn.Flatten(), torch.nn.Linear(4, 5))


def _oracle_grads(raw, state, clf_cpu, y, bn=None, wanted=isp_oracle.PARAM_KEYS):
    st = isp_oracle.cast_state(state, torch.float32, requires_grad=True)
    out, _ = isp_oracle.forward(raw, st, bn=bn)
    loss = torch.nn.functional.cross_entropy(clf_cpu(out), y)
    grads = torch.autograd.grad(loss, [st[k] for k in wanted], allow_unused=True)
    return loss.item(), dict(zip(wanted, grads))


def test_litmodel_train_step_matches_oracle_and_freeze_rules():
    dev = torch.device("cuda:0")
    raw = syn.smooth_scene(4, 64, 96, "drone", seed=11)
    y = torch.tensor([0, 1, 2, 3])
    state = syn.perturbed_state(isp_oracle.default_state(CAM))

    # 1. everything trainable, BatchNorm in train mode (train.py:195-196)
    proc = ParametrizedProcessing(CAM, batch_norm_output=True)
    proc.load_state_dict(state, strict=False)
    clf = _clf()
    lit = r2l_model.LitModel(clf, torch.nn.CrossEntropyLoss(), processor=proc).to(dev).train()
    assert lit.processor.training and lit.processor.batch_norm.training
    loss = lit.update_step((raw.to(dev), y.to(dev)))
    loss.backward()
    bn = dict(training=True, running_mean=torch.zeros(3), running_var=torch.ones(3))
    want_loss, want = _oracle_grads(raw, state, _clf(), y, bn=bn)
    assert abs(loss.item() - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
    named = dict(lit.processor.named_parameters())
    for k in isp_oracle.PARAM_KEYS:
        scale = max(1.0, want[k].abs().max().item())
        assert (named[k].grad.cpu() - want[k]).abs().max().item() <= 1e-4 * scale, k
    assert torch.allclose(lit.processor.batch_norm.running_mean.cpu(), bn["running_mean"], atol=1e-6)

    # 2. freeze_processor: eval mode, no processor gradients, classifier still trains (model.py:64-68, 136-142)
    proc = ParametrizedProcessing(CAM, batch_norm_output=True)
    proc.load_state_dict(state, strict=False)
    lit = r2l_model.LitModel(_clf(), torch.nn.CrossEntropyLoss(), processor=proc, freeze_processor=True).to(dev).train()
    assert not lit.processor.training and lit.classifier.training
    lit.update_step((raw.to(dev), y.to(dev))).backward()
    assert all(p.grad is None for p in lit.processor.parameters())
    assert all(p.grad is not None for p in lit.classifier.parameters())
    assert int(lit.processor.batch_norm.num_batches_tracked) == 0

    # 3. adversarial training on one parameter group: only 'gamma' requires grad, BatchNorm stays in eval mode
    proc = ParametrizedProcessing(CAM, batch_norm_output=True)
    proc.load_state_dict(state, strict=False)
    proc.batch_norm.running_mean.copy_(torch.tensor([0.4, 0.45, 0.5]))
    proc.batch_norm.running_var.copy_(torch.tensor([0.04, 0.05, 0.06]))
    rm, rv = proc.batch_norm.running_mean.clone(), proc.batch_norm.running_var.clone()
    lit = r2l_model.LitModel(_clf(), torch.nn.CrossEntropyLoss(), processor=proc, adv_training=True,
                             adv_parameters='gamma').to(dev).train()
    assert not lit.processor.training
    needs = {n: p.requires_grad for n, p in lit.processor.named_parameters()}
    assert needs == {n: ('gamma' in n) for n in needs}
    loss = lit.update_step((raw.to(dev), y.to(dev)))
    loss.backward()
    for n, p in lit.processor.named_parameters():
        assert (p.grad is not None) == ('gamma' in n), n
    bn = dict(training=False, running_mean=rm, running_var=rv)
    want_loss, want = _oracle_grads(raw, state, _clf(), y, bn=bn, wanted=("gamma_correct",))
    assert abs(loss.item() - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
    got = dict(lit.processor.named_parameters())["gamma_correct"].grad.cpu()
    assert (got - want["gamma_correct"]).abs().max().item() <= 1e-4 * max(1.0, want["gamma_correct"].abs().max().item())
    assert torch.equal(lit.processor.batch_norm.running_mean.cpu(), rm)      # eval mode: statistics untouched
