"""raw2logit_b200 -- B200-native (sm_100a) differentiable ISP: the hot path of raw2logit's pipeline_torch.py.

Public surface: ``raw2logit_b200.pipeline_torch`` (drop-in nn.Modules), ``raw2logit_b200.ops`` (torch custom ops
``torch.ops.raw2logit_isp.*`` + autograd.Function), ``raw2logit_b200.synthetic`` (raw generators, camera presets).
The CUDA library is built in-tree by ``raw2logit_b200._build.build()``; without it every op raises.
"""
__version__ = "0.1.0"
