"""CUDA-graph capture of the ISP training step through the module API.

``GraphedStep(module, example_raw, grad_out)`` runs ``out = module(raw); out.backward(grad_out)`` a few times on a
side stream, then captures the same two calls (C++ autograd node, forward kernel, backward kernel, gradient
accumulation) plus the concatenation of the parameter gradients into ONE ``torch.cuda.CUDAGraph``.  A replay costs a
single graph launch of host time (a few microseconds instead of ~100 us of dispatcher / autograd-engine work per step),
which is what keeps small batches GPU-bound.  Inputs are static: write the next batch into ``step.raw`` (and, when the
cotangent changes, ``step.grad_out``) -- e.g. with a non-blocking copy from pinned host memory -- then ``step.replay()``;
read ``step.out``, ``step.grad_raw`` (fp32 input with ``need_raw_grad``) and ``step.flat_grads`` (the gradients of
``step.params`` in ``named_parameters`` order, one contiguous vector) afterwards, on the same stream.

Everything the captured kernels touch lives in the graph's private memory pool; the module's parameters are read in
place, so optimiser updates between replays are seen.  BatchNorm running statistics and ``num_batches_tracked`` are
updated by the replay, as in eager mode (the forward kernel advances the counter itself).
"""
import torch


class GraphedStep:
    def __init__(self, module, example_raw, grad_out, need_raw_grad=False, warmup=3):
        if not example_raw.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA tensors")
        self.module = module
        self.raw = example_raw.detach().clone()
        if need_raw_grad:
            if not self.raw.is_floating_point():
                raise TypeError("an integer raw batch has no gradient")
            self.raw.requires_grad_(True)
        self.grad_out = grad_out.detach().clone()
        self.params = [p for p in module.parameters() if p.requires_grad]
        kept = [p.grad for p in self.params]

        def run():
            for p in self.params:
                p.grad = None
            self.raw.grad = None
            out = module(self.raw)
            out.backward(self.grad_out)
            return out

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream().wait_stream(side)
        for p in self.params:
            p.grad = None
        self.raw.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = run()
            self.flat_grads = torch.cat([p.grad.reshape(-1) for p in self.params]) if self.params else None
            self.grad_raw = self.raw.grad if need_raw_grad else None
        # the captured gradient tensors belong to the graph; hand the module its own gradients back
        self.param_grads = [p.grad for p in self.params]
        for p, g in zip(self.params, kept):
            p.grad = g

    def replay(self):
        self.graph.replay()
        return self.out

    def assign_grads(self):
        """Point ``p.grad`` of the module's parameters at the graph's (static) gradient tensors, for an optimiser."""
        for p, g in zip(self.params, self.param_grads):
            p.grad = g
