"""CUDA-graph capture of the ISP training step through the module API.

``GraphedStep(module, example_raw, grad_out)`` runs ``out = module(raw); torch.autograd.grad(out, params, grad_out)`` a
few times on a side stream, then captures the same two calls (C++ autograd node, forward kernel, backward kernel) plus
the concatenation of the parameter gradients into ONE ``torch.cuda.CUDAGraph``.  A replay costs a
single graph launch of host time (a few microseconds instead of ~100 us of dispatcher / autograd-engine work per step),
which is what keeps small batches GPU-bound.  Inputs are static: write the next batch into ``step.raw`` (and, when the
cotangent changes, ``step.grad_out``) -- e.g. with a non-blocking copy from pinned host memory -- then ``step.replay()``;
read ``step.out``, ``step.grad_raw`` (fp32 input with ``need_raw_grad``) and ``step.flat_grads`` (the gradients of
``step.params`` in ``named_parameters`` order, one contiguous vector) afterwards, on the same stream.  The step does not
touch ``p.grad`` of the module (``assign_grads()`` points them at the graph's tensors for an optimiser).

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
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        self.params = [p for _, p in named]
        # The captured step differentiates ALIASES of the parameters (new leaves on the same storage, so optimiser updates
        # are seen) with torch.autograd.grad: no AccumulateGrad node takes part.  A parameter's gradient accumulator lives
        # on the stream where an earlier eager step created it (typically the legacy default stream) for as long as any
        # autograd graph of that step is alive; routing a captured gradient through it would make that stream wait on the
        # capturing one, which CUDA forbids.  The module's own .grad tensors are left alone.
        aliases = {n: p.detach().requires_grad_(True) for n, p in named}
        inputs = list(aliases.values()) + ([self.raw] if need_raw_grad else [])

        def run():
            out = torch.func.functional_call(module, aliases, (self.raw,))
            grads = torch.autograd.grad(out, inputs, self.grad_out, allow_unused=True) if inputs else ()
            grads = [g if g is not None else torch.zeros_like(t) for g, t in zip(grads, inputs)]
            return out, grads

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out, grads = run()
            self.param_grads = list(grads[:len(self.params)])
            self.flat_grads = torch.cat([g.reshape(-1) for g in self.param_grads]) if self.params else None
            self.grad_raw = grads[len(self.params)] if need_raw_grad else None

    def replay(self):
        self.graph.replay()
        return self.out

    def assign_grads(self):
        """Point ``p.grad`` of the module's parameters at the graph's (static) gradient tensors, for an optimiser."""
        for p, g in zip(self.params, self.param_grads):
            p.grad = g
