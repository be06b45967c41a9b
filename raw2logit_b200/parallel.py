"""Data parallelism for the ISP path: one process per GPU, the image batch sharded across ranks, and ONE collective
per step -- an all-reduce of every gradient (the 132 ISP scalars ride in the same flat bucket as the task-model
gradients, never as 7 latency-bound 4..324-byte messages).  The reference has no distributed code
(single process, ``gpus=1``, train.py:362); this is the new build's sharding of its batch dimension.

Forward needs no communication (images are independent units); BatchNorm statistics stay per replica, as in the
single-GPU reference (no SyncBN).
"""
import torch
import torch.distributed as dist


def shard_batch(x, rank, world):
    """Rank ``rank`` of ``world`` takes images rank, rank+world, ... of the global batch (first dimension)."""
    return x[rank::world]


def flat_gradients(parameters):
    """All existing gradients of ``parameters`` in one contiguous vector, plus the (param, offset, numel) layout."""
    params = [p for p in parameters if p.grad is not None]
    layout, n = [], 0
    for p in params:
        layout.append((p, n, p.numel()))
        n += p.numel()
    if not params:
        return None, layout
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    return flat, layout


def allreduce_gradients(parameters, world=None, group=None, average=True):
    """One all-reduce over the flat gradient bucket; writes the reduced values back into ``p.grad``.

    Returns the number of elements communicated (0 when nothing had a gradient).  With ``world == 1`` (or no
    initialised process group) it is a no-op."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    flat, layout = flat_gradients(parameters)
    if flat is None:
        return 0
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        for p, off, n in layout:
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
    return flat.numel()


def data_parallel_step(model, batch, optimizer=None, rank=0, world=1, group=None):
    """One training step of ``model`` (a ``raw2logit_b200.model.LitModel``) on this rank's shard of ``batch``:
    forward, loss, backward, single-bucket gradient all-reduce, optional optimizer step.  Returns the local loss."""
    x, y = batch
    x, y = shard_batch(x, rank, world), shard_batch(y, rank, world)
    for p in model.parameters():
        p.grad = None
    loss = model.update_step((x, y))
    loss.backward()
    allreduce_gradients(model.parameters(), world=world, group=group)
    if optimizer is not None:
        optimizer.step()
    return loss.detach()


class PeerExchange:
    """Exchange buffers of the all-reduce that is fused into the backward kernel (``r2l_isp_backward_dp``,
    include/r2l_isp.h): one symmetric-memory allocation per rank, every rank's buffer mapped into every process, so
    the kernel's last CTA pushes its 132 gradients to all ranks over NVLink and adds the slots up itself -- no
    collective launch.  ``next()`` hands out the per-call descriptor (the epoch must advance in lock step on all
    ranks, i.e. every rank makes every call).

    Needs one process per GPU on one node with peer access (``torch.distributed._symmetric_memory``).  The gradients
    of the task model still go through ``allreduce_gradients`` (NCCL); this object serves the ISP-only step."""

    def __init__(self, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised process group")
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = _lib.load().r2l_isp_exchange_bytes(self.world)
        if nbytes == 0:
            raise RuntimeError(f"world size {self.world} is not served by the fused exchange")
        self.buffer = symm.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.buffer.zero_()
        self.handle = symm.rendezvous(self.buffer, group.group_name)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                          # every rank's flags are zero before anyone writes
        self.peers = int(self.handle.buffer_ptrs_dev)                # device array of `world` buffer pointers
        self.calls = 0

    def next(self, average=True):
        """Descriptor for the next ``r2l_isp_backward_dp`` call (``_lib.IspAllreduce``).  The epoch is kept by the
        kernel itself in the exchange buffer (``R2L_EPOCH_DEVICE``), so a captured launch can be replayed."""
        from . import _lib
        self.calls += 1
        return _lib.IspAllreduce(self.world, self.rank, self.peers, _lib.EPOCH_DEVICE,
                                 1.0 / self.world if average else 1.0)


def enable_fused_gradient_exchange(group=None, average=True):
    """From now on the ISP's fused backward (module API: ``ParametrizedProcessing`` under autograd) returns parameter
    gradients that are already summed (``average=False``) or averaged over the ranks of ``group``: the exchange runs
    inside the backward kernel (``PeerExchange``).  Only the 132 in-kernel gradients are exchanged: all-reduce every
    *other* parameter afterwards -- the task model's and, in adversarial mode, ``processor.additive_layer`` (its gradient
    is a separate batch sum and stays rank-local), e.g. ``allreduce_gradients([*task_model.parameters(),
    processor.additive_layer])``.  The staged path (``track_stages=True``) does not take part in the exchange and warns
    when it runs while the exchange is enabled.  Every rank must run the same sequence of ISP backward calls.
    Returns the ``PeerExchange``; ``disable_fused_gradient_exchange()`` restores local gradients."""
    from . import ops
    exchange = PeerExchange(group)
    ops.set_gradient_exchange(exchange, average)
    return exchange


def disable_fused_gradient_exchange():
    from . import ops
    ops.set_gradient_exchange(None)
