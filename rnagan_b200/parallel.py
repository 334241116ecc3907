"""Data-parallel plumbing: one process per GPU, NCCL over NVLink via torch.distributed (gloo on CPU in tests).

Training shards by batch: every rank runs the three steps on its own shard with LOCAL
BatchNorm / latent-standardisation / gradient-norm statistics (what wrapping the reference in DDP would do) and the
parameter gradients are averaged before each optimiser step.  Tile synthesis needs no collective.
"""
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, world)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def allreduce_mean_(tensors, bucket_bytes=256 << 20):
    """In-place mean over ranks of a list of same-dtype tensors, bucketed into flat buffers."""
    if not tensors:
        return
    world = dist.get_world_size()
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket]) if len(bucket) > 1 else bucket[0].reshape(-1)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        if len(bucket) > 1:
            off = 0
            for t in bucket:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
        bucket, size = [], 0
    for t in tensors:
        nb = t.numel() * t.element_size()
        if size + nb > bucket_bytes and bucket:
            flush()
        bucket.append(t)
        size += nb
    flush()


def shard_range(n, rank, world):
    """Contiguous split of n units over `world` ranks (tile synthesis: independent units, no collective)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradSync:
    """Flat fp32 gradient buffer of one network (every ``p.grad`` is a view into it) + overlapped all-reduce.

    ``layer_done(*params)`` is called by the engines as soon as a layer's gradients are final: the layer's slice of
    the flat buffer is all-reduced (SUM) asynchronously on NCCL's stream while the rest of backward keeps running;
    ``finish()`` reduces whatever was not announced, waits for the handles and returns the scale (1/world) that
    the fused Adam applies to the gradients.  With a single process everything is a no-op and the scale is 1.
    """

    def __init__(self, module):
        params = [p for p in module.parameters()]
        self.params = params
        offs, off = {}, 0
        for p in params:
            offs[id(p)] = (off, p.numel())
            off += (p.numel() + 3) // 4 * 4            # keep every tensor 16-byte aligned for the vectorised Adam
        self.offs = offs
        self.flat = torch.zeros(max(off, 4), dtype=torch.float32, device=params[0].device)
        for p in params:
            o, n = offs[id(p)]
            p.grad = self._view(o, n, p)
        self._pending = []
        self._done = set()

    def _view(self, o, n, p):
        """Gradient view with the parameter's own memory layout (conv weights the engines keep channels_last get a
        channels_last gradient: the wgrad kernels write it natively and Adam walks all four tensors in lockstep)."""
        seg = self.flat[o:o + n]
        if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
            N, C, H, W = p.shape
            return seg.view(N, H, W, C).permute(0, 3, 1, 2)
        return seg.view_as(p)

    def rebind(self, p):
        """Re-create the gradient view of one parameter after its layout changed."""
        o, n = self.offs[id(p)]
        p.grad = self._view(o, n, p)

    @staticmethod
    def world():
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def _owns(self, p):
        o, n = self.offs[id(p)]
        return p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * o

    def layer_done(self, *params):
        if self.world() == 1:
            return
        params = [p for p in params if p is not None and id(p) not in self._done]
        if not params:
            return
        if all(self._owns(p) for p in params):
            lo = min(self.offs[id(p)][0] for p in params)
            hi = max(self.offs[id(p)][0] + self.offs[id(p)][1] for p in params)
            # the slice may only cover tensors that are final: require the announced set to be contiguous
            covered = sum((self.offs[id(p)][1] + 3) // 4 * 4 for p in params)
            if covered >= hi - lo:
                self._pending.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
                self._done.update(id(p) for p in params)
                return
        for p in params:
            self._pending.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True))
            self._done.add(id(p))

    def finish(self):
        w = self.world()
        if w == 1:
            return 1.0
        rest = [p for p in self.params if id(p) not in self._done and p.grad is not None]
        # merge contiguous leftovers into as few collectives as possible
        run = []
        for p in rest + [None]:
            if p is not None and self._owns(p) and (not run or self.offs[id(run[-1])][0] + (self.offs[id(run[-1])][1] + 3) // 4 * 4
                                                      == self.offs[id(p)][0]):
                run.append(p)
                continue
            if run:
                lo = self.offs[id(run[0])][0]
                hi = self.offs[id(run[-1])][0] + self.offs[id(run[-1])][1]
                self._pending.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
                run = []
            if p is not None:
                if self._owns(p):
                    run = [p]
                else:
                    self._pending.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True))
        for h in self._pending:
            h.wait()
        self._pending, self._done = [], set()
        return 1.0 / w
