"""Data-parallel plumbing: one process per GPU, NCCL over NVLink via torch.distributed (gloo on CPU in tests).

Training shards by batch: every rank runs the three steps on its own shard with LOCAL
BatchNorm / latent-standardisation / gradient-norm statistics (what wrapping the reference in DDP would do) and the
parameter gradients are averaged before each optimiser step.  Tile synthesis needs no collective.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, world)."""
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


# ------------------------------------------------------------------------------------------------ peer exchange
# NCCL's all-reduce kernels need whole SMs for the ~1 ms they run, and a persistent tile-engine grid then executes the
# CTAs they displaced as a second wave: the all-reduce is paid in full although it "overlaps".
# PeerExchange moves the gradients with the COPY ENGINES instead: the flat gradient buffer lives in symmetric memory
# (torch.distributed._symmetric_memory: every rank maps every peer's buffer over NVLink), and one all-reduce of a
# bucket [lo, hi) is
#     barrier | pull my slice of the bucket from every peer (DMA) | sum the world copies in rank order (one HBM-bound
#     kernel, rg_slices_sum) | barrier | pull every peer's reduced slice (DMA) | barrier
# on a side stream.  No SM is held while bytes move, the summation order is fixed (bit-reproducible, identical on all
# ranks), and per GPU 2 (world-1)/world of the bucket crosses NVLink in each direction -- the same as a ring.
def slice_bounds(n, world, align=4):
    """Contiguous split of [0, n) into `world` slices with `align`-multiple lengths (trailing slices may be short or
    empty); returns [(lo, hi)] per rank."""
    per = ((n + align - 1) // align + world - 1) // world * align
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def exchange_pull_reduce(rank, world, peers, stage, lo, n, copy, reduce):
    """Phase 1 on `rank`: stage[r, :len] <- peer r's copy of this rank's slice of bucket [lo, lo+n), then
    own[lo + slice] = sum_r stage[r] (ascending r).  `peers[r]` is rank r's flat buffer as seen from this rank."""
    a, b = slice_bounds(n, world)[rank]
    if b <= a:
        return 0
    for step in range(world):
        r = (rank - step) % world                # start with the local copy, then walk the peers round-robin
        copy(stage[r, :b - a], peers[r][lo + a:lo + b])
    reduce(stage, world, b - a, peers[rank][lo + a:lo + b])
    return b - a


def exchange_gather(rank, world, peers, lo, n, copy):
    """Phase 2 on `rank`: own[lo + slice_r] <- peer r's reduced slice, for every other rank r."""
    own = peers[rank]
    bounds = slice_bounds(n, world)
    for step in range(1, world):
        r = (rank - step) % world
        a, b = bounds[r]
        if b > a:
            copy(own[lo + a:lo + b], peers[r][lo + a:lo + b])


class _EventHandle:
    def __init__(self, event, device):
        self.event, self.device = event, device

    def wait(self):
        torch.cuda.current_stream(self.device).wait_event(self.event)


class PeerExchange:
    """SUM all-reduce of ranges of one flat fp32 buffer through peer-mapped memory and the copy engines (see above).
    Opt-in (RG_DP_EXCHANGE=ce); every rank must issue the same sequence of `allreduce` calls."""

    def __init__(self, numel, device, group=None, use_nvls=False):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("PeerExchange needs CUDA devices with peer access (there is no CPU path)")
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device, self.numel = device, int(numel)
        self.flat = symm.empty(self.numel, dtype=torch.float32, device=device)
        self.flat.zero_()
        self.hdl = symm.rendezvous(self.flat, group)
        self.peers = [self.flat if r == self.rank else self.hdl.get_buffer(r, (self.numel,), torch.float32, 0)
                      for r in range(self.world)]
        # NVSwitch multicast mapping of the same buffer (0 when the fabric has no multicast support)
        self.mc_ptr = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.nvls = use_nvls and self.mc_ptr != 0
        if use_nvls and not self.nvls and self.rank == 0:
            import warnings
            warnings.warn("RG_DP_EXCHANGE=nvls: no multicast support on this fabric, using the copy-engine exchange")
        per = slice_bounds(self.numel, self.world)[0]
        self.stage = None if self.nvls else torch.empty(self.world, max(per[1] - per[0], 4), dtype=torch.float32,
                                                        device=device)
        self.stream = torch.cuda.Stream(device=device)

    @staticmethod
    def _copy(dst, src):
        dst.copy_(src, non_blocking=True)          # same dtype, contiguous: cudaMemcpyAsync -> copy engine

    @staticmethod
    def _reduce(stage, world, n, out):
        from . import ops
        ops.slices_sum(stage, world, n, out)

    def allreduce(self, lo, hi):
        """Asynchronous SUM over ranks of flat[lo:hi) (lo a multiple of 4 floats; hi is rounded up to one -- the
        flat buffer is padded).  Returns a handle whose wait() orders the current stream after the exchange."""
        hi = min((hi + 3) // 4 * 4, self.numel)
        if lo % 4 != 0 or hi <= lo:
            raise ValueError(f"PeerExchange.allreduce: bad range [{lo}, {hi})")
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        if self.nvls:
            # in-switch reduction: this rank reduces + broadcasts its slice of the bucket through the multicast mapping
            from . import ops
            a, b = slice_bounds(hi - lo, self.world)[self.rank]
            with torch.cuda.stream(self.stream):
                self.hdl.barrier(channel=0)        # every rank's bucket is final and visible
                if b > a:
                    ops.nvls_allreduce(self.mc_ptr, lo + a, b - a)
                self.hdl.barrier(channel=0)        # every slice has been written back to every copy
                done = torch.cuda.Event()
                done.record(self.stream)
            return _EventHandle(done, self.device)
        with torch.cuda.stream(self.stream):
            self.hdl.barrier(channel=0)            # every rank's bucket is final
            exchange_pull_reduce(self.rank, self.world, self.peers, self.stage, lo, hi - lo, self._copy, self._reduce)
            self.hdl.barrier(channel=0)            # every slice is reduced, nobody still reads un-reduced data
            exchange_gather(self.rank, self.world, self.peers, lo, hi - lo, self._copy)
            self.hdl.barrier(channel=0)            # nobody still reads this rank's slice: the buffer may be rewritten
            done = torch.cuda.Event()
            done.record(self.stream)
        return _EventHandle(done, self.device)


def exchange_mode():
    """How data-parallel gradients are summed (RG_DP_EXCHANGE overrides):
      'ce'    PeerExchange over symmetric memory and the copy engines: no SM is held while bytes move
      'nvls'  PeerExchange with the in-switch multimem reduction (rg_nvls_allreduce): each rank's kernel touches 1/world
              of the bytes, so its SM time shrinks with the world size while the copy-engine schedule grows
      'nccl'  torch.distributed all_reduce (always used on CPU / gloo)
    Default: 'ce' up to 3 ranks, 'nvls' from 4 (falls back to 'ce' when the fabric has no multicast).  Measured on B200s
    (profiles/r2_dp_exchange_ab.txt, ms per iteration): 2 GPUs nccl 13.9 / ce 13.7 / nvls 13.0 vs ce 12.1 after the
    kernel work; 8 GPUs nccl 14.4 / ce 14.2, then ce 12.50 / nvls 12.50 with nvls ahead end to end (611 vs 597 steps/s)."""
    mode = os.environ.get("RG_DP_EXCHANGE", "auto").lower()
    if mode == "auto":
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        mode = "nvls" if world >= 4 else "ce"
    return mode


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
        self.xchg = None
        # announced layers are merged into buckets of at least this many floats before they are exchanged (0: every
        # announcement is reduced at once, the NCCL default).  The peer exchange costs three barriers and 2(world-1)
        # copies per bucket whatever its size, and most layers are tiny: RG_DP_BUCKET_MB (default 16) in that mode.
        self.bucket_floats = 0
        self._open = None
        if self.world() > 1 and exchange_mode() in ("ce", "nvls") and params[0].device.type == "cuda":
            try:
                self.xchg = PeerExchange(max(off, 4), params[0].device, use_nvls=exchange_mode() == "nvls")
            except Exception as ex:          # no peer access / symmetric memory on this system: NCCL still works
                import warnings
                warnings.warn(f"rnagan_b200: peer-memory gradient exchange unavailable ({type(ex).__name__}: {ex}); "
                              "using the NCCL all-reduce")
                self.xchg = None
        if self.xchg is not None:
            self.flat = self.xchg.flat
            self.bucket_floats = int(float(os.environ.get("RG_DP_BUCKET_MB", "16")) * (1 << 20)) // 4
        else:
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

    def _reduce_range(self, lo, hi):
        if self.xchg is not None:
            return self.xchg.allreduce(lo, hi)
        return dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True)

    def _flush_open(self):
        if self._open is not None:
            self._pending.append(self._reduce_range(*self._open))
            self._open = None

    def _submit(self, lo, hi):
        """Reduce flat[lo:hi) now, or merge it into the open bucket when bucketing is on (adjacent ranges only: the
        engines announce layers in backward order, i.e. in descending offsets)."""
        if self.bucket_floats <= 0:
            self._pending.append(self._reduce_range(lo, hi))
            return
        r4 = lambda v: (v + 3) // 4 * 4  # noqa: E731   (tensor offsets are padded to 4 floats)
        if self._open is not None and r4(hi) == self._open[0]:
            self._open[0] = lo
        elif self._open is not None and r4(self._open[1]) == lo:
            self._open[1] = hi
        else:
            self._flush_open()
            self._open = [lo, hi]
        if self._open[1] - self._open[0] >= self.bucket_floats:
            self._flush_open()

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
                self._submit(lo, hi)
                self._done.update(id(p) for p in params)
                return
        for p in params:
            self._pending.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True))
            self._done.add(id(p))

    def launch_rest(self):
        """Submit (asynchronously) the reduction of everything the engines did not announce with layer_done."""
        if self.world() == 1:
            return
        self._flush_open()
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
                self._pending.append(self._reduce_range(lo, hi))
                run = []
            if p is not None:
                if self._owns(p):
                    run = [p]
                else:
                    self._pending.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True))
        self._done.update(id(p) for p in rest)

    def wait(self):
        """Order the current stream after every submitted reduction; returns the gradient scale 1/world."""
        w = self.world()
        if w == 1:
            return 1.0
        for h in self._pending:
            h.wait()
        self._pending, self._done = [], set()
        return 1.0 / w

    def finish(self):
        self.launch_rest()
        return self.wait()
