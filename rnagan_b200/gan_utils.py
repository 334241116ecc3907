"""Tile synthesis, drop-in for ``generate_images`` of the reference (src/gan_utils.py:197-244).

Reference semantics kept: CPU uniform(-0.3, 0.3) noise, additive conditioning
with the betaVAE latent, batch standardisation over the WHOLE sample, generator run in chunks of 10 in whatever mode
it is in (train mode => batch statistics per chunk), un-normalise (x+1)/2, NHWC float32 numpy result.
``generate_tiles`` is the throughput path for large jobs (config 4): any chunk size, device-resident result.
"""
import numpy as np
import torch

from . import ops

F32 = torch.float32
BF16 = torch.bfloat16


_PINNED = {}


def _latent(generator, betavae, gene_exp, sample_size, device):
    eng = generator._engine()
    E = generator.encoding_dims
    # pinned staging (the 8 MB draw of a 1024-tile chunk copies at PCIe speed instead of through a pageable bounce
    # buffer): two buffers used alternately, each guarded by the event of its last asynchronous copy, so the host
    # never overwrites noise a queued copy has not read yet and only blocks if it runs two calls ahead of the GPU
    ring = _PINNED.setdefault((sample_size, E), {"i": 0, "buf": [None, None], "ev": [None, None]})
    k = ring["i"] = ring["i"] ^ 1
    if ring["buf"][k] is None:
        ring["buf"][k] = torch.empty(sample_size, E, dtype=F32).pin_memory()
    if ring["ev"][k] is not None:
        ring["ev"][k].synchronize()
    noise = ring["buf"][k]
    noise.uniform_(-0.3, 0.3)                      # same CPU generator stream as torch.FloatTensor(n, E).uniform_()
    noise_d = noise.to(device, non_blocking=True)
    ring["ev"][k] = torch.cuda.Event()
    ring["ev"][k].record(torch.cuda.current_stream(device))
    z = betavae.encode_mean(gene_exp.to(device))
    lat = torch.empty(sample_size, E, dtype=BF16, device=device)
    ops.latent_prep(noise_d, z, lat_bf16=lat)
    return eng, lat


@torch.no_grad()
def generate_tiles(generator, betavae, gene_exp, sample_size, chunk=10, device=None, out=None, u8=False, bgr=False):
    """Returns a device tensor [sample_size, S, S, C] (NHWC): fp32 in [0, 1], or with u8=True the uint8 tiles
    trunc(255 * x) the reference computes on the host before cv2.imwrite (src/generate_tissue_images.py:127-129;
    bgr=True: channels reversed like its cv2.cvtColor(RGB2BGR)) -- written by the generator's last kernel."""
    device = next(generator.parameters()).device if device is None else device
    if next(betavae.parameters()).device != torch.device(device):
        betavae = betavae.to(device)
    eng, lat = _latent(generator, betavae, gene_exp, sample_size, device)
    S, C = eng.size, eng.Cimg
    if out is None:
        out = torch.empty(sample_size, S, S, C, dtype=torch.uint8 if u8 else F32, device=device)
    if (out.dtype == torch.uint8) != bool(u8):
        raise ValueError("generate_tiles: `out` must be uint8 exactly when u8=True")
    for lo in range(0, sample_size, chunk):
        hi = min(sample_size, lo + chunk)
        if hasattr(eng, "w_colT_last"):            # the last kernel writes (x + 1) / 2 (or the uint8 tile) in NHWC itself
            eng.forward(lat[lo:hi], tag=f"synth{hi - lo}", training=generator.training, out=out[lo:hi],
                        unit_nhwc=not u8, u8=u8, bgr=bgr, keep=False)
        elif u8:
            raise NotImplementedError("uint8 tile output is implemented for the transposed-conv DCGANGenerator only")
        else:
            img = eng.forward(lat[lo:hi], tag=f"synth{hi - lo}", training=generator.training)
            ops.tiles_to_unit_nhwc(img, out[lo:hi])
    return out


@torch.no_grad()
def synthesize_job(generator, betavae, profiles, first, last, batch=1024, sink=None, bgr=False, device=None):
    """Large synthesis job (BASELINE config 4: 100k tiles from RNA profiles, batch 1024): tiles [first, last) of the job,
    tile t conditioned on profile row t % len(profiles); each batch is one `generate_images`-style sample (fresh CPU
    noise, latent standardised over the batch, train-mode BN over the batch) whose uint8 NHWC tiles leave the
    generator's last kernel, cross PCIe into one of two pinned host buffers on a copy stream while the next batch
    computes, and are handed to ``sink(first_tile_index, uint8 ndarray [n, S, S, C])`` (a view valid during the call).
    Ranks of a multi-GPU job call this with their own `parallel.shard_range(n, rank, world)`; no collective.
    Returns the number of tiles produced."""
    device = next(generator.parameters()).device if device is None else torch.device(device)
    if next(betavae.parameters()).device != device:
        betavae = betavae.to(device)
    eng = generator._engine()
    S, C = eng.size, eng.Cimg
    profiles_d = profiles.to(device=device, dtype=F32)
    P = profiles_d.shape[0]
    key = ("job", batch, S, C, str(device))
    st = _PINNED.get(key)
    if st is None:
        st = {"dev": [torch.empty(batch, S, S, C, dtype=torch.uint8, device=device) for _ in range(2)],
              "host": [torch.empty(batch, S, S, C, dtype=torch.uint8).pin_memory() for _ in range(2)],
              "copy": torch.cuda.Stream(device=device)}
        _PINNED[key] = st
    cur = torch.cuda.current_stream(device)
    inflight = [None, None]                       # per slot: (copy-done event, first tile, count)

    def drain(k):
        if inflight[k] is not None:
            ev, t0, n = inflight[k]
            ev.synchronize()
            if sink is not None:
                sink(t0, st["host"][k][:n].numpy())
            inflight[k] = None

    k, done = 0, 0
    for t0 in range(first, last, batch):
        n = min(batch, last - t0)
        drain(k)                                  # slot k's host buffer is free and its device buffer has been copied out
        rows = torch.arange(t0, t0 + n, device=device) % P
        generate_tiles(generator, betavae, profiles_d[rows], n, chunk=n, device=device, out=st["dev"][k][:n], u8=True,
                       bgr=bgr)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(st["copy"]):
            st["copy"].wait_event(ready)
            st["host"][k][:n].copy_(st["dev"][k][:n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st["copy"])
        inflight[k] = (ev, t0, n)
        k ^= 1
        done += n
    drain(k)
    drain(k ^ 1)
    return done


def _to_numpy(t):
    """Device tensor -> fresh numpy array through a cached PINNED staging buffer: a pageable `.cpu()` of the 500 MB a
    640-tile call returns runs at 2-3 GB/s and was 80 % of the call; pinned DMA + one host memcpy is ~4x faster."""
    key = ("d2h", t.dtype)
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < t.numel():
        buf = _PINNED[key] = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
    stage = buf[:t.numel()].view(t.shape)
    stage.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return stage.numpy().copy()


def generate_images(trainer, gene_exp=None, sample_size=64, betavae=None):
    """Same call signature and return value as the reference: numpy float32 [sample_size, S, S, 3] in [0, 1]."""
    generator = getattr(trainer, "generator").to(trainer.device)
    if gene_exp is None:
        noise = generator.sampler(sample_size, trainer.device)[0]
        eng = generator._engine()
        lat = ops.cast_pad_bf16(noise.contiguous(), noise.shape[1])
        out = torch.empty(sample_size, eng.size, eng.size, eng.Cimg, dtype=F32, device=trainer.device)
        with torch.no_grad():
            for lo in range(0, sample_size, 10):
                hi = min(sample_size, lo + 10)
                img = eng.forward(lat[lo:hi], tag=f"synth{hi - lo}", training=generator.training)
                ops.tiles_to_unit_nhwc(img, out[lo:hi])
        return _to_numpy(out)
    tiles = generate_tiles(generator, betavae, gene_exp, sample_size, chunk=10, device=trainer.device)
    return _to_numpy(tiles)
