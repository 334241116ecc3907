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
def generate_tiles(generator, betavae, gene_exp, sample_size, chunk=10, device=None, out=None):
    """Returns a device tensor [sample_size, S, S, C] fp32 in [0, 1] (NHWC)."""
    device = next(generator.parameters()).device if device is None else device
    if next(betavae.parameters()).device != torch.device(device):
        betavae = betavae.to(device)
    eng, lat = _latent(generator, betavae, gene_exp, sample_size, device)
    S, C = eng.size, eng.Cimg
    if out is None:
        out = torch.empty(sample_size, S, S, C, dtype=F32, device=device)
    for lo in range(0, sample_size, chunk):
        hi = min(sample_size, lo + chunk)
        if hasattr(eng, "w_colT_last"):            # the last kernel writes (x + 1) / 2 in NHWC itself
            eng.forward(lat[lo:hi], tag=f"synth{hi - lo}", training=generator.training, out=out[lo:hi], unit_nhwc=True)
        else:
            img = eng.forward(lat[lo:hi], tag=f"synth{hi - lo}", training=generator.training)
            ops.tiles_to_unit_nhwc(img, out[lo:hi])
    return out


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
        return out.cpu().numpy()
    tiles = generate_tiles(generator, betavae, gene_exp, sample_size, chunk=10, device=trainer.device)
    return tiles.cpu().numpy()
