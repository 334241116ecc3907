"""Per-epoch sample grid (host side): what torchgan's Trainer writes to `recon` at the end of every epoch [tg] -- the
generator in eval mode on a fixed `test_noise`, tiled `nrow` per row with 2 pixels of padding, min-max normalised, saved as
`{recon}/epoch{N}_{model}.png` (call site: `Trainer(..., sample_size=64, recon=args.image_dir)`,
src/histopathology_gan.py:298-302).  The layout and the uint8 conversion follow torchvision.utils.make_grid / save_image
(checked pixel for pixel in tests/test_host_cpu.py); the PNG is written with zlib so the path needs neither torchvision nor
PIL.  Pure host code on an already computed [N, C, H, W] batch: nothing here is on the training hot path.
"""
import math
import os
import struct
import zlib

import numpy as np
import torch


def make_grid(images, nrow=8, padding=2, normalize=True, pad_value=0.0):
    """[N, C, H, W] float tensor -> [C', H', W'] float grid in torchvision's layout (C' = 3; grey images are repeated)."""
    t = images.detach().to(device="cpu", dtype=torch.float32)
    if t.dim() != 4:
        raise ValueError(f"expected [N, C, H, W], got {tuple(t.shape)}")
    if t.size(1) == 1:
        t = t.expand(-1, 3, -1, -1)
    t = t.clone()
    if normalize:
        lo, hi = float(t.min()), float(t.max())
        t.clamp_(min=lo, max=hi).sub_(lo).div_(max(hi - lo, 1e-5))
    n = t.size(0)
    if n == 1:                                   # torchvision returns a single image as is, without a border
        return t[0]
    xmaps = min(nrow, n)
    ymaps = int(math.ceil(n / xmaps))
    h, w = t.size(2) + padding, t.size(3) + padding
    grid = t.new_full((t.size(1), h * ymaps + padding, w * xmaps + padding), pad_value)
    k = 0
    for y in range(ymaps):
        for x in range(xmaps):
            if k >= n:
                break
            grid[:, y * h + padding:(y + 1) * h, x * w + padding:(x + 1) * w] = t[k]
            k += 1
    return grid


def to_uint8_hwc(grid):
    """torchvision.utils.save_image's conversion: mul(255).add(0.5).clamp(0, 255) -> uint8, HWC."""
    return grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()


def write_png(path, hwc_u8):
    """Minimal RGB / grey 8-bit PNG encoder (filter type 0 on every row)."""
    a = np.ascontiguousarray(hwc_u8)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] not in (1, 3):
        raise ValueError("write_png expects uint8 [H, W, 1 or 3]")
    h, w, c = a.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), a.reshape(h, w * c)], axis=1).tobytes()

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    png = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 0)) +
           chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(png)
    return path


def save_image_grid(images, path, nrow=8, padding=2, normalize=True):
    return write_png(path, to_uint8_hwc(make_grid(images, nrow=nrow, padding=padding, normalize=normalize)))
