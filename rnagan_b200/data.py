"""Device-side half of the data path: the reference decodes uint8 HWC BGR tiles from LMDB and
runs cv2.cvtColor + permute + ConvertImageDtype + Normalize on the CPU inside DataLoader workers
(src/read_data.py:336-343, src/histopathology_gan.py:106-109), then ships fp32 NCHW batches to the GPU.

``DevicePrefetcher`` keeps the tiles uint8 until they are on the device -- a quarter of the host->device bytes -- and
normalises them there with one kernel (``rg_tiles_u8_to_nchw``, bit-identical to the CPU transforms).  The copies of
batch i+1 run on a side stream while batch i trains.  LMDB / lz4 / CSV reading stays on the CPU (out of scope: no GPU
work in it); any DataLoader that yields ``{'image': uint8 [B,S,S,C] or float [B,C,S,S], 'rna_data': float [B,F], ...}``
dicts can be wrapped.
"""
import torch

from . import ops


def normalise_tiles(tiles_u8, bgr=True, out=None):
    """uint8 [B, S, S, C] device tensor -> fp32 NCHW in [-1, 1] (what the reference's `transforms_` produce)."""
    return ops.tiles_u8_to_nchw(tiles_u8, out=out, swap_rb=bgr)


class DevicePrefetcher:
    """Iterate a DataLoader one batch ahead: pinned staging, asynchronous host->device copies on a side stream,
    uint8 tiles normalised on the device.  Yields dicts whose tensors live on `device`."""

    def __init__(self, loader, device, bgr=True):
        self.loader, self.device, self.bgr = loader, torch.device(device), bgr
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher needs a CUDA device (there is no CPU path)")
        self.stream = torch.cuda.Stream(device=self.device)

    def _stage(self, batch):
        out, ev = {}, torch.cuda.Event()
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                if not torch.is_tensor(v):
                    out[k] = v
                    continue
                if v.device.type != "cuda" and not v.is_pinned():
                    v = v.pin_memory()
                d = v.to(self.device, non_blocking=True)
                if k == "image" and d.dtype == torch.uint8:
                    d = normalise_tiles(d.contiguous(), bgr=self.bgr)
                out[k] = d
            ev.record(self.stream)
        return out, ev

    def __iter__(self):
        nxt = None
        for batch in self.loader:
            staged = self._stage(batch)
            if nxt is not None:
                cur, ev = nxt
                torch.cuda.current_stream(self.device).wait_event(ev)
                yield cur
            nxt = staged
        if nxt is not None:
            cur, ev = nxt
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield cur

    def __len__(self):
        return len(self.loader)
