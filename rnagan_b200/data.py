"""Device-side half of the data path: the reference decodes uint8 HWC BGR tiles from LMDB and
runs cv2.cvtColor + permute + ConvertImageDtype + Normalize on the CPU inside DataLoader workers
(src/read_data.py:336-343, src/histopathology_gan.py:106-109), then ships fp32 NCHW batches to the GPU.

``DevicePrefetcher`` keeps the tiles uint8 until they are on the device -- a quarter of the host->device bytes -- and
normalises them there with one kernel (``rg_tiles_u8_to_nchw``, bit-identical to the CPU transforms).  The copies of
batch i+1 run on a side stream while batch i trains.  LMDB / lz4 / CSV reading stays on the CPU (out of scope: no GPU
work in it); any DataLoader that yields ``{'image': uint8 [B,S,S,C] or float [B,C,S,S], 'rna_data': float [B,F], ...}``
dicts can be wrapped.

``RNAScaler`` is the host half the reference runs once per job before the first batch: the zero-safe log transform and the
per-gene standardisation of the expression table (src/histopathology_gan.py:133-151, src/gan_utils.py:79-98,
src/read_data.py:467-498 `normalize_dfs`).  It is setup-time float64 table arithmetic on the [patients, genes] matrix
(no per-step work, nothing for the GPU to accelerate), kept here so the path needs neither pandas nor scikit-learn.
"""
import numpy as np
import torch

from . import ops


def normalise_tiles(tiles_u8, bgr=True, out=None):
    """uint8 [B, S, S, C] device tensor -> fp32 NCHW in [-1, 1] (what the reference's `transforms_` produce)."""
    return ops.tiles_u8_to_nchw(tiles_u8, out=out, swap_rb=bgr)


def log_expression(values):
    """The reference's `_get_log` (src/histopathology_gan.py:133-136): natural log with zeros mapped to 0 -- it routes
    zeros through NaN, so NaN inputs and the NaN that log gives negative counts become 0 as well."""
    x = np.asarray(values, dtype=np.float64)
    out = np.zeros_like(x)
    pos = x > 0
    np.log(x, out=out, where=pos)
    return out


class RNAScaler:
    """`StandardScaler().fit_transform` / `.transform` / `.inverse_transform` on log expression, as the reference drivers
    use it (fit on the training table, src/histopathology_gan.py:148-149; reused for validation / test tables and to map
    decoded profiles back, src/read_data.py:495-496, src/betaVAE_sample.py:132).  Population variance (ddof 0); a gene
    with zero variance keeps scale 1, like scikit-learn's `_handle_zeros_in_scale`."""

    def __init__(self, log=True):
        self.log = log
        self.mean_ = self.var_ = self.scale_ = None
        self.n_samples_seen_ = 0

    def _pre(self, values):
        x = log_expression(values) if self.log else np.asarray(values, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError(f"expected a [patients, genes] table, got shape {x.shape}")
        return x

    def fit(self, values):
        x = self._pre(values)
        if x.shape[0] == 0:
            raise ValueError("cannot fit RNAScaler on an empty table")
        self.n_samples_seen_ = x.shape[0]
        self.mean_ = x.mean(axis=0)
        self.var_ = x.var(axis=0)
        scale = np.sqrt(self.var_)
        # scikit-learn treats a scale below 10 * eps as a constant feature (sklearn.preprocessing._data)
        scale[scale < 10 * np.finfo(np.float64).eps] = 1.0
        self.scale_ = scale
        return self

    def transform(self, values):
        if self.mean_ is None:
            raise RuntimeError("RNAScaler.transform called before fit")
        x = self._pre(values)
        if x.shape[1] != self.mean_.shape[0]:
            raise ValueError(f"table has {x.shape[1]} genes, scaler was fitted on {self.mean_.shape[0]}")
        return (x - self.mean_) / self.scale_

    def fit_transform(self, values):
        return self.fit(values).transform(values)

    def inverse_transform(self, scaled):
        """Back to LOG expression (the reference never undoes the log, src/betaVAE_training.py:196-197)."""
        if self.mean_ is None:
            raise RuntimeError("RNAScaler.inverse_transform called before fit")
        return np.asarray(scaled, dtype=np.float64) * self.scale_ + self.mean_


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

    def _hand_over(self, staged):
        """Order the consumer's stream after the copies and tell the allocator who reads these blocks now."""
        cur, ev = staged
        stream = torch.cuda.current_stream(self.device)
        stream.wait_event(ev)
        for v in cur.values():
            if torch.is_tensor(v) and v.device.type == "cuda":
                v.record_stream(stream)
        return cur

    def __iter__(self):
        nxt = None
        for batch in self.loader:
            staged = self._stage(batch)
            if nxt is not None:
                yield self._hand_over(nxt)
            nxt = staged
        if nxt is not None:
            yield self._hand_over(nxt)

    def __len__(self):
        return len(self.loader)
