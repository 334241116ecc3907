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


# ------------------------------------------------------------------------------------------------ tile reader (host)
# PatchRNADataset of the reference (src/read_data.py:266-372): one LMDB file per slide under
# `{patch_data_path}/{wsi}/{wsi with .svs -> .db}`, key `__keys__` -> lz4-framed pickle of the patch keys, every other
# value an lz4-framed pickle `(name, raw uint8 bytes, shape)` of a BGR tile.  The container formats are read by the
# host functions of the C ABI (csrc/rg_data.cu); `pickle` is the standard library's.
class LMDBFile:
    """Read-only view of one LMDB file (`lmdb.open(path, subdir=False, readonly=True, lock=False)` +
    `txn.get` / `txn.stat()['entries']`, src/read_data.py:314-320, 346-351)."""

    def __init__(self, path):
        from . import _lib
        import os
        self._L = _lib.lib()
        self.path = path
        self._h = self._L.rg_lmdb_open(os.fsencode(path))
        if not self._h:
            raise OSError(self._L.rg_last_error().decode("utf-8", "replace"))

    def stat(self):
        import ctypes
        e, p, d = ctypes.c_ulonglong(), ctypes.c_uint(), ctypes.c_uint()
        from . import _lib
        _lib.check(self._L.rg_lmdb_stat(self._h, ctypes.byref(e), ctypes.byref(p), ctypes.byref(d)), "rg_lmdb_stat")
        return {"entries": e.value, "psize": p.value, "depth": d.value}

    def get(self, key, default=None):
        """The value stored under `key` (bytes) as a bytes copy, or `default`."""
        import ctypes
        from . import _lib
        val, n = ctypes.c_void_p(), ctypes.c_size_t()
        rc = self._L.rg_lmdb_get(self._h, key, len(key), ctypes.byref(val), ctypes.byref(n))
        if rc == -5:                                   # RG_ENOTFOUND
            return default
        _lib.check(rc, "rg_lmdb_get")
        return ctypes.string_at(val.value, n.value)

    def close(self):
        if getattr(self, "_h", None):
            self._L.rg_lmdb_close(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def lz4f_decompress(data):
    """`lz4framed.decompress(data)` (src/read_data.py:318, 332): one LZ4 frame -> bytes."""
    import ctypes
    from . import _lib
    L = _lib.lib()
    size = ctypes.c_longlong(-1)
    cap = max(1 << 16, 4 * len(data))
    while True:
        buf = ctypes.create_string_buffer(cap)
        n = L.rg_lz4f_decompress(data, len(data), buf, cap, ctypes.byref(size))
        if n >= 0:
            return buf.raw[:n]
        if n == -2 and cap < (1 << 34):
            cap = max(2 * cap, size.value if size.value > 0 else 0)
            continue
        raise ValueError("lz4f_decompress: malformed LZ4 frame")


class PatchRNADataset(torch.utils.data.Dataset):
    """Drop-in for the reference's PatchRNADataset (src/read_data.py:266-372): same constructor arguments, same per-slide
    patch sampling (`random.sample` on the global `random` generator, in CSV row order), same item dictionary
    `{'image', 'rna_data', 'labels'}`.

    `raw=True` (not in the reference) returns the tile as it is stored -- uint8 HWC, BGR -- for
    `DevicePrefetcher(bgr=True)`, which normalises on the GPU; the default applies the reference's host path: BGR->RGB,
    `permute(2, 0, 1)` and `transforms` (src/read_data.py:336-343)."""

    def __init__(self, patch_data_path, csv_path, img_size, transforms=None, max_patches_total=300, quick=False, le=None,
                 raw=False):
        self.patch_data_path, self.csv_path, self.img_size = patch_data_path, csv_path, img_size
        self.transforms, self.max_patches_total, self.quick, self.le, self.raw = transforms, max_patches_total, quick, le, raw
        self.keys, self.images, self.filenames, self.labels, self.lmdbs_path, self.rna_data_arrays = [], [], [], [], [], []
        self._open = {}
        self._preprocess()

    def _preprocess(self):
        import os
        import pickle
        import random
        import pandas as pd
        if isinstance(self.csv_path, str):
            csv_file = pd.read_csv(self.csv_path)
            csv_file["patch_data_path"] = [self.patch_data_path] * csv_file.shape[0]
            csv_file["labels"] = [0] * csv_file.shape[0]
        else:
            csv_file = self.csv_path
        if self.quick:
            csv_file = csv_file.sample(150)
        rna_cols = [c for c in csv_file.columns if "rna_" in c]
        for _, row in csv_file.iterrows():
            wsi = row["wsi_file_name"]
            rna = torch.tensor(row[rna_cols].values.astype(np.float32), dtype=torch.float32)
            label = np.asarray(row["labels"])
            if self.le is not None:
                label = self.le.transform(label.reshape(-1, 1))
            label = torch.tensor(label, dtype=torch.float32)
            path = os.path.join(row["patch_data_path"], wsi, wsi.replace(".svs", ".db"))
            try:
                with LMDBFile(path) as db:
                    n_patches = db.stat()["entries"] - 1
                    keys = pickle.loads(lz4f_decompress(db.get(b"__keys__")))
                picked = random.sample(list(range(n_patches)), min(n_patches, self.max_patches_total))
            except Exception:
                print("Error with db {}".format(path))
                continue
            for i in picked:
                self.images.append(i)
                self.filenames.append(wsi)
                self.labels.append(label)
                self.lmdbs_path.append(path)
                self.keys.append(keys[i])
                self.rna_data_arrays.append(rna)

    def decompress_and_deserialize(self, lmdb_value):
        import pickle
        try:
            _, img_arr, img_shape = pickle.loads(lz4f_decompress(lmdb_value))
        except Exception:
            return None
        image = np.frombuffer(img_arr, dtype=np.uint8).reshape(img_shape)
        if self.raw:
            return torch.from_numpy(np.copy(image))                       # uint8 HWC, BGR as stored
        return torch.from_numpy(np.ascontiguousarray(image[..., ::-1])).permute(2, 0, 1)   # cv2.COLOR_BGR2RGB + CHW

    def __len__(self):
        return len(self.images)

    def _db(self, path):
        db = self._open.get(path)
        if db is None:
            db = self._open[path] = LMDBFile(path)        # kept open per worker (the reference re-opens per item)
        return db

    def __getitem__(self, idx):
        image = self.decompress_and_deserialize(self._db(self.lmdbs_path[idx]).get(self.keys[idx]))
        if image is None:
            print(self.lmdbs_path[idx])
        elif not self.raw and self.transforms is not None:
            image = self.transforms(image)
        return {"image": image, "rna_data": self.rna_data_arrays[idx], "labels": self.labels[idx]}

    def __getstate__(self):                               # DataLoader workers: handles are per process
        st = dict(self.__dict__)
        st["_open"] = {}
        return st
