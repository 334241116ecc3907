"""Host half of the tile data path (csrc/rg_data.cu + rnagan_b200/data.py) -- no GPU needed:

  * the LZ4 frame decoder against frames produced by the system liblz4 (the library behind the reference's `lz4framed`),
    linked and independent blocks, checksums, content size, stored blocks, multi-block inputs; malformed input is refused;
  * the LMDB reader against files laid out by the small writer below (py-lmdb / liblmdb are not installed in this image,
    so files written by liblmdb itself are NOT covered: parity with it is unpinned): both meta pages (the newer one
    wins), a branch level, inline and overflow values, missing keys;
  * PatchRNADataset end to end on two synthetic slides: same sampling, keys, BGR->RGB / CHW / transforms semantics as
    src/read_data.py:266-372, plus the raw uint8 path the device prefetcher consumes.
"""
import ctypes
import ctypes.util
import os
import pickle
import random
import struct

import numpy as np
import pytest
import torch

from rnagan_b200 import data as D

PSIZE = 4096


# ------------------------------------------------------------------------------------------------ LZ4 via the system lib
def _liblz4():
    name = ctypes.util.find_library("lz4")
    for cand in ([name] if name else []) + ["liblz4.so.1"]:
        try:
            return ctypes.CDLL(cand)
        except OSError:
            continue
    return None


class _FrameInfo(ctypes.Structure):
    _fields_ = [("blockSizeID", ctypes.c_int), ("blockMode", ctypes.c_int), ("contentChecksumFlag", ctypes.c_int),
                ("frameType", ctypes.c_int), ("contentSize", ctypes.c_ulonglong), ("dictID", ctypes.c_uint),
                ("blockChecksumFlag", ctypes.c_int)]


class _Prefs(ctypes.Structure):
    _fields_ = [("frameInfo", _FrameInfo), ("compressionLevel", ctypes.c_int), ("autoFlush", ctypes.c_uint),
                ("favorDecSpeed", ctypes.c_uint), ("reserved", ctypes.c_uint * 3)]


def _compress(lib, raw, prefs=None):
    lib.LZ4F_compressFrameBound.restype = ctypes.c_size_t
    lib.LZ4F_compressFrameBound.argtypes = [ctypes.c_size_t, ctypes.c_void_p]
    lib.LZ4F_compressFrame.restype = ctypes.c_size_t
    lib.LZ4F_compressFrame.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]
    p = ctypes.byref(prefs) if prefs is not None else None
    cap = lib.LZ4F_compressFrameBound(len(raw), p)
    buf = ctypes.create_string_buffer(cap)
    n = lib.LZ4F_compressFrame(buf, cap, raw, len(raw), p)
    assert n <= cap
    return buf.raw[:n]


def _payloads():
    rng = np.random.default_rng(0)
    tile = rng.integers(0, 256, size=(256, 256, 3), dtype=np.uint8)
    tile[64:192, :, :] = tile[64:65, :, :]                        # long repeats: matches that span 64 KiB blocks
    return [b"", b"a", b"abcabcabcabcabcabcabcabcabc" * 50, bytes(rng.integers(0, 4, size=300000, dtype=np.uint8)),
            tile.tobytes(), pickle.dumps(("name.png", tile.tobytes(), tile.shape))]


def test_lz4_frame_decoder_matches_system_liblz4():
    lib = _liblz4()
    if lib is None:
        pytest.skip("system liblz4 not present")
    variants = [None]
    for block_mode in (0, 1):                                     # 0 linked, 1 independent
        for size_id in (4, 7):                                    # 64 KiB / 4 MiB blocks
            for csum in (0, 1):
                p = _Prefs()
                p.frameInfo.blockSizeID, p.frameInfo.blockMode = size_id, block_mode
                p.frameInfo.contentChecksumFlag, p.frameInfo.blockChecksumFlag = csum, csum
                variants.append(p)
    for raw in _payloads():
        for prefs in variants:
            if prefs is not None:
                prefs.frameInfo.contentSize = len(raw) if len(raw) % 2 else 0
            frame = _compress(lib, raw, prefs)
            assert D.lz4f_decompress(frame) == raw


def test_lz4_frame_decoder_stored_blocks_and_errors():
    raw = bytes(range(256)) * 3
    # hand-built frame: version 01, independent blocks, no checksums, one STORED block (high bit of the size)
    frame = struct.pack("<IBBB", 0x184D2204, 0x60, 0x40, 0x00) + struct.pack("<I", 0x80000000 | len(raw)) + raw + b"\0" * 4
    assert D.lz4f_decompress(frame) == raw
    # one hand-coded compressed block: 4 literals "abcd", then a match of length 12 at offset 4, then 1 literal "e"
    block = bytes([0x48]) + b"abcd" + struct.pack("<H", 4) + bytes([0x10]) + b"e"
    frame = struct.pack("<IBBB", 0x184D2204, 0x60, 0x40, 0x00) + struct.pack("<I", len(block)) + block + b"\0" * 4
    assert D.lz4f_decompress(frame) == b"abcd" * 4 + b"e"
    for bad in (b"", b"\x04\x22\x4d\x18", b"not a frame at all....", frame[:-6]):
        with pytest.raises(ValueError):
            D.lz4f_decompress(bad)


# ------------------------------------------------------------------------------------------------ LMDB test-side writer
def _page(pgno, flags, ptrs_nodes=None, overflow_pages=None, payload=b""):
    """One page: header pgno(8) pad(2) flags(2) lower(2) upper(2) (or pages(4)), node pointers growing up, nodes packed
    from the end of the page down (even offsets)."""
    if overflow_pages is not None:
        body = struct.pack("<QHHI", pgno, 0, flags, overflow_pages) + payload
        return body + b"\0" * (overflow_pages * PSIZE - len(body))
    page = bytearray(PSIZE)
    upper = PSIZE
    ptrs = []
    for node in ptrs_nodes:
        node = node + b"\0" * (len(node) & 1)
        upper -= len(node)
        page[upper:upper + len(node)] = node
        ptrs.append(upper)
    lower = 16 + 2 * len(ptrs)
    assert lower <= upper
    page[0:16] = struct.pack("<QHHHH", pgno, 0, flags, lower, upper)
    for i, off in enumerate(ptrs):
        page[16 + 2 * i:18 + 2 * i] = struct.pack("<H", off)
    return bytes(page)


def _meta(pgno, txnid, root, entries, depth, last_pg):
    free_db = struct.pack("<IHHQQQQQ", PSIZE, 0, 0, 0, 0, 0, 0, 0xFFFFFFFFFFFFFFFF)
    main_db = struct.pack("<IHHQQQQQ", 0, 0, depth, 1, 1, 0, entries, root)
    body = struct.pack("<QHHHH", pgno, 0, 0x08, 0, 0) + struct.pack("<IIQQ", 0xBEEFC0DE, 1, 0, 1 << 30) + free_db + \
        main_db + struct.pack("<QQ", last_pg, txnid)
    return body + b"\0" * (PSIZE - len(body))


def write_lmdb(path, items, leaf_fill=5, inline_max=1000):
    """items: {key bytes: value bytes}.  Two-level tree (one branch root, `leaf_fill` keys per leaf), values longer than
    `inline_max` on overflow pages.  Meta page 0 is a stale empty snapshot (txn 1), meta page 1 the live one (txn 2)."""
    keys = sorted(items)
    pages = {}
    next_pg = [2]

    def alloc(n=1):
        p = next_pg[0]
        next_pg[0] += n
        return p

    leaves = []
    for i in range(0, len(keys), leaf_fill):
        chunk = keys[i:i + leaf_fill]
        pg = alloc()
        nodes = []
        for k in chunk:
            v = items[k]
            if len(v) > inline_max:
                npages = (16 + len(v) + PSIZE - 1) // PSIZE
                opg = alloc(npages)
                pages[opg] = _page(opg, 0x04, overflow_pages=npages, payload=v)
                nodes.append(struct.pack("<HHHH", len(v) & 0xFFFF, len(v) >> 16, 0x01, len(k)) + k + struct.pack("<Q", opg))
            else:
                nodes.append(struct.pack("<HHHH", len(v) & 0xFFFF, len(v) >> 16, 0, len(k)) + k + v)
        pages[pg] = _page(pg, 0x02, nodes)
        leaves.append((chunk[0], pg))
    root = alloc()
    bnodes = []
    for i, (first, pg) in enumerate(leaves):
        k = b"" if i == 0 else first
        bnodes.append(struct.pack("<HHHH", pg & 0xFFFF, (pg >> 16) & 0xFFFF, (pg >> 32) & 0xFFFF, len(k)) + k)
    pages[root] = _page(root, 0x01, bnodes)
    last = next_pg[0] - 1
    with open(path, "wb") as f:
        f.write(_meta(0, 1, 0xFFFFFFFFFFFFFFFF, 0, 0, 1))
        f.write(_meta(1, 2, root, len(keys), 2, last))
        pg = 2
        while pg <= last:
            blob = pages[pg]
            f.write(blob)
            pg += len(blob) // PSIZE


def test_lmdb_reader_tree_overflow_and_meta_selection(tmp_path):
    rng = np.random.default_rng(1)
    items = {f"k{i:04d}".encode(): bytes(rng.integers(0, 256, size=int(rng.integers(1, 900)), dtype=np.uint8))
             for i in range(57)}
    items[b"__keys__"] = b"x" * 5000                              # overflow value, two pages
    items[b"big"] = bytes(rng.integers(0, 256, size=200000, dtype=np.uint8))
    items[b""] = b"empty key"
    path = str(tmp_path / "slide.db")
    write_lmdb(path, items)
    with D.LMDBFile(path) as db:
        st = db.stat()
        assert st["entries"] == len(items) and st["psize"] == PSIZE and st["depth"] == 2
        for k, v in items.items():
            assert db.get(k) == v, k
        for missing in (b"k9999", b"a", b"k0000x", b"zzzz", b"k00"):
            assert db.get(missing) is None
            assert db.get(missing, b"dflt") == b"dflt"
    with pytest.raises(OSError):
        D.LMDBFile(str(tmp_path / "absent.db"))
    bad = tmp_path / "garbage.db"
    bad.write_bytes(b"\0" * 8192)
    with pytest.raises(OSError):
        D.LMDBFile(str(bad))


def test_patch_rna_dataset_matches_reference_semantics(tmp_path):
    """Two synthetic slides written the way the reference's tiling step stores them (src/read_data.py:318-336):
    `__keys__` -> lz4(pickle(list of keys)), key -> lz4(pickle((name, bytes, shape))), BGR uint8 tiles."""
    import pandas as pd
    lib = _liblz4()
    if lib is None:
        pytest.skip("system liblz4 not present (fixtures are compressed with it)")
    rng = np.random.default_rng(2)
    size, genes = 32, 6
    rows, truth = [], {}
    for s, n_tiles in enumerate((7, 4)):
        wsi = f"SLIDE-{s}.svs"
        os.makedirs(tmp_path / wsi)
        keys = [f"{wsi}_patch_{i}".encode() for i in range(n_tiles)]
        items = {b"__keys__": _compress(lib, pickle.dumps(keys))}
        for k in keys:
            tile = rng.integers(0, 256, size=(size, size, 3), dtype=np.uint8)
            truth[k] = tile
            items[k] = _compress(lib, pickle.dumps((k.decode() + ".png", tile.tobytes(), tile.shape)))
        write_lmdb(str(tmp_path / wsi / wsi.replace(".svs", ".db")), items, leaf_fill=3, inline_max=600)
        rows.append({"wsi_file_name": wsi, **{f"rna_G{j}": float(rng.normal()) for j in range(genes)}})
    rows.append({"wsi_file_name": "MISSING.svs", **{f"rna_G{j}": 0.0 for j in range(genes)}})   # skipped with a message
    csv = str(tmp_path / "table.csv")
    pd.DataFrame(rows).to_csv(csv, index=False)

    def halve(img):                                               # a stand-in for the reference's transforms pipeline
        return img.float() / 255.0

    random.seed(5)
    ds = D.PatchRNADataset(str(tmp_path), csv, size, transforms=halve, max_patches_total=5)
    # same sampling as the reference: random.sample(range(n), min(n, max)) per slide, in CSV order
    random.seed(5)
    want = [(0, i) for i in random.sample(list(range(7)), 5)] + [(1, i) for i in random.sample(list(range(4)), 4)]
    assert len(ds) == 9 and [(int(f.split("-")[1][0]), i) for f, i in zip(ds.filenames, ds.images)] == want
    for idx in range(len(ds)):
        item = ds[idx]
        tile = truth[ds.keys[idx]]
        assert set(item) == {"image", "rna_data", "labels"}
        ref = torch.from_numpy(np.ascontiguousarray(tile[..., ::-1])).permute(2, 0, 1).float() / 255.0
        assert torch.equal(item["image"], ref)                    # cv2.COLOR_BGR2RGB + permute(2,0,1) + transforms
        assert item["rna_data"].dtype == torch.float32 and item["rna_data"].shape == (genes,)
        assert float(item["labels"]) == 0.0
    random.seed(5)
    raw = D.PatchRNADataset(str(tmp_path), csv, size, max_patches_total=5, raw=True)
    assert torch.equal(raw[3]["image"], torch.from_numpy(truth[raw.keys[3]])) and raw[3]["image"].dtype == torch.uint8
    # a DataLoader over it (default collate) yields the batches the train_ops consume
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=4)))
    assert batch["image"].shape == (4, 3, size, size) and batch["rna_data"].shape == (4, genes)
