"""Shared helpers for the parity tests: golden loading, sampling indices (must match oracle/make_golden.py)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED_G, SEED_D, SEED_V, SEED_BATCH, SEED_RUN = 11, 12, 13, 14, 99
CONFIGS = {"mini32": (32, 8, 256, 2), "mini64": (64, 4, 192, 1),
           # BASELINE config-2 shapes (gan_run_lung.json): 256x256 tiles, 19198 genes
           "full256": (256, 8, 19198, 1)}


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name), allow_pickle=False))


def sample_idx(n, k=256):
    return np.unique(np.linspace(0, n - 1, num=min(k, n)).astype(np.int64))


def sample_of(t, k=256):
    v = t.detach().double().flatten().cpu()
    return v[torch.from_numpy(sample_idx(v.numel(), k))].numpy()


def stats_of(t):
    v = t.detach().double().flatten().cpu()
    return v.sum().item(), v.norm().item()


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cosine(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
