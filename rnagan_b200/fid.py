"""Evaluation math of the reference's FID script (src/fid.py:100-163), host side.

The reference scores generated tiles with the Frechet distance between Gaussians fitted to Inception-v3 pool features.
The feature extractor needs the pretrained torchvision weights (not reachable offline, out of scope); what is here is the
part that follows it: the activation statistics (src/fid.py:100-112) and the distance (src/fid.py:115-163) for features
from ANY extractor.  The distance uses a symmetric formulation instead of scipy's general matrix square root:
Tr sqrt(C1 C2) = sum_i sqrt(lambda_i(C1^1/2 C2 C1^1/2)), two `eigh` calls on symmetric matrices -- real by construction,
so there is no imaginary residue to discard and no singular-product retry; it agrees with the reference formula to
rounding (tests/test_host_cpu.py).  float64 numpy, run once per evaluation: nothing for the GPU here.
"""
import numpy as np


def activation_statistics(features):
    """[N, D] features -> (mean [D], covariance [D, D] with the unbiased 1/(N-1) normalisation of np.cov)."""
    f = np.asarray(features, dtype=np.float64)
    if f.ndim != 2 or f.shape[0] < 2:
        raise ValueError(f"need [N >= 2, D] features, got {f.shape}")
    mu = f.mean(axis=0)
    c = f - mu
    return mu, c.T @ c / (f.shape[0] - 1)


def _sym_sqrt(m):
    w, v = np.linalg.eigh((m + m.T) * 0.5)
    return (v * np.sqrt(np.clip(w, 0.0, None))) @ v.T


def frechet_distance(mu1, sigma1, mu2, sigma2):
    """d^2 = |mu1 - mu2|^2 + Tr(C1) + Tr(C2) - 2 Tr sqrt(C1 C2) for symmetric positive semi-definite C1, C2."""
    mu1, mu2 = np.atleast_1d(np.asarray(mu1, np.float64)), np.atleast_1d(np.asarray(mu2, np.float64))
    c1, c2 = np.atleast_2d(np.asarray(sigma1, np.float64)), np.atleast_2d(np.asarray(sigma2, np.float64))
    if mu1.shape != mu2.shape or c1.shape != c2.shape or c1.shape != (mu1.shape[0], mu1.shape[0]):
        raise ValueError("mean vectors / covariance matrices have different dimensions")
    r = _sym_sqrt(c1)
    lam = np.linalg.eigvalsh((r @ c2 @ r + (r @ c2 @ r).T) * 0.5)
    tr_covmean = np.sqrt(np.clip(lam, 0.0, None)).sum()
    d = mu1 - mu2
    return float(d @ d + np.trace(c1) + np.trace(c2) - 2.0 * tr_covmean)


def fid_from_features(real_features, generated_features):
    return frechet_distance(*activation_statistics(generated_features), *activation_statistics(real_features))
