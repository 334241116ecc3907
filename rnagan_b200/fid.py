"""The reference's FID script (src/fid.py:33-235), host side.

The reference scores generated tiles with the Frechet distance between Gaussians fitted to Inception-v3 `Mixed_7c`
features.  Everything it defines is here under the same names: `PartialInceptionNetwork` (torchvision's architecture with
a forward hook; the pretrained ImageNet weights cannot be downloaded in this environment, so they are passed in as a
state_dict / file instead of `pretrained=True`), `preprocess_image[s]`, `get_activations`,
`calculate_activation_statistics`, `calculate_frechet_distance`, `calculate_fid`.  The distance uses a symmetric
formulation instead of scipy's general matrix square root: Tr sqrt(C1 C2) = sum_i sqrt(lambda_i(C1^1/2 C2 C1^1/2)), two
`eigh` calls on symmetric matrices -- real by construction, so there is no imaginary residue to discard and no
singular-product retry; it agrees with the reference formula to rounding (tests/test_host_cpu.py).  Evaluation only
(stock torch ops + float64 numpy, run once per evaluation): not part of the sm_100a hot path.
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


# ------------------------------------------------------------------------------------------------ feature extractor
# src/fid.py:33-98: torchvision's Inception-v3 (pretrained ImageNet weights), activations of `Mixed_7c` captured with a
# forward hook and average-pooled to [N, 2048].  The architecture comes from torchvision (a dependency of the reference
# too); its pretrained weights cannot be downloaded here, so the constructor takes a state_dict / file instead of
# `pretrained=True` and refuses to score with random weights unless asked to.  Evaluation only: stock torch ops, not part
# of the sm_100a hot path.
class PartialInceptionNetwork:
    """Same `forward` contract as the reference class: x [N, 3, 299, 299] float32 in [0, 1] -> [N, 2048] activations."""

    def __init__(self, weights=None, transform_input=True, allow_random_weights=False):
        import torch
        from torchvision.models import inception_v3
        self._torch = torch
        # pretrained=True in the reference builds the net with transform_input=True, aux_logits=True
        self.inception_network = inception_v3(weights=None, aux_logits=True, transform_input=True, init_weights=False)
        if weights is None and not allow_random_weights:
            raise ValueError("PartialInceptionNetwork needs the torchvision Inception-v3 ImageNet weights (a state_dict "
                             "or the path of `inception_v3_google-*.pth`): they cannot be downloaded in this "
                             "environment.  Pass allow_random_weights=True only to exercise the pipeline.")
        if weights is not None:
            state = torch.load(weights, map_location="cpu") if isinstance(weights, (str, bytes)) else weights
            self.inception_network.load_state_dict(state)
        self.inception_network.Mixed_7c.register_forward_hook(self.output_hook)
        self.transform_input = transform_input
        self.mixed_7c_output = None

    def output_hook(self, module, input, output):
        self.mixed_7c_output = output                      # N x 2048 x 8 x 8

    def to(self, device):
        self.inception_network.to(device)
        return self

    def eval(self):
        self.inception_network.eval()
        return self

    def __call__(self, x):
        torch = self._torch
        assert x.shape[1:] == (3, 299, 299), "Expected input shape to be: (N,3,299,299), but got {}".format(x.shape)
        x = x * 2 - 1                                      # src/fid.py:54
        with torch.no_grad():
            self.inception_network(x)
            act = torch.nn.functional.adaptive_avg_pool2d(self.mixed_7c_output, (1, 1))
        return act.view(x.shape[0], 2048)

    forward = __call__


def preprocess_image(im):
    """src/fid.py:165-187: [H, W, 3] float32 in [0, 1] or uint8 -> float32 tensor [3, 299, 299] in [0, 1]
    (cv2.resize, bilinear -- the reference's default interpolation)."""
    import cv2
    import torch
    assert im.ndim == 3 and im.shape[2] == 3
    if im.dtype == np.uint8:
        im = im.astype(np.float32) / 255
    im = cv2.resize(np.ascontiguousarray(im, dtype=np.float32), (299, 299))
    return torch.from_numpy(np.ascontiguousarray(np.rollaxis(im, axis=2)))


def preprocess_images(images, use_multiprocessing=False):
    """src/fid.py:190-216 ([N, H, W, 3] -> [N, 3, 299, 299]); a thread pool instead of a process pool (cv2.resize releases
    the GIL)."""
    import torch
    if use_multiprocessing:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor() as pool:
            out = list(pool.map(preprocess_image, images))
    else:
        out = [preprocess_image(im) for im in images]
    return torch.stack(out, dim=0)


def get_activations(images, batch_size, device="cuda:0", network=None):
    """src/fid.py:67-98: [N, 3, 299, 299] float32 -> numpy [N, 2048]."""
    assert images.shape[1:] == (3, 299, 299)
    net = (network if network is not None else PartialInceptionNetwork()).to(device).eval()
    out = np.zeros((images.shape[0], 2048), dtype=np.float32)
    for lo in range(0, images.shape[0], batch_size):
        out[lo:lo + batch_size] = net(images[lo:lo + batch_size].to(device)).detach().cpu().numpy()
    return out


def calculate_activation_statistics(images, batch_size, device="cuda:0", network=None):
    """src/fid.py:100-112."""
    act = get_activations(images, batch_size, device=device, network=network)
    return np.mean(act, axis=0), np.cov(act, rowvar=False)


calculate_frechet_distance = frechet_distance      # src/fid.py:115-163 (the eps retry is unnecessary here)


def calculate_fid(images1, images2, use_multiprocessing, batch_size, device="cuda:0", network=None):
    """src/fid.py:219-235: FID between two [N, H, W, 3] image stacks (float32 in [0, 1] or uint8)."""
    a = preprocess_images(images1, use_multiprocessing)
    b = preprocess_images(images2, use_multiprocessing)
    mu1, s1 = calculate_activation_statistics(a, batch_size, device=device, network=network)
    mu2, s2 = calculate_activation_statistics(b, batch_size, device=device, network=network)
    return calculate_frechet_distance(mu1, s1, mu2, s2)
