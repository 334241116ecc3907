"""Tensor-level wrappers over the C ABI: torch supplies device memory and streams, nothing else.

Every function takes CUDA tensors, checks dtype/contiguity, and launches asynchronously on the current stream.
"""
import torch

from . import _lib

BF16 = torch.bfloat16


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, dtype, name, rows_strided=False):
    if t.device.type != "cuda":
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if rows_strided:
        if t.dim() != 2 or t.stride(1) != 1:
            raise ValueError(f"{name} must be a 2-D tensor with unit column stride")
    elif not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


_ws_cache = {}

# When set to a list, every tcgen05 contraction records (kind, algorithmic flops, start event, end event): bench.py
# uses it to measure per-kernel roofline fractions live, on the launching stream, outside the timed region.
PROFILE = None


def _prof(kind, flops, fn):
    if PROFILE is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    PROFILE.append((kind, float(flops), e0, e1))
    return r


def _workspace(nbytes, device):
    """Grow-only fp32 scratch per device (split-K partials); caller-owned from the library's point of view."""
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    cur = _ws_cache.get(key)
    if cur is None or cur.numel() * 4 < nbytes:
        cur = torch.empty((max(nbytes, 1) + 3) // 4, dtype=torch.float32, device=device)
        _ws_cache[key] = cur
    return cur


def up_pad(cs):
    return max(16, (cs + 15) // 16 * 16)


# ------------------------------------------------------------------------------------------------ packing
def pack_link(W, w_down=None, w_up=None, want_down=True, want_up=True):
    """W: fp32 [Cp, Cs, 4, 4] -> (w_down bf16 [Cp, 16*Cs], w_up bf16 [4, Cs_pad, 4*Cp])."""
    _chk(W, torch.float32, "W")
    Cp, Cs = W.shape[0], W.shape[1]
    if want_down and w_down is None:
        w_down = torch.empty(Cp, 16 * Cs, dtype=BF16, device=W.device)
    if want_up and w_up is None:
        w_up = torch.zeros(4, up_pad(Cs), 4 * Cp, dtype=BF16, device=W.device)
    _lib.check(_lib.lib().rg_pack_link(_p(W), _p(w_down) if want_down else None, _p(w_up) if want_up else None,
                                       Cp, Cs, _st()), "rg_pack_link")
    return w_down, w_up


def is_native4(t):
    """True when a 4-D tensor [N, C, H, W] is stored channels_last, i.e. physically [N][H][W][C] -- the engine's
    native weight layout (same element order as the packed w_down operand and as the wgrad accumulator rows)."""
    return t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def phys2d(W):
    """[Cp, 16*Cs] view of the physical memory of a native (channels_last) weight [Cp, Cs, 4, 4]."""
    return W.permute(0, 2, 3, 1).reshape(W.shape[0], -1)


def _wgrad_layout(dW):
    if dW.dtype != torch.float32 or dW.device.type != "cuda":
        raise ValueError("dW must be a CUDA fp32 tensor")
    if dW.is_contiguous():
        return 0
    if is_native4(dW):
        return 1
    raise ValueError("dW must be contiguous or channels_last")


def pack_up_from_down(w_down, w_up, Cs):
    """bf16 w_down [Cp, 16*Cs] -> bf16 w_up [4, Cs_pad, 4*Cp]."""
    _chk(w_down, BF16, "w_down"); _chk(w_up, BF16, "w_up")
    _lib.check(_lib.lib().rg_pack_up_from_down(_p(w_down), _p(w_up), w_down.shape[0], Cs, _st()),
               "rg_pack_up_from_down")
    return w_up


def pack_up9_from_down(w_down, Cs, out=None):
    """bf16 w_down [Cp, 16*Cs] (Cs == 64) -> merged-phase operand bf16 [9, Cp/64, 2, 4, 32, 64] of conv_up."""
    _chk(w_down, BF16, "w_down")
    Cp = w_down.shape[0]
    if out is None:
        out = torch.empty(9, Cp // 64, 2, 4, 32, 64, dtype=BF16, device=w_down.device)
    _lib.check(_lib.lib().rg_pack_up9_from_down(_p(w_down), _p(out), Cp, Cs, _st()), "rg_pack_up9_from_down")
    return out


def pack_proj(W, out=None):
    """W: fp32 [E, C0, 4, 4] -> bf16 [16*C0, E]."""
    _chk(W, torch.float32, "W")
    E, C0 = W.shape[0], W.shape[1]
    if out is None:
        out = torch.empty(16 * C0, E, dtype=BF16, device=W.device)
    _lib.check(_lib.lib().rg_pack_proj(_p(W), _p(out), E, C0, _st()), "rg_pack_proj")
    return out


def pack_edge(W, out=None):
    """W: fp32 [Cp, Cimg, 4, 4] -> bf16 [Cp, 64] (k = tap*4 + c)."""
    _chk(W, torch.float32, "W")
    Cp, Cimg = W.shape[0], W.shape[1]
    if out is None:
        out = torch.empty(Cp, 64, dtype=BF16, device=W.device)
    _lib.check(_lib.lib().rg_pack_edge(_p(W), _p(out), Cp, Cimg, _st()), "rg_pack_edge")
    return out


def cast_pad_bf16(src, cols_pad=None, out=None):
    _chk(src, torch.float32, "src")
    rows, cols = src.shape
    cols_pad = cols if cols_pad is None else cols_pad
    if out is None:
        out = torch.empty(rows, cols_pad, dtype=BF16, device=src.device)
    _lib.check(_lib.lib().rg_cast_pad_bf16(_p(src), _p(out), rows, cols, cols_pad, _st()), "rg_cast_pad_bf16")
    return out


# ------------------------------------------------------------------------------------------------ contractions
_stats_cache = {}


def stats_ws(C, device, slot=0):
    """fp32 [parts, 2, C] scratch a convolution epilogue fills with per-CTA channel sums / sums of squares."""
    key = (device.index, C, slot)
    t = _stats_cache.get(key)
    if t is None:
        t = torch.zeros(_lib.lib().rg_stats_parts(), 2, C, dtype=torch.float32, device=device)
        _stats_cache[key] = t
    return t


def _aux(aux):
    """aux: None, ("lrelu", h, slope) or ("bn", a, mean, rstd, scale, shift, slope) -> (ctypes pointer | None, keepalive)."""
    if aux is None:
        return None, None
    import ctypes
    if aux[0] == "lrelu":
        st = _lib.EpilogueAux(aux[1].data_ptr(), 1, None, None, None, None, float(aux[2]))
    else:
        _, a, mean, rstd, scale, shift, slope = aux
        st = _lib.EpilogueAux(a.data_ptr(), 2, mean.data_ptr(), rstd.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                              float(slope))
    return ctypes.addressof(st), st


def reduce_partials(stats, out):
    """out[k, c] = sum over the per-CTA rows of stats [parts, K, C] (fixed order)."""
    _lib.check(_lib.lib().rg_reduce_partials(_p(stats), out.numel(), _p(out), _st()), "rg_reduce_partials")


def gemm_nt_bwd(A, Bw, out, stats=None, aux=None):
    """bf16 out[M, N] = A[M, K] @ Bw[N, K]^T with fused statistics / elementwise backward (see rg_epilogue_aux)."""
    _chk(A, BF16, "A", True); _chk(Bw, BF16, "Bw", True)
    M, K = A.shape
    N = Bw.shape[0]
    ap, keep = _aux(aux)
    _prof("gemm_nt", 2.0 * M * N * K, lambda: _lib.check(
        _lib.lib().rg_gemm_nt_bwd(_p(A), A.stride(0), _p(Bw), Bw.stride(0), _p(out), M, N, K, out.stride(0), _p(stats),
                                  ap, _st()), "rg_gemm_nt_bwd"))
    return out


def conv_down(hi, w_down, out=None, stats=None, aux=None):
    """hi bf16 [B, 2H, 2W, Cs], w_down bf16 [Cp, 16*Cs] -> lo bf16 [B, H, W, Cp] (+ fused BN statistics)."""
    _chk(hi, BF16, "hi"); _chk(w_down, BF16, "w_down")
    B, H2, W2, Cs = hi.shape
    Cp = w_down.shape[0]
    H, W = H2 // 2, W2 // 2
    if out is None:
        out = torch.empty(B, H, W, Cp, dtype=BF16, device=hi.device)
    ap, keep = _aux(aux)
    _prof("conv_down", 2.0 * B * H * W * Cp * 16 * Cs, lambda: _lib.check(
        _lib.lib().rg_conv_down(_p(hi), _p(w_down), _p(out), B, H, W, Cs, Cp, _p(stats), ap, _st()), "rg_conv_down"))
    return out


def conv_up(lo, w, Cs, out=None, stats=None, aux=None):
    """lo bf16 [B, H, W, Cp] -> hi bf16 [B, 2H, 2W, Cs].  w: w_down bf16 [Cp, 16*Cs] (2-D; read MN-major),
    w_up bf16 [4, Cs_pad, 4*Cp] (3-D; K-major) or the merged-phase w_up9 (6-D; Cs == 64, >= 256 low-res pixels)."""
    _chk(lo, BF16, "lo"); _chk(w, BF16, "w")
    B, H, W, Cp = lo.shape
    if out is None:
        out = torch.empty(B, 2 * H, 2 * W, Cs, dtype=BF16, device=lo.device)
    ap, keep = _aux(aux)
    _prof("conv_up", 2.0 * B * H * W * Cp * 16 * Cs, lambda: _lib.check(
        _lib.lib().rg_conv_up(_p(lo), _p(w), {2: 1, 3: 0, 6: 2}[w.dim()], _p(out), B, H, W, Cp, Cs, _p(stats), ap,
                              _st()),
        "rg_conv_up"))
    return out


def conv_up_img(lo, w_up, Cimg, bias=None, act_tanh=False, out=None):
    """lo bf16 [B, H, W, Cp] -> fp32 NCHW image [B, Cimg, 2H, 2W] (+bias, tanh)."""
    _chk(lo, BF16, "lo"); _chk(w_up, BF16, "w_up")
    B, H, W, Cp = lo.shape
    if out is None:
        out = torch.empty(B, Cimg, 2 * H, 2 * W, dtype=torch.float32, device=lo.device)
    _prof("conv_up_img", 2.0 * B * H * W * Cp * 16 * Cimg, lambda: _lib.check(
        _lib.lib().rg_conv_up_img(_p(lo), _p(w_up), _p(out), _p(bias), int(act_tanh), B, H, W, Cp, Cimg, _st()),
        "rg_conv_up_img"))
    return out


def conv_wgrad(lo, hi, dW, alpha=1.0, alpha_dev=None, beta=0.0):
    """dW fp32 [Cp, Cs, 4, 4] (contiguous or channels_last) = beta*dW + alpha * sum lo (x) hi@tap."""
    _chk(lo, BF16, "lo"); _chk(hi, BF16, "hi")
    native = _wgrad_layout(dW)
    B, H, W, Cp = lo.shape
    Cs = hi.shape[3]
    L = _lib.lib()
    nbytes = L.rg_conv_wgrad_ws_bytes(B, H, W, Cp, Cs)
    ws = _workspace(nbytes, lo.device)
    _prof("conv_wgrad", 2.0 * B * H * W * Cp * 16 * Cs, lambda: _lib.check(
        L.rg_conv_wgrad(_p(lo), _p(hi), _p(dW), _p(ws), ws.numel() * 4, B, H, W, Cp, Cs, float(alpha), _p(alpha_dev),
                        float(beta), native, _st()), "rg_conv_wgrad"))
    return dW


def proj_wgrad(z, da0, dW, alpha=1.0, alpha_dev=None, beta=0.0):
    """dW fp32 [E, C0, 4, 4] (contiguous or channels_last) = sum_b z[b, e] * da0[b, kh, kw, c]."""
    _chk(z, BF16, "z"); _chk(da0, BF16, "da0")
    native = _wgrad_layout(dW)
    B, E = z.shape
    C0 = da0.shape[-1]
    L = _lib.lib()
    nbytes = L.rg_proj_wgrad_ws_bytes(B, E, C0)
    ws = _workspace(nbytes, z.device)
    _prof("proj_wgrad", 2.0 * B * E * 16 * C0, lambda: _lib.check(
        L.rg_proj_wgrad(_p(z), _p(da0), _p(dW), _p(ws), ws.numel() * 4, B, E, C0, float(alpha), _p(alpha_dev),
                        float(beta), native, _st()), "rg_proj_wgrad"))
    return dW


def gemm_nt(A, Bw, out=None, col_scale=None, col_shift=None, slope=1.0, out_f32=False, N=None, K=None, tanh=False):
    """C[M, N] = lrelu((A[M, K] @ Bw[N, K]^T) * col_scale + col_shift); rows of A / Bw / out may be strided."""
    _chk(A, BF16, "A", True); _chk(Bw, BF16, "Bw", True)
    M = A.shape[0]
    K = A.shape[1] if K is None else K
    N = Bw.shape[0] if N is None else N
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32 if out_f32 else BF16, device=A.device)
    _prof("gemm_nt", 2.0 * M * N * K, lambda: _lib.check(
        _lib.lib().rg_gemm_nt_ld(_p(A), A.stride(0), _p(Bw), Bw.stride(0), _p(out), M, N, K, out.stride(0),
                                 _p(col_scale), _p(col_shift), float(slope),
                                 int(out.dtype == torch.float32) | (2 if tanh else 0), _st()),
        "rg_gemm_nt_ld"))
    return out


def gemm_nn(A, Bw, out=None, col_scale=None, col_shift=None, slope=1.0, out_f32=False, N=None, K=None):
    """C[M, N] = A[M, K] @ Bw[K, N] with Bw row-major (the input gradient of an nn.Linear with weight [K, N])."""
    _chk(A, BF16, "A", True); _chk(Bw, BF16, "Bw", True)
    M = A.shape[0]
    K = A.shape[1] if K is None else K
    N = Bw.shape[1] if N is None else N
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32 if out_f32 else BF16, device=A.device)
    _prof("gemm_nn", 2.0 * M * N * K, lambda: _lib.check(
        _lib.lib().rg_gemm_nn(_p(A), A.stride(0), _p(Bw), Bw.stride(0), _p(out), M, N, K, out.stride(0),
                              _p(col_scale), _p(col_shift), float(slope), int(out.dtype == torch.float32), _st()),
        "rg_gemm_nn"))
    return out


def gemm_tn(A, Bm, out=None, alpha=1.0, alpha_dev=None, beta=0.0, M=None, N=None):
    """C[M, N] fp32 (dense) = beta*C + alpha * A[R, M]^T @ Bm[R, N]; rows of A / Bm may be strided."""
    _chk(A, BF16, "A", True); _chk(Bm, BF16, "Bm", True)
    R = A.shape[0]
    M = A.shape[1] if M is None else M
    N = Bm.shape[1] if N is None else N
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    L = _lib.lib()
    nbytes = L.rg_gemm_tn_ws_bytes(R, M, N)
    ws = _workspace(nbytes, A.device)
    _prof("gemm_tn", 2.0 * R * M * N, lambda: _lib.check(
        L.rg_gemm_tn_ld(_p(A), A.stride(0), _p(Bm), Bm.stride(0), _p(out), _p(ws), ws.numel() * 4, R, M, N,
                        float(alpha), _p(alpha_dev), float(beta), _st()), "rg_gemm_tn_ld"))
    return out


# ------------------------------------------------------------------------------------------------ HBM-bound ops
F32 = torch.float32


def _red_ws(M, C, device):
    return _workspace(_lib.lib().rg_reduce_ws_bytes(M, C), device)


def bn_stats(a, M, C, sums):
    ws = _red_ws(M, C, a.device)
    _lib.check(_lib.lib().rg_bn_stats(_p(a), M, C, _p(ws), ws.numel() * 4, _p(sums), _st()), "rg_bn_stats")


def bn_finalize(sums, gamma, beta, M, C, eps, momentum, rmean, rvar, nbt, mean, rstd, scale, shift):
    _lib.check(_lib.lib().rg_bn_finalize(_p(sums), _p(gamma), _p(beta), M, C, float(eps), float(momentum), _p(rmean),
                                         _p(rvar), _p(nbt), _p(mean), _p(rstd), _p(scale), _p(shift), _st()),
               "rg_bn_finalize")


def bn_finalize_partials(stats, gamma, beta, M, C, eps, momentum, rmean, rvar, nbt, sums, mean, rstd, scale, shift):
    _lib.check(_lib.lib().rg_bn_finalize_partials(_p(stats), _p(gamma), _p(beta), M, C, float(eps), float(momentum),
                                                  _p(rmean), _p(rvar), _p(nbt), _p(sums), _p(mean), _p(rstd),
                                                  _p(scale), _p(shift), _st()), "rg_bn_finalize_partials")


def bn_act(a, scale, shift, slope, h, M, C):
    _lib.check(_lib.lib().rg_bn_act(_p(a), _p(scale), _p(shift), float(slope), _p(h), M, C, _st()), "rg_bn_act")


def bn_bwd_reduce(dh, a, mean, rstd, scale, shift, slope, M, C, sums):
    ws = _red_ws(M, C, a.device)
    _lib.check(_lib.lib().rg_bn_bwd_reduce(_p(dh), _p(a), _p(mean), _p(rstd), _p(scale), _p(shift), float(slope), M, C,
                                           _p(ws), ws.numel() * 4, _p(sums), _st()), "rg_bn_bwd_reduce")


def bn_bwd_apply(dh, a, add, mean, rstd, scale, shift, slope, sums, M, C, da, du_out=None):
    _lib.check(_lib.lib().rg_bn_bwd_apply(_p(dh), _p(a), _p(add), _p(mean), _p(rstd), _p(scale), _p(shift),
                                          float(slope), _p(sums), M, C, _p(da), _p(du_out), _st()), "rg_bn_bwd_apply")


def bn_param_grads(sums, dgamma, dbeta, C, acc_gamma, acc_beta):
    _lib.check(_lib.lib().rg_bn_param_grads(_p(sums), _p(dgamma), _p(dbeta), C, float(acc_gamma), float(acc_beta),
                                            _st()), "rg_bn_param_grads")


def lrelu_bwd(dh, h, slope, da, M, C):
    _lib.check(_lib.lib().rg_lrelu_bwd(_p(dh), _p(h), float(slope), _p(da), M, C, _st()), "rg_lrelu_bwd")


def col_sum(x, M, C, tmp, out, acc):
    ws = _red_ws(M, C, x.device)
    _lib.check(_lib.lib().rg_col_sum(_p(x), M, C, _p(ws), ws.numel() * 4, _p(tmp), _p(out), float(acc), _st()),
               "rg_col_sum")


def bn_gp_reduce(ggI, a, gO, mean, rstd, M, C, q):
    ws = _red_ws(M, C, a.device)
    _lib.check(_lib.lib().rg_bn_gp_reduce(_p(ggI), _p(a), _p(gO), _p(mean), _p(rstd), M, C, _p(ws), ws.numel() * 4,
                                          _p(q), _st()), "rg_bn_gp_reduce")


def bn_gp_apply(ggI, a, gO, mean, rstd, gamma, scale, shift, slope, s, q, M, C, A_dh, A_a, dgamma, dgamma_acc):
    _lib.check(_lib.lib().rg_bn_gp_apply(_p(ggI), _p(a), _p(gO), _p(mean), _p(rstd), _p(gamma), _p(scale), _p(shift),
                                         float(slope), _p(s), _p(q), M, C, _p(A_dh), _p(A_a), _p(dgamma),
                                         float(dgamma_acc), _st()), "rg_bn_gp_apply")


def latent_prep(noise, z, lat_bf16=None, lat_f32=None):
    _chk(noise, F32, "noise"); _chk(z, F32, "z")
    B, E = noise.shape
    _lib.check(_lib.lib().rg_latent_prep(_p(noise), _p(z), B, E, z.shape[0], _p(lat_bf16), _p(lat_f32), _st()),
               "rg_latent_prep")


def im2col_img(x, col, y=None, mode=0, eps_dev=None, mul_dev=None, mixed_out=None):
    _chk(x, F32, "x")
    B, Cimg, S, _ = x.shape
    _lib.check(_lib.lib().rg_im2col_img(_p(x), _p(y), mode, _p(eps_dev), _p(mul_dev), B, Cimg, S, _p(col),
                                        _p(mixed_out), _st()), "rg_im2col_img")


_partial_cache = {}


def img_channel_sum(x, out, y=None, mode=0, acc=0.0):
    B, Cimg, S, _ = x.shape
    part = _partial_cache.get(x.device.index)
    if part is None:
        part = torch.empty(1024, dtype=F32, device=x.device)
        _partial_cache[x.device.index] = part
    _lib.check(_lib.lib().rg_img_channel_sum(_p(x), _p(y), mode, B, Cimg, S, _p(part), part.numel(), _p(out),
                                             float(acc), _st()), "rg_img_channel_sum")


def pack_edge_t(W, out):
    """W fp32 [Cp, Cimg, 4, 4] -> bf16 [rows, Cp] (row = tap*Cimg + c), B operand of the dgrad-form image GEMM."""
    Cp, Cimg = W.shape[0], W.shape[1]
    _lib.check(_lib.lib().rg_pack_edge_t(_p(W), _p(out), Cp, Cimg, out.shape[0], _st()), "rg_pack_edge_t")
    return out


def conv_up_img_col(lo, w_colT, Cimg, col, out, bias=None, act_tanh=False, unit_nhwc=False, u8=False, bgr=False):
    """Image-side transposed conv as GEMM (K = Cp only, each input pixel read once) + col2im:
    lo bf16 [B, H, W, Cp] -> out fp32 NCHW [B, Cimg, 2H, 2W]; col: fp32 scratch [B*H*W, 16*Cimg].
    unit_nhwc: out is instead the synthesis result (x + 1) / 2 as fp32 NHWC [B, 2H, 2W, Cimg];
    u8: out is the uint8 NHWC tile trunc(255 * (x + 1) / 2) (bgr: channel order reversed for cv2.imwrite)."""
    B, H, W, Cp = lo.shape
    N = 16 * Cimg
    gemm_nt(lo.view(B * H * W, Cp), w_colT, out=col, N=N)
    _lib.check(_lib.lib().rg_col2im_img(_p(col), col.stride(0), _p(bias),
                                        int(act_tanh) | (2 if unit_nhwc else 0) | (4 if u8 else 0) | (8 if bgr else 0), B,
                                        Cimg, H, W, _p(out), _st()), "rg_col2im_img")
    return out


# ---- fused image-side convolutions (csrc/rg_img.cu): need 64 channels on the wide side
def img_conv_up_pack(W, out=None):
    """W fp32 [64, Cimg, 4, 4] -> the packed mma.sync B-fragment table of img_conv_up (uint8 buffer)."""
    _chk(W, F32, "W")
    L = _lib.lib()
    if out is None:
        out = torch.empty(L.rg_img_conv_up_pack_bytes(), dtype=torch.uint8, device=W.device)
    _lib.check(L.rg_img_conv_up_pack(_p(W), W.shape[0], W.shape[1], _p(out), _st()), "rg_img_conv_up_pack")
    return out


def img_conv_up(lo, W, out, bias=None, act_tanh=False, unit_nhwc=False, u8=False, bgr=False, Cimg=None, bn=None):
    """lo bf16 [B, H, W, 64]; W: fp32 [64, Cimg, 4, 4] (packed on the fly) or the table of img_conv_up_pack (then pass
    Cimg) -> out: fp32 NCHW [B, Cimg, 2H, 2W] (default), fp32 NHWC (x+1)/2 (unit_nhwc) or uint8 NHWC trunc(255 (x+1)/2)
    (u8).  ConvTranspose2d(64, Cimg, 4, 2, 1) + bias + tanh.  bn = (scale, shift, slope): `lo` is the pre-BatchNorm
    activation and lrelu(scale * lo + shift) is applied on the fly."""
    _chk(lo, BF16, "lo")
    B, H, Wd, Cp = lo.shape
    if W.dtype == F32:
        Cimg = W.shape[1]
        W = img_conv_up_pack(W)
    flags = int(act_tanh) | (2 if unit_nhwc else 0) | (4 if u8 else 0) | (8 if bgr else 0)
    bn_scale, bn_shift, bn_slope = bn if bn is not None else (None, None, 1.0)      # lo = pre-BatchNorm activation
    _prof("img_conv_up", 2.0 * B * H * Wd * Cp * 16 * Cimg, lambda: _lib.check(
        _lib.lib().rg_img_conv_up(_p(lo), _p(W), _p(bias), flags, B, H, Wd, Cp, Cimg, _p(out), _p(bn_scale), _p(bn_shift),
                                  float(bn_slope), _st()), "rg_img_conv_up"))
    return out


def img_conv_down(x, W, out, y=None, mode=0, eps_dev=None, mul_dev=None, bias=None, slope=1.0, mask_src=None,
                  mask_slope=1.0):
    """x fp32 NCHW [B, Cimg, S, S] (transformed per `mode`, see rg_img_conv_down), W fp32 [64, Cimg, 4, 4] ->
    out bf16 [B, S/2, S/2, 64] = lrelu(conv2d(x', W, stride 2, pad 1) + bias) (* LeakyReLU' mask of mask_src)."""
    _chk(x, F32, "x"); _chk(W, F32, "W"); _chk(out, BF16, "out")
    B, Cimg, S, _ = x.shape
    Cp = W.shape[0]
    _prof("img_conv_down", 2.0 * B * (S // 2) ** 2 * Cp * 16 * Cimg, lambda: _lib.check(
        _lib.lib().rg_img_conv_down(_p(x), _p(y), mode, _p(eps_dev), _p(mul_dev), _p(W), _p(bias), float(slope),
                                    _p(mask_src), float(mask_slope), B, Cimg, S, Cp, _p(out), _st()),
        "rg_img_conv_down"))
    return out


def img_conv_wgrad(act, x, dW, y=None, mode=0, eps_dev=None, mul_dev=None, acc=0.0, dbias=None, acc_bias=0.0):
    """dW fp32 [64, Cimg, 4, 4] (contiguous) = acc*dW + sum act (x) x'@tap; optional dbias[64] = acc_bias*dbias + sum act."""
    _chk(act, BF16, "act"); _chk(x, F32, "x"); _chk(dW, F32, "dW")
    B, Cimg, S, _ = x.shape
    L = _lib.lib()
    ws = _workspace(L.rg_img_conv_wgrad_ws_bytes(), x.device)
    _prof("img_conv_wgrad", 2.0 * B * (S // 2) ** 2 * 64 * 16 * Cimg, lambda: _lib.check(
        L.rg_img_conv_wgrad(_p(act), _p(x), _p(y), mode, _p(eps_dev), _p(mul_dev), B, Cimg, S, act.shape[-1], _p(ws),
                            ws.numel() * 4, _p(dW), float(acc), _p(dbias), float(acc_bias), _st()), "rg_img_conv_wgrad"))
    return dW


def unpack_edge_grad(dcol, dW, acc=0.0):
    Cp, Cimg = dW.shape[0], dW.shape[1]
    _lib.check(_lib.lib().rg_unpack_edge_grad(_p(dcol), _p(dW), Cp, Cimg, float(acc), _st()), "rg_unpack_edge_grad")


def pack_head(W, w_head):
    _lib.check(_lib.lib().rg_pack_head(_p(W), _p(w_head), W.shape[1], _st()), "rg_pack_head")


def head_fwd(h5, w_head, B, K, slope, a6, out):
    _lib.check(_lib.lib().rg_head_fwd(_p(h5), _p(w_head), B, K, float(slope), _p(a6), _p(out), _st()), "rg_head_fwd")


def head_bwd_data(a6, dout_const, w_head, B, K, slope, da6, dh5, dout=None):
    _lib.check(_lib.lib().rg_head_bwd_data(_p(a6), _p(dout), float(dout_const), _p(w_head), B, K, float(slope),
                                           _p(da6), _p(dh5), _st()), "rg_head_bwd_data")


def head_wgrad(da6, x, B, K, C, dW, acc):
    _lib.check(_lib.lib().rg_head_wgrad(_p(da6), _p(x), B, K, C, _p(dW), float(acc), _st()), "rg_head_wgrad")


def wgan_loss(a, sign_a, loss_out, b=None, sign_b=0.0):
    _lib.check(_lib.lib().rg_wgan_loss(_p(a), float(sign_a), _p(b), float(sign_b), a.numel(), _p(loss_out), _st()),
               "rg_wgan_loss")


def gp_norm(g, lambd, partial, out3):
    _lib.check(_lib.lib().rg_gp_norm(_p(g), g.numel(), float(lambd), _p(partial), partial.numel(), _p(out3), _st()),
               "rg_gp_norm")


def tiles_u8_to_nchw(tiles, out=None, swap_rb=False):
    """uint8 [B, S, S, C] (HWC, optionally BGR) on the device -> fp32 NCHW [B, C, S, S] in [-1, 1]."""
    if tiles.device.type != "cuda" or tiles.dtype != torch.uint8 or not tiles.is_contiguous() or tiles.dim() != 4:
        raise ValueError("tiles must be a contiguous CUDA uint8 tensor [B, S, S, C] (no CPU fallback)")
    B, S, S2, C = tiles.shape
    if S != S2:
        raise ValueError("tiles must be square")
    if out is None:
        out = torch.empty(B, C, S, S, dtype=torch.float32, device=tiles.device)
    _lib.check(_lib.lib().rg_tiles_u8_to_nchw(_p(tiles), _p(out), B, C, S, int(swap_rb), _st()), "rg_tiles_u8_to_nchw")
    return out


def tiles_to_unit_nhwc(img, out):
    B, C, S, _ = img.shape
    _lib.check(_lib.lib().rg_tiles_to_unit_nhwc(_p(img), _p(out), B, C, S, _st()), "rg_tiles_to_unit_nhwc")


def clamp_(p, lo, hi):
    _lib.check(_lib.lib().rg_clamp(_p(p), p.numel(), float(lo), float(hi), _st()), "rg_clamp")


def slices_sum(stage, nparts, n, out):
    """out[:n] = sum_r stage[r, :n] in ascending r (stage: fp32 [nparts, stride] contiguous; n % 4 == 0)."""
    _lib.check(_lib.lib().rg_slices_sum(_p(stage), int(nparts), int(stage.stride(0)), int(n), _p(out), _st()),
               "rg_slices_sum")
    return out


def nvls_allreduce(mc_ptr, offset, n, max_ctas=0):
    """In-switch SUM of floats [offset, offset+n) of a symmetric buffer through its multicast address (int)."""
    _lib.check(_lib.lib().rg_nvls_allreduce(int(mc_ptr), int(offset), int(n), int(max_ctas), _st()), "rg_nvls_allreduce")


class AdamTable:
    """Chunk table over a fixed list of (param, grad, exp_avg, exp_avg_sq) fp32 tensors for rg_adam_step."""

    CHUNK = 1 << 16

    def __init__(self, params, grads, ms, vs, shadows=None):
        import ctypes
        L = _lib.lib()
        n = len(params)
        shadows = list(shadows) if shadows is not None else [None] * n
        sizes = [p.numel() for p in params]
        max_chunks = sum((s + self.CHUNK - 1) // self.CHUNK for s in sizes)
        arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
        host = torch.empty(L.rg_adam_table_bytes(max_chunks), dtype=torch.uint8).pin_memory()
        sh = (ctypes.c_void_p * n)(*[None if t is None else t.data_ptr() for t in shadows])
        # a shadow with more columns than its 2-D parameter is a pitched copy (zero-padded K): Adam writes it row by row
        cols = [p.shape[1] if (s_ is not None and p.dim() == 2 and s_.dim() == 2 and s_.shape[0] == p.shape[0]
                               and s_.shape[1] > p.shape[1]) else 0 for p, s_ in zip(params, shadows)]
        pitch = [s_.shape[1] if c else 0 for c, s_ in zip(cols, shadows)]
        for p, s_, c in zip(params, shadows, cols):
            if s_ is not None and not c and s_.numel() != p.numel():
                raise ValueError("AdamTable: a shadow must have the parameter's element count or be a row-padded 2-D copy")
        nch = L.rg_adam_build_table_pitched(arr(params), arr(grads), arr(ms), arr(vs), sh, (ctypes.c_int * n)(*cols),
                                            (ctypes.c_int * n)(*pitch), (ctypes.c_int64 * n)(*sizes), n, self.CHUNK,
                                            host.data_ptr(), max_chunks)
        if nch <= 0:
            _lib.check(nch if nch < 0 else -1, "rg_adam_build_table")
        self.num_chunks = nch
        self.table = host.to(params[0].device, non_blocking=False)
        self.keep = (params, grads, ms, vs)
        self.shadows = shadows
        self.ptrs = tuple(t.data_ptr() for ts in self.keep for t in ts)

    def step_dyn(self, dyn, beta1, beta2, eps, clamp=None, grad_scale=1.0):
        """The step with {lr, step count} in the device tensor `dyn` (fp32 [2]); the kernel advances the count itself."""
        lo, hi = (clamp if clamp is not None else (0.0, 0.0))
        _lib.check(_lib.lib().rg_adam_step_dyn(_p(self.table), self.num_chunks, _p(dyn), float(beta1), float(beta2),
                                               float(eps), int(clamp is not None), float(lo), float(hi),
                                               float(grad_scale), _st()), "rg_adam_step_dyn")

    def step(self, lr, beta1, beta2, eps, step, clamp=None, grad_scale=1.0):
        lo, hi = (clamp if clamp is not None else (0.0, 0.0))
        _lib.check(_lib.lib().rg_adam_step(_p(self.table), self.num_chunks, float(lr), float(beta1), float(beta2),
                                           float(eps), int(step), int(clamp is not None), float(lo), float(hi),
                                           float(grad_scale), _st()), "rg_adam_step")


# ------------------------------------------------------------------------------------------------ betaVAE training
def mul_cast_pad_bf16(src, out, mul=None, scale=1.0):
    rows, cols = src.shape
    _lib.check(_lib.lib().rg_mul_cast_pad_bf16(_p(src), _p(mul), float(scale), _p(out), rows, cols, out.shape[1],
                                               _st()), "rg_mul_cast_pad_bf16")
    return out


def vae_reparam(mulv, eps, z, partial):
    B, Z = eps.shape
    n = _lib.lib().rg_vae_reparam(_p(mulv), _p(eps), B, Z, _p(z), _p(partial), partial.numel(), _st())
    if n <= 0:
        _lib.check(n if n < 0 else -1, "rg_vae_reparam")
    return n


def vae_recon(pre, x, gscale, dpre, partial):
    B, F = x.shape
    n = _lib.lib().rg_vae_recon(_p(pre), pre.stride(0), _p(x), B, F, float(gscale), _p(dpre), _p(partial),
                                partial.numel(), _st())
    if n <= 0:
        _lib.check(n if n < 0 else -1, "rg_vae_recon")
    return n


def vae_latent_grad(dz, mulv, eps, kscale, dcat):
    B, Z = eps.shape
    _lib.check(_lib.lib().rg_vae_latent_grad(_p(dz), _p(mulv), _p(eps), B, Z, float(kscale), _p(dcat), _st()),
               "rg_vae_latent_grad")


def vae_loss_finalize(p_sse, n1, p_kld, n2, B, F, beta, out3):
    _lib.check(_lib.lib().rg_vae_loss_finalize(_p(p_sse), n1, _p(p_kld), n2, B, F, float(beta), _p(out3), _st()),
               "rg_vae_loss_finalize")


# ------------------------------------------------------------------------------------------------ resize-conv generator
def upsample2x_reflectpad(h, out):
    B, H, W, C = h.shape
    _lib.check(_lib.lib().rg_upsample2x_reflectpad(_p(h), _p(out), B, H, W, C, _st()), "rg_upsample2x_reflectpad")
    return out


def upsample2x_reflectpad_bwd(du, out):
    B, H, W, C = out.shape
    _lib.check(_lib.lib().rg_upsample2x_reflectpad_bwd(_p(du), _p(out), B, H, W, C, _st()),
               "rg_upsample2x_reflectpad_bwd")
    return out


def pack_conv3(W, out):
    Cout, Cin = W.shape[0], W.shape[1]
    _lib.check(_lib.lib().rg_pack_conv3(_p(W), _p(out), Cout, Cin, out.shape[0], _st()), "rg_pack_conv3")
    return out


def conv3x3(u, w3, out, bias=None, stats=None):
    """u bf16 [B, Ho+2, Wo+2, Cin] -> out bf16 NHWC [B, Ho, Wo, Cout] or fp32 NCHW [B, Cimg, Ho, Wo]."""
    B, Hp, Wp, Cin = u.shape
    Ho, Wo = Hp - 2, Wp - 2
    if out.dtype == torch.float32:
        Cimg = out.shape[1]
        _prof("conv3x3", 2.0 * B * Ho * Wo * Cimg * 9 * Cin, lambda: _lib.check(
            _lib.lib().rg_conv3x3_img(_p(u), _p(w3), _p(out), _p(bias), B, Ho, Wo, Cin, Cimg, _st()), "rg_conv3x3_img"))
    else:
        Cout = out.shape[3]
        _prof("conv3x3", 2.0 * B * Ho * Wo * Cout * 9 * Cin, lambda: _lib.check(
            _lib.lib().rg_conv3x3(_p(u), _p(w3), _p(out), _p(bias), B, Ho, Wo, Cin, Cout, _p(stats), _st()),
            "rg_conv3x3"))
    return out


def conv3x3_dgrad(da, w3, du):
    B, Ho, Wo, Cout = da.shape
    Cin = du.shape[3]
    _prof("conv3x3_dgrad", 2.0 * B * Ho * Wo * Cout * 9 * Cin, lambda: _lib.check(
        _lib.lib().rg_conv3x3_dgrad(_p(da), _p(w3), _p(du), B, Ho, Wo, Cin, Cout, _st()), "rg_conv3x3_dgrad"))
    return du


def conv3x3_wgrad(da, u, dW, beta=0.0):
    B, Ho, Wo, Cout = da.shape
    Cin = u.shape[3]
    L = _lib.lib()
    ws = _workspace(L.rg_conv3x3_wgrad_ws_bytes(B, Ho, Wo, Cin, Cout), da.device)
    _prof("conv3x3_wgrad", 2.0 * B * Ho * Wo * Cout * 9 * Cin, lambda: _lib.check(
        L.rg_conv3x3_wgrad(_p(da), _p(u), _p(dW), _p(ws), ws.numel() * 4, B, Ho, Wo, Cin, Cout, float(beta), _st()),
        "rg_conv3x3_wgrad"))
    return dW


def upg_last_bwd(u, dout, W, dW, du):
    B, Cimg, S, _ = dout.shape
    C = u.shape[3]
    L = _lib.lib()
    ws = _workspace(L.rg_upg_last_ws_bytes(B, S, C, Cimg), u.device)
    _lib.check(L.rg_upg_last_bwd(_p(u), _p(dout), _p(W), B, S, C, Cimg, _p(dW), _p(du), _p(ws), ws.numel() * 4, _st()),
               "rg_upg_last_bwd")
