"""ctypes binding of librnagan_b200.so (the C ABI declared in include/rnagan_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is raised.
The product path never routes through oracle/ or a PyTorch re-implementation.
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RG_LIB_PATH") or os.path.join(_HERE, "librnagan_b200.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["rg_gemm_api.cu", "rg_ops.cu", "rg_img.cu", "rg_data.cu"]

_c = ctypes
_vp, _i, _f, _sz = _c.c_void_p, _c.c_int, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/rnagan_b200.h one to one
SIGNATURES = {
    "rg_version": (_i, []),
    "rg_last_error": (_c.c_char_p, []),
    "rg_check_device": (_i, []),
    "rg_launch_count": (_c.c_longlong, []),
    "rg_debug_set_prof": (None, [_vp]),
    "rg_pack_link": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "rg_pack_proj": (_i, [_vp, _vp, _i, _i, _vp]),
    "rg_pack_up_from_down": (_i, [_vp, _vp, _i, _i, _vp]),
    "rg_up9_elems": (_sz, [_i]),
    "rg_pack_up9_from_down": (_i, [_vp, _vp, _i, _i, _vp]),
    "rg_pack_edge": (_i, [_vp, _vp, _i, _i, _vp]),
    "rg_cast_pad_bf16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "rg_stats_ws_bytes": (_sz, [_i]),
    "rg_stats_parts": (_i, []),
    "rg_conv_down": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rg_conv_up": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rg_gemm_nt_bwd": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rg_reduce_partials": (_i, [_vp, _i, _vp, _vp]),
    "rg_conv_up_img": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "rg_conv_wgrad_ws_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "rg_conv_wgrad": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _f, _vp, _f, _i, _vp]),
    "rg_proj_wgrad_ws_bytes": (_sz, [_i, _i, _i]),
    "rg_proj_wgrad": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _vp, _f, _i, _vp]),
    "rg_gemm_nt": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _f, _i, _vp]),
    "rg_gemm_nt_ld": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _f, _i, _vp]),
    "rg_gemm_nn": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _f, _i, _vp]),
    "rg_gemm_tn_ld": (_i, [_vp, _i, _vp, _i, _vp, _vp, _sz, _i, _i, _i, _f, _vp, _f, _vp]),
    "rg_gemm_tn_ws_bytes": (_sz, [_i, _i, _i]),
    "rg_gemm_tn": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _vp, _f, _vp]),
    "rg_reduce_ws_bytes": (_sz, [_i, _i]),
    "rg_bn_stats": (_i, [_vp, _i, _i, _vp, _sz, _vp, _vp]),
    "rg_bn_finalize": (_i, [_vp, _vp, _vp, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rg_bn_finalize_partials": (_i, [_vp, _vp, _vp, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rg_bn_act": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _vp]),
    "rg_bn_bwd_reduce": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _vp, _sz, _vp, _vp]),
    "rg_bn_bwd_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _vp, _vp, _vp]),
    "rg_bn_param_grads": (_i, [_vp, _vp, _vp, _i, _f, _f, _vp]),
    "rg_lrelu_bwd": (_i, [_vp, _vp, _f, _vp, _i, _i, _vp]),
    "rg_col_sum": (_i, [_vp, _i, _i, _vp, _sz, _vp, _vp, _f, _vp]),
    "rg_bn_gp_reduce": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _sz, _vp, _vp]),
    "rg_bn_gp_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _i, _vp, _vp, _vp, _f, _vp]),
    "rg_latent_prep": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "rg_im2col_img": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "rg_img_channel_sum": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _f, _vp]),
    "rg_col2im_img": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rg_img_conv_up": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp]),
    "rg_img_conv_up_pack_bytes": (_sz, []),
    "rg_img_conv_up_pack": (_i, [_vp, _i, _i, _vp, _vp]),
    "rg_img_conv_down": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _f, _vp, _f, _i, _i, _i, _i, _vp, _vp]),
    "rg_img_conv_wgrad_ws_bytes": (_sz, []),
    "rg_img_conv_wgrad": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp, _f, _vp, _f, _vp]),
    "rg_pack_edge_t": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "rg_unpack_edge_grad": (_i, [_vp, _vp, _i, _i, _f, _vp]),
    "rg_pack_head": (_i, [_vp, _vp, _i, _vp]),
    "rg_head_fwd": (_i, [_vp, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "rg_head_bwd_data": (_i, [_vp, _vp, _f, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "rg_head_wgrad": (_i, [_vp, _vp, _i, _i, _i, _vp, _f, _vp]),
    "rg_wgan_loss": (_i, [_vp, _f, _vp, _f, _i, _vp, _vp]),
    "rg_gp_norm": (_i, [_vp, _sz, _f, _vp, _i, _vp, _vp]),
    "rg_adam_table_bytes": (_i, [_i]),
    "rg_adam_build_table": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i]),
    "rg_adam_build_table_pitched": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _i]),
    "rg_adam_step": (_i, [_vp, _i, _f, _f, _f, _f, _i, _i, _f, _f, _f, _vp]),
    "rg_adam_step_dyn": (_i, [_vp, _i, _vp, _f, _f, _f, _i, _f, _f, _f, _vp]),
    "rg_clamp": (_i, [_vp, _sz, _f, _f, _vp]),
    "rg_slices_sum": (_i, [_vp, _i, _sz, _sz, _vp, _vp]),
    "rg_nvls_allreduce": (_i, [_vp, _sz, _sz, _i, _vp]),
    "rg_tiles_u8_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rg_lmdb_open": (_vp, [_c.c_char_p]),
    "rg_lmdb_close": (None, [_vp]),
    "rg_lmdb_stat": (_i, [_vp, _c.POINTER(_c.c_ulonglong), _c.POINTER(_c.c_uint), _c.POINTER(_c.c_uint)]),
    "rg_lmdb_get": (_i, [_vp, _c.c_char_p, _sz, _c.POINTER(_vp), _c.POINTER(_sz)]),
    "rg_lz4f_decompress": (_c.c_longlong, [_c.c_char_p, _sz, _vp, _sz, _c.POINTER(_c.c_longlong)]),
    "rg_tiles_to_unit_nhwc": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "rg_upsample2x_reflectpad": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rg_upsample2x_reflectpad_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rg_pack_conv3": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "rg_conv3x3": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "rg_conv3x3_img": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rg_conv3x3_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rg_conv3x3_wgrad_ws_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "rg_conv3x3_wgrad": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _f, _vp]),
    "rg_upg_last_ws_bytes": (_sz, [_i, _i, _i, _i]),
    "rg_upg_last_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "rg_mul_cast_pad_bf16": (_i, [_vp, _vp, _f, _vp, _i, _i, _i, _vp]),
    "rg_vae_reparam": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "rg_vae_recon": (_i, [_vp, _i, _vp, _i, _i, _f, _vp, _vp, _i, _vp]),
    "rg_vae_latent_grad": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp, _vp]),
    "rg_vae_loss_finalize": (_i, [_vp, _i, _vp, _i, _i, _i, _f, _vp, _vp]),
}


class EpilogueAux(ctypes.Structure):
    """rg_epilogue_aux of include/rnagan_b200.h."""
    _fields_ = [("aux", _vp), ("mode", _i), ("mean", _vp), ("rstd", _vp), ("scale", _vp), ("shift", _vp),
                ("slope", _f)]


_lib = None


def build(verbose=False):
    """Compile every CUDA source for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [
        "nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH,
    ] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building librnagan_b200.so")
    global _lib
    _lib = None
    return LIB_PATH


def _needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh")) and os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def lib():
    """Return the loaded library, failing loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class RgError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib().rg_last_error().decode("utf-8", "replace")
        raise RgError(f"{what} failed with status {rc}: {msg}")
