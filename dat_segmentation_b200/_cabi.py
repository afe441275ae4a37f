"""ctypes binding of the C ABI in include/dat_b200.h (libdat_b200.so).

The library is the product; this file only marshals raw device pointers and the
current CUDA stream.  There is no fallback: if the shared library is missing the
import of the CUDA path fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DAT_B200_LIB: another build of the same library (A/B timing of two builds in one process environment: tools/ only)
LIB_PATH = os.environ.get("DAT_B200_LIB") or os.path.join(_HERE, "libdat_b200.so")

DAT_F32, DAT_BF16 = 0, 1
HEAD_DIM = 32


class BlockDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("n_heads", C.c_int32),
                ("n_groups", C.c_int32), ("stride", C.c_int32), ("ksize", C.c_int32),
                ("table_h", C.c_int32), ("table_w", C.c_int32),
                ("offset_range_factor", C.c_float), ("x_dtype", C.c_int32),
                ("act_dtype", C.c_int32), ("pe_mode", C.c_int32), ("no_off", C.c_int32)]


# dat_block_desc.pe_mode (include/dat_b200.h)
PE_RPE, PE_NONE, PE_DWC, PE_FIXED, PE_LOGCPB = 0, 1, 2, 3, 4


PARAM_FIELDS = ("off_dw_w", "off_dw_b", "off_ln_g", "off_ln_b", "off_pw_w", "wq", "bq", "wk", "bk",
                "wv", "bv", "wo", "bo", "rpe_table", "pe_b", "pe_w2")
# state-dict key of every field (reference names, dat_blocks.py:51-104)
PARAM_KEYS = ("conv_offset.0.weight", "conv_offset.0.bias", "conv_offset.1.norm.weight",
              "conv_offset.1.norm.bias", "conv_offset.3.weight", "proj_q.weight", "proj_q.bias",
              "proj_k.weight", "proj_k.bias", "proj_v.weight", "proj_v.bias", "proj_out.weight",
              "proj_out.bias", "rpe_table")
SAVED_FIELDS = ("q", "t_dw", "off_raw", "pos", "xs", "k", "v", "o", "lse")


# optional bf16 operand copies of the four projection weights (dat_block_params.w*_bf16; 0 = the library casts)
BF16_FIELDS = ("wq_bf16", "wk_bf16", "wv_bf16", "wo_bf16")


class BlockParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS + BF16_FIELDS]


class CastItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", C.c_int64)]


class BlockGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


class BlockSaved(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in SAVED_FIELDS]


class DatError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libdat_b200.so (built by dat_segmentation_b200/build.py).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DatError(f"{LIB_PATH} is missing — build it with `python -m dat_segmentation_b200.build` "
                       "(there is no CPU or PyTorch fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    dp = C.POINTER(BlockDesc)
    L.dat_last_error.restype = C.c_char_p
    L.dat_version.restype = C.c_char_p
    L.dat_launch_count.restype = C.c_uint64
    L.dat_sample_grid.argtypes = [dp, C.POINTER(i32), C.POINTER(i32)]
    L.dat_block_fwd_workspace_bytes.argtypes = [dp]
    L.dat_block_fwd_workspace_bytes.restype = C.c_size_t
    L.dat_block_bwd_workspace_bytes.argtypes = [dp]
    L.dat_block_bwd_workspace_bytes.restype = C.c_size_t
    L.dat_block_forward.argtypes = [dp, C.POINTER(BlockParams), vp, vp, C.POINTER(BlockSaved), vp,
                                    C.c_size_t, vp]
    L.dat_block_backward.argtypes = [dp, C.POINTER(BlockParams), vp, vp, C.POINTER(BlockSaved), vp,
                                     C.POINTER(BlockGrads), vp, C.c_size_t, vp]
    L.dat_pointwise_fwd.argtypes = [vp, i32, f32p, f32p, vp, i32, i64, i32, i32, vp]
    L.dat_pointwise_fwd_tc.argtypes = [vp, i32, vp, f32p, vp, i32, i64, i32, i32, vp]
    L.dat_pointwise_fwd_tc_residual.argtypes = [vp, i32, vp, f32p, f32p, f32p, i64, f32p, i64, i32, i32, vp]
    L.dat_pointwise_fwd_tc_residual.restype = C.c_int
    L.dat_cast_bf16.argtypes = [f32p, vp, i64, vp]
    L.dat_cast_transpose_bf16.argtypes = [f32p, vp, i32, i32, vp]
    L.dat_cast_bf16_multi.argtypes = [vp, i32, vp]
    L.dat_pointwise_dgrad_tc.argtypes = [vp, vp, vp, i32, i64, i32, i32, vp]
    L.dat_cast_bf16_multi.restype = L.dat_pointwise_dgrad_tc.restype = C.c_int
    L.dat_pointwise_wgrad_tc_workspace_bytes.argtypes = [i64, i32, i32]
    L.dat_pointwise_wgrad_tc_workspace_bytes.restype = C.c_size_t
    L.dat_pointwise_wgrad_tc.argtypes = [vp, vp, f32p, f32p, i64, i32, i32, vp, C.c_size_t, vp]
    L.dat_bias_grad.argtypes = [vp, i32, f32p, i64, i32, vp, C.c_size_t, vp]
    L.dat_cast_transpose_bf16.restype = L.dat_pointwise_wgrad_tc.restype = L.dat_bias_grad.restype = C.c_int
    L.dat_debug_gemm_timing.argtypes = [C.POINTER(C.c_uint64)]
    L.dat_debug_gemm_timing.restype = C.c_int
    L.dat_debug_attn_bwd_timing.argtypes = [C.POINTER(C.c_uint64)]
    L.dat_debug_attn_bwd_timing.restype = C.c_int
    L.dat_offset_pos_fwd.argtypes = [dp, C.POINTER(BlockParams), vp, f32p, f32p, f32p, vp]
    L.dat_ref_points.argtypes = [i32, i32, f32p, f32p, vp]
    L.dat_sample_fwd.argtypes = [dp, vp, f32p, vp, vp, vp]
    L.dat_gather_kv_fwd.argtypes = [dp, vp, f32p, vp, vp, f32p, f32p, vp, vp, vp, vp]
    L.dat_attention_fwd.argtypes = [dp, vp, vp, vp, f32p, f32p, vp, f32p, vp, C.c_size_t, i32, vp]
    L.dat_attention_fwd_workspace_bytes.argtypes = [dp]
    L.dat_attention_fwd_workspace_bytes.restype = C.c_size_t
    L.dat_rpe_bias.argtypes = [dp, f32p, f32p, f32p, vp]
    L.dat_layernorm_fwd.argtypes = [vp, i32, f32p, f32p, vp, i32, f32p, f32p, i64, i32, C.c_float, vp]
    L.dat_layernorm_bwd_workspace_bytes.argtypes = [i64, i32]
    L.dat_layernorm_bwd_workspace_bytes.restype = C.c_size_t
    L.dat_layernorm_bwd.argtypes = [vp, i32, vp, i32, f32p, f32p, f32p, vp, vp, f32p, f32p, i64, i32, vp,
                                    C.c_size_t, vp]
    L.dat_scale_residual.argtypes = [vp, i32, vp, i32, f32p, vp, i32, i64, i64, vp]
    L.dat_scale_residual.restype = C.c_int
    L.dat_dwconv_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.dat_dwconv_workspace_bytes.restype = C.c_size_t
    L.dat_dwconv_fwd.argtypes = [vp, i32, f32p, f32p, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp,
                                 C.c_size_t, vp]
    L.dat_gelu_bwd.argtypes = [vp, vp, vp, i32, i64, vp]
    L.dat_dwconv_wgrad.argtypes = [vp, i32, vp, i32, f32p, f32p, i32, i32, i32, i32, i32, vp, C.c_size_t, vp]
    L.dat_dwconv_fwd.restype = L.dat_gelu_bwd.restype = L.dat_dwconv_wgrad.restype = C.c_int
    L.dat_dwconv_bwd.argtypes = [vp, i32, vp, vp, i32, f32p, vp, f32p, f32p, i32, i32, i32, i32, i32, i32, vp,
                                 C.c_size_t, vp]
    L.dat_dwconv_bwd.restype = C.c_int
    L.dat_layernorm_fwd.restype = C.c_int
    L.dat_residual_layernorm_fwd.argtypes = [vp, f32p, i64, vp, i32, f32p, f32p, vp, vp, i32, f32p, f32p, i64, i32,
                                             C.c_float, vp]
    L.dat_residual_layernorm_bwd.argtypes = [vp, i32, vp, i32, f32p, f32p, f32p, vp, vp, vp, f32p, i64, f32p, f32p,
                                             i64, i32, vp, C.c_size_t, vp]
    L.dat_residual_layernorm_fwd.restype = L.dat_residual_layernorm_bwd.restype = C.c_int
    L.dat_layernorm_bwd.restype = C.c_int
    L.dat_conv3x3s2_kp.argtypes = [i32]
    L.dat_conv3x3s2_kp.restype = i32
    L.dat_im2col3x3s2.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, vp]
    L.dat_col2im3x3s2.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    L.dat_conv_weight_pack.argtypes = [f32p, vp, i32, i32, vp]
    L.dat_conv_weight_unpack.argtypes = [f32p, f32p, i32, i32, vp]
    L.dat_gelu_fwd.argtypes = [vp, i32, vp, i32, i64, vp]
    L.dat_gelu_bwd_mixed.argtypes = [vp, i32, vp, vp, i32, i64, vp]
    L.dat_gelu_bwd_mixed.restype = C.c_int
    L.dat_transpose_pc.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    L.dat_transpose_pc.restype = C.c_int
    L.dat_pointwise_wgrad_workspace_bytes.argtypes = [i64, i32, i32]
    L.dat_pointwise_wgrad_workspace_bytes.restype = C.c_size_t
    L.dat_pointwise_wgrad.argtypes = [vp, i32, vp, i32, f32p, f32p, i64, i32, i32, vp, C.c_size_t, vp]
    for name in ("dat_im2col3x3s2", "dat_col2im3x3s2", "dat_conv_weight_pack", "dat_conv_weight_unpack", "dat_gelu_fwd",
                 "dat_pointwise_wgrad"):
        getattr(L, name).restype = C.c_int
    for name in ("dat_sample_grid", "dat_block_forward", "dat_block_backward", "dat_pointwise_fwd",
                 "dat_pointwise_fwd_tc", "dat_cast_bf16",
                 "dat_offset_pos_fwd", "dat_ref_points", "dat_sample_fwd", "dat_gather_kv_fwd", "dat_attention_fwd",
                 "dat_rpe_bias"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().dat_last_error().decode("utf-8", "replace")
        raise DatError(f"{what} failed (code {rc}): {msg}")


def exported_symbols():
    """Names declared in include/dat_b200.h (used by the CPU-side symbol test)."""
    return ["dat_sample_grid", "dat_block_fwd_workspace_bytes", "dat_block_bwd_workspace_bytes",
            "dat_last_error", "dat_version", "dat_launch_count", "dat_block_forward", "dat_block_backward",
            "dat_pointwise_fwd", "dat_pointwise_fwd_tc", "dat_pointwise_fwd_tc_residual", "dat_cast_bf16", "dat_debug_gemm_timing", "dat_debug_attn_bwd_timing",
            "dat_cast_transpose_bf16", "dat_pointwise_wgrad_tc_workspace_bytes", "dat_pointwise_wgrad_tc",
            "dat_cast_bf16_multi", "dat_pointwise_dgrad_tc",
            "dat_bias_grad",
            "dat_offset_pos_fwd",
            "dat_ref_points", "dat_sample_fwd", "dat_gather_kv_fwd",
            "dat_attention_fwd", "dat_attention_fwd_workspace_bytes", "dat_rpe_bias",
            "dat_layernorm_fwd", "dat_layernorm_bwd_workspace_bytes", "dat_layernorm_bwd",
            "dat_residual_layernorm_fwd", "dat_residual_layernorm_bwd",
            "dat_dwconv_workspace_bytes", "dat_dwconv_fwd", "dat_gelu_bwd", "dat_dwconv_wgrad",
            "dat_dwconv_bwd", "dat_scale_residual",
            "dat_conv3x3s2_kp", "dat_im2col3x3s2", "dat_col2im3x3s2", "dat_conv_weight_pack", "dat_conv_weight_unpack",
            "dat_gelu_fwd", "dat_gelu_bwd_mixed", "dat_transpose_pc", "dat_pointwise_wgrad_workspace_bytes", "dat_pointwise_wgrad"]
