"""ctypes binding of ``libjl_b200.so`` (C ABI declared in ``include/jl_b200.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing, or a call returns a
negative status, this module raises.  The structures below mirror the header field for field;
``tests/test_abi.py`` checks their size / offsets against the C compiler's.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JL_B200_LIB: tuning aid — an alternative build of the same sources (scripts/build_variant.sh) for A/B runs; default = the in-tree library
LIB_PATH = os.environ.get("JL_B200_LIB") or os.path.join(_HERE, "libjl_b200.so")

JL_OK, JL_EINVAL, JL_EUNSUPPORTED_SHAPE, JL_ECUDA, JL_EUNSUPPORTED = 0, -1, -2, -3, -4
JL_DT_BF16, JL_DT_F32 = 0, 1
JL_EPI_NONE, JL_EPI_GELU, JL_EPI_RELU, JL_EPI_GELU_BWD, JL_EPI_RELU_BWD, JL_EPI_GLU, JL_EPI_GELU_DGELU, JL_EPI_MUL_AUX, JL_EPI_ARGMAX = range(9)
JL_LAYOUT_K, JL_LAYOUT_MN = 0, 1
JL_CTC_SUM, JL_CTC_MEAN = 0, 1
JL_MEL_BINS, JL_MEL_MAXW, JL_MEL_FRAMES_PER_CTA = 80, 32, 32
DEFAULT_ATTN_IMPL = 0     # what the library starts with (0 = tcgen05 kernels, 1 = mma.sync kernels)

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class MelCmvnParams(C.Structure):
    _fields_ = [("wave", vp), ("wave_stride", i64), ("num_samples", vp), ("batch", i32), ("max_frames", i32),
                ("window", vp), ("twiddle", vp), ("mel_lo", vp), ("mel_cnt", vp), ("mel_w", vp), ("feats", vp),
                ("feats_bf16", vp), ("attention_mask", vp), ("frame_lengths", vp), ("apply_cmvn", i32)]


class GemmParams(C.Structure):
    _fields_ = [("a", vp), ("lda", i64), ("b", vp), ("ldb", i64), ("a_layout", i32), ("b_layout", i32),
                ("c", vp), ("ldc", i64), ("bias", vp), ("residual", vp), ("ldr", i64), ("aux", vp), ("ldaux", i64),
                ("aux_out", vp), ("ldaux_out", i64), ("row_lengths", vp), ("rows_per_seq", i32),
                ("m", i32), ("n", i32), ("k", i32), ("epilogue", i32), ("out_dtype", i32), ("alpha", f32),
                ("workspace", vp), ("workspace_bytes", i64)]


class LayerNormFwdParams(C.Structure):
    _fields_ = [("x", vp), ("ldx", i64), ("gamma", vp), ("beta", vp), ("y", vp), ("ldy", i64), ("mean", vp),
                ("rstd", vp), ("rows", i32), ("d", i32), ("eps", f32), ("act", i32)]


class LayerNormBwdParams(C.Structure):
    _fields_ = [("dy", vp), ("lddy", i64), ("x", vp), ("ldx", i64), ("gamma", vp), ("mean", vp), ("rstd", vp),
                ("dres", vp), ("lddres", i64), ("dx", vp), ("lddx", i64), ("dgamma", vp), ("dbeta", vp),
                ("partial", vp), ("rows", i32), ("d", i32)]


class ColReduceJob(C.Structure):
    _fields_ = [("dy", vp), ("lddy", i64), ("x", vp), ("ldx", i64), ("mean", vp), ("rstd", vp), ("rows", i32), ("cols", i32),
                ("out_sum", vp), ("out_dot", vp)]


class WFAdapterFwdParams(C.Structure):
    _fields_ = [("h", vp), ("ldh", i64), ("out", vp), ("ldo", i64), ("bd_scaled", vp), ("s", vp), ("t", vp), ("ad_pad", vp),
                ("c_d", vp), ("bu", vp), ("au_pad", vp), ("c_u", vp), ("row_lengths", vp), ("rows_per_seq", i32),
                ("mean", vp), ("rstd", vp), ("rows", i32), ("d", i32), ("r", i32), ("b", i32), ("eps", f32),
                ("t1_out", vp), ("u_out", vp), ("t2_out", vp)]


class WFAdapterPackParams(C.Structure):
    _fields_ = [("down_B", vp), ("down_A", vp), ("up_A", vp), ("gamma", vp), ("beta", vp), ("bd_scaled", vp), ("s", vp), ("t", vp),
                ("ad_pad", vp), ("au_pad", vp), ("sets", i32), ("d", i32), ("r", i32), ("b", i32)]


class AttAdapterFwdParams(C.Structure):
    _fields_ = [("h", vp), ("ldh", i64), ("out", vp), ("ldo", i64), ("wqkv_scaled", vp), ("s", vp), ("tb", vp), ("wo", vp), ("bo", vp),
                ("lengths", vp), ("cu_seqlens", vp), ("total_rows", i32), ("batch", i32), ("seq", i32), ("d", i32), ("scale", f32), ("eps", f32),
                ("zero_padded_rows", i32), ("qkv_out", vp), ("a_out", vp), ("mean", vp), ("rstd", vp), ("lse", vp), ("col_split", i32)]


class LnFoldPackParams(C.Structure):
    _fields_ = [("w", vp), ("bias", vp), ("gamma", vp), ("beta", vp), ("w_scaled", vp), ("s", vp), ("tb", vp), ("n", i32), ("d", i32)]


class LnProjBwdParams(C.Structure):
    _fields_ = [("dy", vp), ("lddy", i64), ("y", vp), ("ldy", i64), ("w", vp), ("s", vp), ("tb", vp), ("gamma", vp), ("h", vp), ("ldh", i64),
                ("mean", vp), ("rstd", vp), ("dres", vp), ("lddres", i64), ("dx", vp), ("lddx", i64), ("dz", vp), ("lddz", i64),
                ("rows", i32), ("n", i32), ("d", i32), ("col_partial", vp),
                ("dy_scaled", vp), ("lddys", i64), ("wgrad_partial", vp), ("col_split", i32),
                ("num_runs", i32), ("w_sets", i32), ("run_start", i32 * 9), ("run_set", i32 * 8)]


class FusionParams(C.Structure):
    _fields_ = [("h", vp), ("ldh", i64), ("y", vp), ("ldy", i64), ("y_stride", i64), ("q", vp), ("ldq", i64), ("key", vp), ("ldkey", i64),
                ("key_stride", i64), ("out", vp), ("ldo", i64), ("alpha", vp), ("dout", vp), ("lddout", i64), ("dy", vp), ("lddy", i64),
                ("dy_stride", i64), ("dq", vp), ("lddq", i64), ("dkey", vp), ("lddkey", i64), ("dkey_stride", i64),
                ("rows", i32), ("d", i32), ("b", i32), ("num_adapters", i32), ("scale", f32), ("row_lengths", vp), ("rows_per_seq", i32)]


class AttnFwdParams(C.Structure):
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("ld_qkv", i64), ("o", vp), ("ld_o", i64), ("lse", vp),
                ("lengths", vp), ("batch", i32), ("seq", i32), ("heads", i32), ("scale", f32), ("cu_seqlens", vp), ("total_rows", i32)]


class AttnBwdParams(C.Structure):
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("ld_qkv", i64), ("o", vp), ("d_o", vp), ("ld_o", i64), ("lse", vp),
                ("dq", vp), ("dk", vp), ("dv", vp), ("ld_dqkv", i64), ("delta", vp), ("lengths", vp),
                ("batch", i32), ("seq", i32), ("heads", i32), ("scale", f32), ("cu_seqlens", vp), ("total_rows", i32)]


class CtcParams(C.Structure):
    _fields_ = [("logits", vp), ("ld_logits", i64), ("logits_dtype", i32), ("labels", vp), ("max_label_len", i32),
                ("input_lengths", vp), ("batch", i32), ("seq", i32), ("vocab", i32), ("blank", i32),
                ("reduction", i32), ("zero_infinity", i32), ("nll", vp), ("loss", vp), ("grad", vp), ("ld_grad", i64),
                ("grad_dtype", i32), ("cu_seqlens", vp)]


class CtcGreedyParams(C.Structure):
    _fields_ = [("logits", vp), ("ld_logits", i64), ("logits_dtype", i32), ("input_lengths", vp), ("batch", i32),
                ("seq", i32), ("vocab", i32), ("blank", i32), ("frame_ids", vp), ("out_ids", vp), ("out_lengths", vp),
                ("cu_seqlens", vp)]


class AdamWParams(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("param_bf16", vp), ("n", i64),
                ("lr", f32), ("beta1", f32), ("beta2", f32), ("eps", f32), ("weight_decay", f32), ("grad_scale", f32),
                ("step", i32), ("hyper_dev", vp)]


# name -> (restype, argtypes); every symbol include/jl_b200.h declares
SYMBOLS = {
    "jl_version": (C.c_int, []),
    "jl_last_error": (C.c_char_p, []),
    "jl_launch_count": (C.c_int64, []),
    "jl_launch_count_reset": (None, []),
    "jl_mel_cmvn_workspace_bytes": (C.c_int, [C.POINTER(MelCmvnParams), C.POINTER(C.c_size_t)]),
    "jl_mel_cmvn_fwd": (C.c_int, [C.POINTER(MelCmvnParams), vp, vp]),
    "jl_gemm_bf16": (C.c_int, [C.POINTER(GemmParams), vp]),
    "jl_debug_gemm_ref": (C.c_int, [C.POINTER(GemmParams), vp]),
    "jl_debug_set_gemm_mode": (None, [C.c_int]),
    "jl_debug_set_gemm_bn": (None, [C.c_int]),
    "jl_debug_set_gemm_tail": (None, [C.c_int]),
    "jl_gemm_workspace_zero_bytes": (C.c_int, [C.POINTER(GemmParams), C.POINTER(C.c_size_t)]),
    "jl_debug_set_attn_impl": (None, [C.c_int]),
    "jl_debug_set_pdl": (None, [C.c_int]),
    "jl_debug_set_lnproj_split": (None, [i32]),
    "jl_gemm_workspace_bytes": (C.c_int, [C.POINTER(GemmParams), C.POINTER(C.c_size_t)]),
    "jl_layernorm_fwd": (C.c_int, [C.POINTER(LayerNormFwdParams), vp]),
    "jl_layernorm_bwd_workspace_bytes": (C.c_int, [C.POINTER(LayerNormBwdParams), C.POINTER(C.c_size_t)]),
    "jl_layernorm_bwd": (C.c_int, [C.POINTER(LayerNormBwdParams), vp]),
    "jl_layernorm_wgrad": (C.c_int, [C.POINTER(LayerNormBwdParams), vp]),
    "jl_colreduce_multi": (C.c_int, [C.POINTER(ColReduceJob), i32, vp]),
    "jl_wfadapter_fwd": (C.c_int, [C.POINTER(WFAdapterFwdParams), vp]),
    "jl_wfadapter_pack": (C.c_int, [C.POINTER(WFAdapterPackParams), vp]),
    "jl_attadapter_fwd": (C.c_int, [C.POINTER(AttAdapterFwdParams), vp]),
    "jl_lnfold_pack": (C.c_int, [C.POINTER(LnFoldPackParams), vp]),
    "jl_lnfold_pack_multi": (C.c_int, [C.POINTER(LnFoldPackParams), i32, vp]),
    "jl_lnproj_bwd": (C.c_int, [C.POINTER(LnProjBwdParams), vp]),
    "jl_lnproj_bwd_reduce": (C.c_int, [vp, i32, i32, vp, vp, vp, i32, i32, i32, vp]),
    "jl_lnproj_wgrad": (C.c_int, [vp, i64, vp, i32, i32, i32, vp, vp, vp, vp]),
    "jl_lnproj_wgrad_prep": (C.c_int, [vp, i64, vp, vp, i32, i32, vp, i64, vp, vp]),
    "jl_fusion_combine_fwd": (C.c_int, [C.POINTER(FusionParams), vp]),
    "jl_fusion_combine_bwd": (C.c_int, [C.POINTER(FusionParams), vp]),
    "jl_attn_fwd": (C.c_int, [C.POINTER(AttnFwdParams), vp]),
    "jl_attn_bwd": (C.c_int, [C.POINTER(AttnBwdParams), vp]),
    "jl_ctc_workspace_bytes": (C.c_int, [C.POINTER(CtcParams), C.POINTER(C.c_size_t)]),
    "jl_ctc_fwd": (C.c_int, [C.POINTER(CtcParams), vp, vp]),
    "jl_ctc_greedy": (C.c_int, [C.POINTER(CtcGreedyParams), vp]),
    "jl_ctc_greedy_from_partials": (C.c_int, [vp, vp, i64, i32, vp, vp, i32, i32, i32, vp, vp, vp, vp]),
    "jl_im2col_k5s2": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
    "jl_embed_positions": (C.c_int, [vp, f32, vp, vp, i32, i32, i32, vp]),
    "jl_embed_positions_packed": (C.c_int, [vp, vp, f32, vp, vp, i32, i32, i32, vp]),
    "jl_transpose_bf16": (C.c_int, [vp, i64, vp, i64, i32, i32, vp]),
    "jl_wave_stats": (C.c_int, [vp, i64, vp, i32, i32, vp, vp]),
    "jl_wave_im2col": (C.c_int, [vp, i64, vp, i32, i32, vp, vp, i32, i32, i32, vp]),
    "jl_im2col_1d": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "jl_colsum_bf16": (C.c_int, [vp, i64, vp, i32, i32, vp, vp]),
    "jl_colsum_workspace_bytes": (C.c_int, [i32, i32, C.POINTER(C.c_size_t)]),
    "jl_cast_f32_to_bf16": (C.c_int, [vp, vp, i64, vp]),
    "jl_add_bf16": (C.c_int, [vp, vp, vp, i64, vp]),
    "jl_adamw_bucket": (C.c_int, [C.POINTER(AdamWParams), vp]),
    "jl_adamw_advance": (C.c_int, [vp, vp]),
    "jl_comm_unique_id": (C.c_int, [vp]),
    "jl_comm_init": (C.c_int, [vp, i32, i32, C.POINTER(vp)]),
    "jl_comm_allreduce": (C.c_int, [vp, vp, C.c_size_t, vp]),
    "jl_comm_rank": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32)]),
    "jl_comm_destroy": (C.c_int, [vp]),
}
COMM_ID_BYTES = 128


class JLError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libjl_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libjl_b200.so (built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != JL_OK:
        msg = load().jl_last_error()
        raise JLError(rc, msg.decode("utf-8", "replace") if msg else "")


def launch_count() -> int:
    return int(load().jl_launch_count())


def launch_count_reset() -> None:
    load().jl_launch_count_reset()
