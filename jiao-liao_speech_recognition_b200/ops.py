"""Tensor-level wrappers over the C ABI (``include/jl_b200.h``).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every computation below is a call into
``libjl_b200.so``.  All wrappers enqueue on ``torch.cuda.current_stream()`` and never synchronise.  There is no
fallback path: CPU tensors are rejected.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32

# When set to a list, every gemm() call appends (params, kept-alive tensors, 2·M·N·K): bench.py replays the list to
# time the step's tensor-core launches on their own (roofline of the dominant kernel).
GEMM_TRACE = None
# True while the deferred weight-gradient products of a backward pass are being issued (modeling._SideBranch.join): split-K is off,
# every product keeps its whole K range on its own CTAs and the products of all layers run concurrently on a pool of streams.
GEMM_NO_SPLIT = False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str, ndim: Optional[int] = None) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (libjl_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got {tuple(t.shape)}")


def _rows2d(t: torch.Tensor, name: str) -> None:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")


# ----------------------------------------------------------------------------------------------- mel + CMVN
def mel_cmvn(wave: torch.Tensor, num_samples: torch.Tensor, tables: dict, max_frames: int, apply_cmvn: bool = True,
             want_bf16: bool = False):
    """wave [B, N] fp32, num_samples [B] int32 → (feats fp32 [B, F, 80], feats bf16 | None, mask int32 [B, F], frame_lengths int32 [B])."""
    _need(wave, F32, "wave", 2)
    _need(num_samples, I32, "num_samples", 1)
    if wave.stride(1) != 1:
        raise ValueError("wave: inner stride must be 1")
    b = wave.shape[0]
    dev = wave.device
    feats = torch.empty((b, max_frames, L.JL_MEL_BINS), dtype=F32, device=dev)
    feats16 = torch.empty((b, max_frames, L.JL_MEL_BINS), dtype=BF16, device=dev) if want_bf16 else None
    mask = torch.empty((b, max_frames), dtype=I32, device=dev)
    flen = torch.empty((b,), dtype=I32, device=dev)
    p = L.MelCmvnParams(wave=wave.data_ptr(), wave_stride=wave.stride(0), num_samples=num_samples.data_ptr(), batch=b,
                        max_frames=max_frames, window=tables["window"].data_ptr(), twiddle=tables["twiddle"].data_ptr(),
                        mel_lo=tables["mel_lo"].data_ptr(), mel_cnt=tables["mel_cnt"].data_ptr(), mel_w=tables["mel_w"].data_ptr(),
                        feats=feats.data_ptr(), feats_bf16=_ptr(feats16), attention_mask=mask.data_ptr(),
                        frame_lengths=flen.data_ptr(), apply_cmvn=1 if apply_cmvn else 0)
    lib = L.load()
    nbytes = C.c_size_t(0)
    L.check(lib.jl_mel_cmvn_workspace_bytes(C.byref(p), C.byref(nbytes)))
    ws = torch.empty((max(nbytes.value, 16),), dtype=torch.uint8, device=dev)
    L.check(lib.jl_mel_cmvn_fwd(C.byref(p), ws.data_ptr(), _stream()))
    return feats, feats16, mask, flen


# ----------------------------------------------------------------------------------------------- GEMM
_TAIL_WS: dict = {}


def _tail_workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Scratch for the pair kernel's tail split: its leading arrival counters must be zero on entry (the kernel leaves them
    zero), and products that may run concurrently must not share it — one zero-initialised buffer per (device, stream),
    grown on demand."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _TAIL_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros((max(nbytes, 24 << 20),), dtype=torch.uint8, device=device)
        _TAIL_WS[key] = ws
    return ws


def gemm(a: torch.Tensor, b: torch.Tensor, *, bias: Optional[torch.Tensor] = None, epilogue: int = L.JL_EPI_NONE,
         residual: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None, aux_out: Optional[torch.Tensor] = None,
         row_lengths: Optional[torch.Tensor] = None, rows_per_seq: int = 0, out: Optional[torch.Tensor] = None,
         out_dtype=BF16, alpha: float = 1.0, a_layout: int = L.JL_LAYOUT_K, b_layout: int = L.JL_LAYOUT_K,
         reference: bool = False) -> torch.Tensor:
    """C[M, N] = epilogue(alpha · A · Bᵀ + bias).  A is [M, K] (layout K) or [K, M] (layout MN); B is [N, K] or [K, N]."""
    _need(a, BF16, "a")
    _need(b, BF16, "b")
    _rows2d(a, "a")
    _rows2d(b, "b")
    m, k = (a.shape[0], a.shape[1]) if a_layout == L.JL_LAYOUT_K else (a.shape[1], a.shape[0])
    n, kb = (b.shape[0], b.shape[1]) if b_layout == L.JL_LAYOUT_K else (b.shape[1], b.shape[0])
    if k != kb:
        raise ValueError(f"gemm: K mismatch ({k} vs {kb})")
    n_out = n // 2 if epilogue == L.JL_EPI_GLU else n
    if out is None:
        out = torch.empty((m, n_out), dtype=out_dtype, device=a.device)
    else:
        _rows2d(out, "out")
        if tuple(out.shape) != (m, n_out):
            raise ValueError(f"gemm: out has shape {tuple(out.shape)}, expected {(m, n_out)}")
    if bias is not None:
        _need(bias, F32, "bias", 1)
    for t, nm in ((residual, "residual"), (aux, "aux"), (aux_out, "aux_out")):
        if t is not None:
            _need(t, BF16, nm)
            _rows2d(t, nm)
    if row_lengths is not None:
        _need(row_lengths, I32, "row_lengths", 1)
    p = L.GemmParams(a=a.data_ptr(), lda=a.stride(0), b=b.data_ptr(), ldb=b.stride(0), a_layout=a_layout, b_layout=b_layout,
                     c=out.data_ptr(), ldc=out.stride(0), bias=_ptr(bias),
                     residual=_ptr(residual), ldr=0 if residual is None else residual.stride(0),
                     aux=_ptr(aux), ldaux=0 if aux is None else aux.stride(0),
                     aux_out=_ptr(aux_out), ldaux_out=0 if aux_out is None else aux_out.stride(0),
                     row_lengths=_ptr(row_lengths), rows_per_seq=rows_per_seq, m=m, n=n, k=k, epilogue=epilogue,
                     out_dtype=L.JL_DT_BF16 if out.dtype == BF16 else L.JL_DT_F32, alpha=alpha)
    if out.dtype not in (BF16, F32):
        raise TypeError("gemm: out must be bf16 or fp32")
    lib = L.load()
    ws = None
    if not reference and not GEMM_NO_SPLIT:
        nbytes, zbytes = C.c_size_t(0), C.c_size_t(0)
        L.check(lib.jl_gemm_workspace_bytes(C.byref(p), C.byref(nbytes)))
        if nbytes.value:
            L.check(lib.jl_gemm_workspace_zero_bytes(C.byref(p), C.byref(zbytes)))
            if zbytes.value:
                ws = _tail_workspace(a.device, nbytes.value)       # persistent, zero counters, one per stream
            else:
                ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=a.device)
            p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes.value
    fn = lib.jl_debug_gemm_ref if reference else lib.jl_gemm_bf16
    L.check(fn(C.byref(p), _stream()))
    if GEMM_TRACE is not None:
        GEMM_TRACE.append((p, (a, b, out, bias, residual, aux, aux_out, row_lengths, ws), 2.0 * m * n * k))
    return out


def lm_head_argmax(h: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None):
    """SURVEY §8 f1 (inference): the lm_head product h [M, d] · w [V, d]ᵀ + bias with the JL_EPI_ARGMAX epilogue — the [M, V]
    logits are never written; returns (pmax fp32 [ceil(V/32), M], pidx int32 [ceil(V/32), M]): per frame and chunk of 32 vocabulary
    entries the maximum logit and its (first) index.  ``ctc_greedy_from_partials`` turns them into token ids."""
    _need(h, BF16, "h")
    _need(w, BF16, "w")
    _rows2d(h, "h")
    _rows2d(w, "w")
    m, k = h.shape
    v, kb = w.shape
    if k != kb:
        raise ValueError(f"lm_head_argmax: K mismatch ({k} vs {kb})")
    if bias is not None:
        _need(bias, F32, "bias", 1)
    chunks = (v + 31) // 32
    ld = (m + 31) // 32 * 32
    pmax = torch.empty((chunks, ld), dtype=F32, device=h.device)
    pidx = torch.empty((chunks, ld), dtype=I32, device=h.device)
    p = L.GemmParams(a=h.data_ptr(), lda=h.stride(0), b=w.data_ptr(), ldb=w.stride(0), a_layout=L.JL_LAYOUT_K, b_layout=L.JL_LAYOUT_K,
                     c=pmax.data_ptr(), ldc=ld, bias=_ptr(bias), residual=None, ldr=0, aux=None, ldaux=0, aux_out=pidx.data_ptr(), ldaux_out=ld,
                     row_lengths=None, rows_per_seq=0, m=m, n=v, k=k, epilogue=L.JL_EPI_ARGMAX, out_dtype=L.JL_DT_F32, alpha=1.0)
    L.check(L.load().jl_gemm_bf16(C.byref(p), _stream()))
    if GEMM_TRACE is not None:
        GEMM_TRACE.append((p, (h, w, pmax, pidx, bias), 2.0 * m * v * k))
    return pmax, pidx


def ctc_greedy_from_partials(pmax: torch.Tensor, pidx: torch.Tensor, input_lengths: torch.Tensor, batch: int, seq: int, blank: int = 0,
                             cu_seqlens: Optional[torch.Tensor] = None):
    """(pmax, pidx) of ``lm_head_argmax`` → (out_ids [B, seq] int32 with -1 tail, out_lengths [B] int32, frame_ids [B, seq] int32)."""
    _need(pmax, F32, "pmax", 2)
    _need(pidx, I32, "pidx", 2)
    _need(input_lengths, I32, "input_lengths", 1)
    if pmax.shape != pidx.shape or pmax.stride(1) != 1 or pidx.stride(1) != 1 or pmax.stride(0) != pidx.stride(0):
        raise ValueError("ctc_greedy_from_partials: pmax / pidx must be equally shaped chunk-major matrices")
    if cu_seqlens is not None:
        _need(cu_seqlens, I32, "cu_seqlens", 1)
    dev = pmax.device
    frame_ids = torch.empty((batch, seq), dtype=I32, device=dev)
    out_ids = torch.empty((batch, seq), dtype=I32, device=dev)
    out_len = torch.empty((batch,), dtype=I32, device=dev)
    L.check(L.load().jl_ctc_greedy_from_partials(pmax.data_ptr(), pidx.data_ptr(), pmax.stride(0), pmax.shape[0], input_lengths.data_ptr(),
                                                 _ptr(cu_seqlens), batch, seq, blank, frame_ids.data_ptr(), out_ids.data_ptr(),
                                                 out_len.data_ptr(), _stream()))
    return out_ids, out_len, frame_ids


def replay_gemm_trace(trace) -> None:
    """Re-issue recorded GEMM launches (same operands, same shapes) on the current stream."""
    lib = L.load()
    s = _stream()
    for p, _, _ in trace:
        L.check(lib.jl_gemm_bf16(C.byref(p), s))


# ----------------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, save_stats: bool = False,
                  out: Optional[torch.Tensor] = None, gelu: bool = False):
    _need(x, BF16, "x")
    _rows2d(x, "x")
    _need(gamma, F32, "gamma", 1)
    _need(beta, F32, "beta", 1)
    rows, d = x.shape
    y = torch.empty((rows, d), dtype=BF16, device=x.device) if out is None else out
    mean = torch.empty((rows,), dtype=F32, device=x.device) if save_stats else None
    rstd = torch.empty((rows,), dtype=F32, device=x.device) if save_stats else None
    p = L.LayerNormFwdParams(x=x.data_ptr(), ldx=x.stride(0), gamma=gamma.data_ptr(), beta=beta.data_ptr(), y=y.data_ptr(),
                             ldy=y.stride(0), mean=_ptr(mean), rstd=_ptr(rstd), rows=rows, d=d, eps=eps,
                             act=L.JL_EPI_GELU if gelu else L.JL_EPI_NONE)
    L.check(L.load().jl_layernorm_fwd(C.byref(p), _stream()))
    return y, mean, rstd


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor,
                  dres: Optional[torch.Tensor] = None, want_wgrad: bool = False, dgamma: Optional[torch.Tensor] = None,
                  dbeta: Optional[torch.Tensor] = None):
    for t, nm in ((dy, "dy"), (x, "x")):
        _need(t, BF16, nm)
        _rows2d(t, nm)
    rows, d = x.shape
    dx = torch.empty((rows, d), dtype=BF16, device=x.device)
    if want_wgrad:
        dgamma = torch.empty((d,), dtype=F32, device=x.device) if dgamma is None else dgamma
        dbeta = torch.empty((d,), dtype=F32, device=x.device) if dbeta is None else dbeta
    p = L.LayerNormBwdParams(dy=dy.data_ptr(), lddy=dy.stride(0), x=x.data_ptr(), ldx=x.stride(0), gamma=gamma.data_ptr(),
                             mean=mean.data_ptr(), rstd=rstd.data_ptr(), dres=_ptr(dres),
                             lddres=0 if dres is None else dres.stride(0), dx=dx.data_ptr(), lddx=dx.stride(0),
                             dgamma=_ptr(dgamma) if want_wgrad else None, dbeta=_ptr(dbeta) if want_wgrad else None,
                             partial=None, rows=rows, d=d)
    lib = L.load()
    partial = None
    if want_wgrad:
        nbytes = C.c_size_t(0)
        L.check(lib.jl_layernorm_bwd_workspace_bytes(C.byref(p), C.byref(nbytes)))
        partial = torch.empty((max(nbytes.value // 4, 4),), dtype=F32, device=x.device)
        p.partial = partial.data_ptr()
    L.check(lib.jl_layernorm_bwd(C.byref(p), _stream()))
    return dx, dgamma, dbeta


def layernorm_wgrad(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, dgamma: torch.Tensor,
                    dbeta: Optional[torch.Tensor] = None) -> None:
    """dγ = Σ dy ∘ x̂ and dβ = Σ dy (one launch)."""
    for t, nm in ((dy, "dy"), (x, "x")):
        _need(t, BF16, nm)
        _rows2d(t, nm)
    rows, d = x.shape
    p = L.LayerNormBwdParams(dy=dy.data_ptr(), lddy=dy.stride(0), x=x.data_ptr(), ldx=x.stride(0), gamma=None, mean=mean.data_ptr(),
                             rstd=rstd.data_ptr(), dres=None, lddres=0, dx=None, lddx=0, dgamma=dgamma.data_ptr(), dbeta=_ptr(dbeta),
                             partial=None, rows=rows, d=d)
    L.check(L.load().jl_layernorm_wgrad(C.byref(p), _stream()))


def colreduce_multi(jobs) -> None:
    """Up to 4 column reductions in one launch.  Each job is a dict: ``dy`` [rows, cols] bf16, ``out_sum`` [cols] fp32 (Σ_r dy) and /
    or — with ``x`` [rows, cols] bf16, ``mean``, ``rstd`` [rows] fp32 — ``out_dot`` [cols] fp32 (Σ_r dy · (x − mean) · rstd)."""
    if not 1 <= len(jobs) <= 4:
        raise ValueError("colreduce_multi: 1..4 jobs")
    arr = (L.ColReduceJob * len(jobs))()
    for i, j in enumerate(jobs):
        dy = j["dy"]
        _need(dy, BF16, "dy")
        _rows2d(dy, "dy")
        x = j.get("x")
        if x is not None:
            _need(x, BF16, "x")
            _rows2d(x, "x")
            if x.shape != dy.shape:
                raise ValueError("colreduce_multi: x and dy must have the same shape")
        for nm in ("out_sum", "out_dot"):
            o = j.get(nm)
            if o is not None:
                _need(o, F32, nm, 1)
                if o.numel() != dy.shape[1] or not o.is_contiguous():
                    raise ValueError(f"colreduce_multi: {nm} must be a contiguous fp32 vector of {dy.shape[1]} elements")
        arr[i] = L.ColReduceJob(dy=dy.data_ptr(), lddy=dy.stride(0), x=_ptr(x), ldx=0 if x is None else x.stride(0), mean=_ptr(j.get("mean")),
                                rstd=_ptr(j.get("rstd")), rows=dy.shape[0], cols=dy.shape[1], out_sum=_ptr(j.get("out_sum")),
                                out_dot=_ptr(j.get("out_dot")))
    L.check(L.load().jl_colreduce_multi(arr, len(jobs), _stream()))


# ----------------------------------------------------------------------------------------------- fused WFAdapter
def wfadapter_fwd(h: torch.Tensor, pack: dict, eps: float, row_lengths: Optional[torch.Tensor] = None, rows_per_seq: int = 0,
                  save_stats: bool = False, out: Optional[torch.Tensor] = None, mean: Optional[torch.Tensor] = None,
                  rstd: Optional[torch.Tensor] = None, t1: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None,
                  t2: Optional[torch.Tensor] = None):
    """out = h + WFAdapter(h) in one kernel.  ``pack`` holds the kernel-layout factors (``wfadapter_pack`` /
    modeling.JLEngine._wf_pack).  Training: ``mean`` / ``rstd`` [rows] fp32 and ``t1`` [rows, r], ``u`` [rows, b], ``t2`` [rows, r]
    bf16 (contiguous) receive the LayerNorm statistics and the intermediates the backward pass needs."""
    _need(h, BF16, "h")
    _rows2d(h, "h")
    rows, d = h.shape
    if out is None:
        out = torch.empty((rows, d), dtype=BF16, device=h.device)
    else:
        _need(out, BF16, "out")
        _rows2d(out, "out")
        if tuple(out.shape) != (rows, d):
            raise ValueError(f"wfadapter_fwd: out has shape {tuple(out.shape)}, expected {(rows, d)}")
    if save_stats and mean is None:
        mean = torch.empty((rows,), dtype=F32, device=h.device)
        rstd = torch.empty((rows,), dtype=F32, device=h.device)
    r, b = pack["r"], pack["b"]
    for t, nm, w in ((t1, "t1", r), (u, "u", b), (t2, "t2", r)):
        if t is not None:
            _need(t, BF16, nm, 2)
            if tuple(t.shape) != (rows, w) or not t.is_contiguous():
                raise ValueError(f"wfadapter_fwd: {nm} must be a contiguous [{rows}, {w}] bf16 tensor")
    p = L.WFAdapterFwdParams(h=h.data_ptr(), ldh=h.stride(0), out=out.data_ptr(), ldo=out.stride(0), bd_scaled=pack["bd"].data_ptr(),
                             s=pack["s"].data_ptr(), t=pack["t"].data_ptr(), ad_pad=pack["ad"].data_ptr(), c_d=pack["c_d"].data_ptr(),
                             bu=pack["bu"].data_ptr(), au_pad=pack["au"].data_ptr(), c_u=pack["c_u"].data_ptr(),
                             row_lengths=_ptr(row_lengths), rows_per_seq=rows_per_seq, mean=_ptr(mean), rstd=_ptr(rstd), rows=rows, d=d,
                             r=r, b=b, eps=eps, t1_out=_ptr(t1), u_out=_ptr(u), t2_out=_ptr(t2))
    L.check(L.load().jl_wfadapter_fwd(C.byref(p), _stream()))
    return out, mean, rstd


def wfadapter_pack(down_B: torch.Tensor, down_A: torch.Tensor, up_A: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bufs: Optional[dict] = None) -> dict:
    """Kernel-layout operands of ``wfadapter_fwd`` for every factor set, derived on the device (one launch): ``down_B`` [K, r, d],
    ``down_A`` [K, b, r], ``up_A`` [K, d, r] bf16; ``gamma`` / ``beta`` fp32 [d].  → {"bd" [K, r, d] bf16, "s" / "t" [K, r] fp32,
    "ad" [K, b, 64], "au" [K, d, 64] bf16}; ``bufs`` (a previous result) is overwritten in place — the buffers a captured graph
    holds stay valid."""
    for t_, nm in ((down_B, "down_B"), (down_A, "down_A"), (up_A, "up_A")):
        _need(t_, BF16, nm, 3)
        if not t_.is_contiguous():
            raise ValueError(f"wfadapter_pack: {nm} must be contiguous")
    _need(gamma, F32, "gamma", 1)
    _need(beta, F32, "beta", 1)
    k, r, d = down_B.shape
    b = down_A.shape[1]
    dev = down_B.device
    if bufs is None:
        bufs = {"bd": torch.empty((k, r, d), dtype=BF16, device=dev), "s": torch.empty((k, r), dtype=F32, device=dev),
                "t": torch.empty((k, r), dtype=F32, device=dev), "ad": torch.empty((k, b, 64), dtype=BF16, device=dev),
                "au": torch.empty((k, d, 64), dtype=BF16, device=dev)}
    p = L.WFAdapterPackParams(down_B=down_B.data_ptr(), down_A=down_A.data_ptr(), up_A=up_A.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                              bd_scaled=bufs["bd"].data_ptr(), s=bufs["s"].data_ptr(), t=bufs["t"].data_ptr(), ad_pad=bufs["ad"].data_ptr(),
                              au_pad=bufs["au"].data_ptr(), sets=k, d=d, r=r, b=b)
    L.check(L.load().jl_wfadapter_pack(C.byref(p), _stream()))
    return bufs


# ----------------------------------------------------------------------------------------------- fused AttAdapter forward
def lnfold_pack(w: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, bufs: Optional[dict] = None) -> dict:
    """LayerNorm folding of the projection ``w`` [n, d] bf16 (+ ``bias`` [n] fp32) that follows LayerNorm(γ, β): → {"w" bf16 [n, d] =
    W ⊙ γ, "s" fp32 [n] its row sums, "tb" fp32 [n] = W β + bias}, one launch; ``bufs`` (a previous result) is overwritten in place."""
    _need(w, BF16, "w", 2)
    if not w.is_contiguous():
        raise ValueError("lnfold_pack: w must be contiguous")
    _need(gamma, F32, "gamma", 1)
    _need(beta, F32, "beta", 1)
    if bias is not None:
        _need(bias, F32, "bias", 1)
    n, d = w.shape
    if bufs is None:
        bufs = {"w": torch.empty((n, d), dtype=BF16, device=w.device), "s": torch.empty((n,), dtype=F32, device=w.device),
                "tb": torch.empty((n,), dtype=F32, device=w.device)}
    p = L.LnFoldPackParams(w=w.data_ptr(), bias=_ptr(bias), gamma=gamma.data_ptr(), beta=beta.data_ptr(), w_scaled=bufs["w"].data_ptr(),
                           s=bufs["s"].data_ptr(), tb=bufs["tb"].data_ptr(), n=n, d=d)
    L.check(L.load().jl_lnfold_pack(C.byref(p), _stream()))
    return bufs


LNFOLD_MAX_JOBS = 48


def lnfold_pack_multi(jobs: list) -> None:
    """``lnfold_pack`` for several projections in one launch per 48: ``jobs`` = [(w, bias | None, gamma, beta, bufs)], ``bufs`` the
    existing result dicts, overwritten in place."""
    for i0 in range(0, len(jobs), LNFOLD_MAX_JOBS):
        part = jobs[i0:i0 + LNFOLD_MAX_JOBS]
        arr = (L.LnFoldPackParams * len(part))()
        for i, (w, bias, gamma, beta, bufs) in enumerate(part):
            _need(w, BF16, "w", 2)
            if not w.is_contiguous() or bufs["w"].shape != w.shape:
                raise ValueError("lnfold_pack_multi: w must be contiguous and match its buffers")
            arr[i] = L.LnFoldPackParams(w=w.data_ptr(), bias=_ptr(bias), gamma=gamma.data_ptr(), beta=beta.data_ptr(), w_scaled=bufs["w"].data_ptr(),
                                        s=bufs["s"].data_ptr(), tb=bufs["tb"].data_ptr(), n=w.shape[0], d=w.shape[1])
        L.check(L.load().jl_lnfold_pack_multi(arr, len(part), _stream()))


def attadapter_fwd(h: torch.Tensor, pack: dict, wo: torch.Tensor, bo: torch.Tensor, lengths: Optional[torch.Tensor], batch: int, seq: int,
                   eps: float, zero_padded_rows: bool = False, training: bool = False, cu_seqlens: Optional[torch.Tensor] = None,
                   col_split: int = 0):
    """out = h + AttAdapter(h) in one kernel (utterances of <= 256 frames).  ``pack`` = ``lnfold_pack`` of the concatenated q|k|v
    projection ([192, d]); ``wo`` [d, 64] bf16, ``bo`` [d] fp32.  Returns (out, saved) with saved = (mean, rstd, qkv [rows, 192],
    a [rows, 64], lse) when ``training`` else None."""
    _need(h, BF16, "h", 2)
    _rows2d(h, "h")
    _need(wo, BF16, "wo", 2)
    _need(bo, F32, "bo", 1)
    rows, d = h.shape
    if wo.shape != (d, 64) or not wo.is_contiguous() or pack["w"].shape != (192, d):
        raise ValueError("attadapter_fwd: wo must be [d, 64] and the packed q|k|v projection [192, d]")
    if cu_seqlens is None and rows != batch * seq:
        raise ValueError(f"attadapter_fwd: h has {rows} rows, expected {batch * seq}")
    dev = h.device
    out = torch.empty((rows, d), dtype=BF16, device=dev)
    saved = None
    mean = rstd = qkv = a = lse = None
    if training:
        mean = torch.empty((rows,), dtype=F32, device=dev)
        rstd = torch.empty((rows,), dtype=F32, device=dev)
        qkv = torch.empty((rows, 192), dtype=BF16, device=dev)
        a = torch.empty((rows, 64), dtype=BF16, device=dev)
        lse = torch.empty((batch, 1, seq) if cu_seqlens is None else (1, rows), dtype=F32, device=dev)
        saved = (mean, rstd, qkv, a, lse)
    p = L.AttAdapterFwdParams(h=h.data_ptr(), ldh=h.stride(0), out=out.data_ptr(), ldo=out.stride(0), wqkv_scaled=pack["w"].data_ptr(),
                              s=pack["s"].data_ptr(), tb=pack["tb"].data_ptr(), wo=wo.data_ptr(), bo=bo.data_ptr(), lengths=_ptr(lengths),
                              cu_seqlens=_ptr(cu_seqlens), total_rows=rows if cu_seqlens is not None else 0, batch=batch, seq=seq, d=d,
                              scale=0.125, eps=eps, zero_padded_rows=1 if zero_padded_rows else 0, qkv_out=_ptr(qkv), a_out=_ptr(a),
                              mean=_ptr(mean), rstd=_ptr(rstd), lse=_ptr(lse))
    p.col_split = col_split
    L.check(L.load().jl_attadapter_fwd(C.byref(p), _stream()))
    return out, saved


def lnproj_bwd_reduce(col_partial: torch.Tensor, dgamma: Optional[torch.Tensor], dbeta: Optional[torch.Tensor], dbias: Optional[torch.Tensor],
                      accumulate: bool = False, tile_offset: int = 0, num_tiles: Optional[int] = None) -> None:
    """Finish the column sums ``lnproj_bwd(..., want_cols=True)`` left per row tile: dgamma = Σ dz·x̂, dbeta = Σ dz, dbias = Σ dres
    (fp32 [d] outputs, any may be None), summed in fixed order.  ``accumulate``: add to dgamma / dbeta instead of overwriting them.
    ``tile_offset`` / ``num_tiles``: sum only that range of row tiles (one run of a multi-run ``lnproj_bwd``)."""
    _need(col_partial, F32, "col_partial", 3)
    three, tiles, d = col_partial.shape
    if three != 3 or not col_partial.is_contiguous():
        raise ValueError("lnproj_bwd_reduce: col_partial must be a contiguous [3, row_tiles, d] tensor")
    for t_, nm in ((dgamma, "dgamma"), (dbeta, "dbeta"), (dbias, "dbias")):
        if t_ is not None:
            _need(t_, F32, nm, 1)
            if t_.numel() != d or not t_.is_contiguous():
                raise ValueError(f"lnproj_bwd_reduce: {nm} must be a contiguous [d] tensor")
    n_t = tiles if num_tiles is None else num_tiles
    L.check(L.load().jl_lnproj_bwd_reduce(col_partial.data_ptr(), n_t, d, _ptr(dgamma), _ptr(dbeta), _ptr(dbias), 1 if accumulate else 0, tile_offset, tiles,
                                          _stream()))


def lnproj_bwd(dy: torch.Tensor, y: torch.Tensor, w: torch.Tensor, pack: dict, gamma: torch.Tensor, h: torch.Tensor, mean: torch.Tensor,
               rstd: torch.Tensor, dres: torch.Tensor, want_dz: bool = False, want_cols: bool = False, want_wgrad_operands: bool = False,
               out: Optional[torch.Tensor] = None, col_split: int = 0, runs: Optional[list] = None):
    """Backward through LayerNorm → projection (W [n, d], n a multiple of 8, at most 192) in one kernel: dx = LayerNorm'(dy · W) + dres.
    ``y`` = the projection output saved by the forward pass, ``pack`` = ``lnfold_pack`` of this projection.  → (dx, dz | None);
    dz = dy · W (bf16) only when ``want_dz``.  ``want_cols``: → (dx, dz | None, col_partial [3, ⌈rows/128⌉, d] fp32) — per-row-tile column
    sums for ``lnproj_bwd_reduce`` (the LayerNorm weight gradients and the bias gradient behind ``dres`` without reading dz again)."""
    _need(dy, BF16, "dy", 2)
    _need(y, BF16, "y", 2)
    _need(w, BF16, "w", 2)
    _need(h, BF16, "h", 2)
    _need(dres, BF16, "dres", 2)
    for t_, nm in ((dy, "dy"), (y, "y"), (h, "h"), (dres, "dres")):
        _rows2d(t_, nm)
    rows, n = dy.shape
    d = h.shape[1]
    w_sets = 1
    if runs is not None:
        # ``runs`` = [(row_start, row_end, set)]: row ranges in ascending order, each with its own block of w ([sets · n, d]) and of
        # pack["s"] / pack["tb"] ([sets · n]); row tiles never straddle a run; col_partial gets one row per tile of every run
        if w.shape[0] % n or w.shape[1] != d or len(runs) == 0 or len(runs) > 8:
            raise ValueError("lnproj_bwd: with runs, w must be [sets * n, d] and 1..8 runs given")
        w_sets = w.shape[0] // n
        w_ok = True
    else:
        w_ok = w.shape == (n, d)
    if y.shape != (rows, n) or not w_ok or not w.is_contiguous() or h.shape[0] != rows or dres.shape != (rows, d):
        raise ValueError("lnproj_bwd: shapes do not match")
    if out is None:
        dx = torch.empty((rows, d), dtype=BF16, device=h.device)
    else:
        _need(out, BF16, "out", 2)
        _rows2d(out, "out")
        if out.shape != (rows, d):
            raise ValueError("lnproj_bwd: out must be [rows, d]")
        dx = out
    dz = torch.empty((rows, d), dtype=BF16, device=h.device) if want_dz else None
    p = L.LnProjBwdParams(dy=dy.data_ptr(), lddy=dy.stride(0), y=y.data_ptr(), ldy=y.stride(0), w=w.data_ptr(), s=pack["s"].data_ptr(),
                          tb=pack["tb"].data_ptr(), gamma=gamma.data_ptr(), h=h.data_ptr(), ldh=h.stride(0), mean=mean.data_ptr(),
                          rstd=rstd.data_ptr(), dres=dres.data_ptr(), lddres=dres.stride(0), dx=dx.data_ptr(), lddx=dx.stride(0),
                          dz=_ptr(dz), lddz=dz.stride(0) if dz is not None else 0, rows=rows, n=n, d=d)
    tiles = (rows + 127) // 128
    if runs is not None:
        tiles = sum((e - s_ + 127) // 128 for s_, e, _ in runs)
        p.num_runs, p.w_sets = len(runs), w_sets
        for i, (s_, e, k_) in enumerate(runs):
            p.run_start[i], p.run_start[i + 1], p.run_set[i] = s_, e, k_
            if i and s_ != runs[i - 1][1]:
                raise ValueError("lnproj_bwd: runs must be adjacent row ranges in ascending order")
    cols = torch.empty((3, tiles, d), dtype=F32, device=h.device) if want_cols else None
    p.col_partial = _ptr(cols)
    p.col_split = col_split
    dys = wpart = None
    if want_wgrad_operands:
        # ``dys`` = dy ⊙ rstd (bf16), ``wpart`` [2, 4·⌈rows/128⌉, n]: operands of ``lnproj_wgrad`` (the projection's weight / bias gradient)
        dys = torch.empty((rows, n), dtype=BF16, device=h.device)
        wpart = torch.empty((2, 4 * ((rows + 127) // 128), n), dtype=F32, device=h.device)
        p.dy_scaled, p.lddys, p.wgrad_partial = dys.data_ptr(), dys.stride(0), wpart.data_ptr()
    L.check(L.load().jl_lnproj_bwd(C.byref(p), _stream()))
    out = (dx, dz, cols) if want_cols else (dx, dz)
    return out + (dys, wpart) if want_wgrad_operands else out


def lnproj_wgrad_prep(dy: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor):
    """→ (dy ⊙ rstd bf16 [rows, n], wpart fp32 [2, 4·⌈rows/128⌉, n]): the operands of ``lnproj_wgrad`` from a kernel of their own."""
    _need(dy, BF16, "dy", 2)
    _rows2d(dy, "dy")
    rows, n = dy.shape
    dys = torch.empty((rows, n), dtype=BF16, device=dy.device)
    wpart = torch.empty((2, 4 * ((rows + 127) // 128), n), dtype=F32, device=dy.device)
    L.check(L.load().jl_lnproj_wgrad_prep(dy.data_ptr(), dy.stride(0), mean.data_ptr(), rstd.data_ptr(), rows, n, dys.data_ptr(), dys.stride(0),
                                          wpart.data_ptr(), _stream()))
    return dys, wpart


def lnproj_wgrad(m0: torch.Tensor, wpart: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dbias: Optional[torch.Tensor]) -> None:
    """Finish the weight gradient of a projection that follows a LayerNorm, in place: ``m0`` [n, d] fp32 = (dy ⊙ rstd)ᵀ · h (from
    ``gemm`` with MN-major operands) → dW = (m0 − v 1ᵀ) ⊙ γ + cs βᵀ; ``dbias`` [n] ← cs.  ``wpart`` from ``lnproj_bwd``."""
    _need(m0, F32, "m0", 2)
    _need(wpart, F32, "wpart", 3)
    n, d = m0.shape
    if wpart.shape[0] != 2 or wpart.shape[2] != n or not wpart.is_contiguous() or m0.stride(1) != 1:
        raise ValueError("lnproj_wgrad: shapes do not match")
    if dbias is not None:
        _need(dbias, F32, "dbias", 1)
        if dbias.numel() != n or not dbias.is_contiguous():
            raise ValueError("lnproj_wgrad: dbias must be a contiguous [n] tensor")
    L.check(L.load().jl_lnproj_wgrad(m0.data_ptr(), m0.stride(0), wpart.data_ptr(), wpart.shape[1], n, d, gamma.data_ptr(), beta.data_ptr(),
                                     _ptr(dbias), _stream()))


# ----------------------------------------------------------------------------------------------- AdapterFusion combine (f4)
def _fusion_params(y, q, key, alpha, scale):
    kk, rows, d = y.shape
    b = q.shape[1]
    for t, nm in ((y, "y"), (key, "key")):
        _need(t, BF16, nm, 3)
        if not t.is_contiguous():
            raise ValueError(f"fusion: {nm} must be contiguous [K, rows, C]")
    _need(q, BF16, "q", 2)
    _rows2d(q, "q")
    _need(alpha, F32, "alpha", 2)
    if key.shape != (kk, rows, b) or q.shape[0] != rows or alpha.shape != (rows, kk) or not alpha.is_contiguous():
        raise ValueError("fusion: shape mismatch between y / q / key / alpha")
    return L.FusionParams(y=y.data_ptr(), ldy=d, y_stride=rows * d, q=q.data_ptr(), ldq=q.stride(0), key=key.data_ptr(), ldkey=b,
                          key_stride=rows * b, alpha=alpha.data_ptr(), rows=rows, d=d, b=b, num_adapters=kk, scale=scale)


def fusion_combine_fwd(h: torch.Tensor, y: torch.Tensor, q: torch.Tensor, key: torch.Tensor, scale: float,
                       row_lengths: Optional[torch.Tensor] = None, rows_per_seq: int = 0):
    """out = h + Σ_k softmax_k(q · key_k · scale) y_k per row.  h [rows, d], y [K, rows, d], q [rows, b], key [K, rows, b] (bf16) →
    (out [rows, d] bf16, alpha [rows, K] fp32)."""
    _need(h, BF16, "h", 2)
    _rows2d(h, "h")
    rows, d = h.shape
    alpha = torch.empty((rows, y.shape[0]), dtype=F32, device=h.device)
    out = torch.empty((rows, d), dtype=BF16, device=h.device)
    p = _fusion_params(y, q, key, alpha, scale)
    if y.shape[1] != rows or y.shape[2] != d:
        raise ValueError("fusion: y must be [K, rows, d]")
    p.h, p.ldh, p.out, p.ldo = h.data_ptr(), h.stride(0), out.data_ptr(), out.stride(0)
    p.row_lengths, p.rows_per_seq = _ptr(row_lengths), rows_per_seq
    L.check(L.load().jl_fusion_combine_fwd(C.byref(p), _stream()))
    return out, alpha


def fusion_combine_bwd(dout: torch.Tensor, y: torch.Tensor, q: torch.Tensor, key: torch.Tensor, alpha: torch.Tensor, scale: float):
    """→ (dy [K, rows, d] = α_k · dout, dq [rows, b], dkey [K, rows, b]), all bf16; the caller adds dkey_k · W_k to dy_k."""
    _need(dout, BF16, "dout", 2)
    _rows2d(dout, "dout")
    kk, rows, d = y.shape
    b = q.shape[1]
    dy = torch.empty((kk, rows, d), dtype=BF16, device=y.device)
    dq = torch.empty((rows, b), dtype=BF16, device=y.device)
    dkey = torch.empty((kk, rows, b), dtype=BF16, device=y.device)
    p = _fusion_params(y, q, key, alpha, scale)
    p.dout, p.lddout = dout.data_ptr(), dout.stride(0)
    p.dy, p.lddy, p.dy_stride = dy.data_ptr(), d, rows * d
    p.dq, p.lddq = dq.data_ptr(), b
    p.dkey, p.lddkey, p.dkey_stride = dkey.data_ptr(), b, rows * b
    L.check(L.load().jl_fusion_combine_bwd(C.byref(p), _stream()))
    return dy, dq, dkey


# ----------------------------------------------------------------------------------------------- attention
def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, lengths: Optional[torch.Tensor], batch: int, seq: int, heads: int,
             scale: float, want_lse: bool = False, cu_seqlens: Optional[torch.Tensor] = None):
    """q/k/v: [B*seq, heads*64] views (unit inner stride, common row stride).  Returns (o [B*seq, heads*64] bf16, lse | None).
    Packed layout (``cu_seqlens`` [B + 1] int32): q/k/v/o are [total, heads*64], utterance b = rows [cu[b], cu[b+1]), ``seq`` is the
    upper bound on an utterance's length, lse is [heads, total]."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, BF16, nm)
        _rows2d(t, nm)
    if not (q.stride(0) == k.stride(0) == v.stride(0)):
        raise ValueError("attn: q, k, v must share one row stride")
    rows = q.shape[0]
    if cu_seqlens is None:
        if rows != batch * seq:
            raise ValueError(f"attn: q has {rows} rows, expected {batch * seq}")
    else:
        _need(cu_seqlens, I32, "cu_seqlens", 1)
        if cu_seqlens.numel() != batch + 1:
            raise ValueError(f"attn: cu_seqlens must have {batch + 1} entries")
    if q.shape[1] != heads * 64:
        raise ValueError(f"attn: q has shape {tuple(q.shape)}, expected {heads * 64} columns (head_dim is 64)")
    o = torch.empty((rows, heads * 64), dtype=BF16, device=q.device)
    lse = None
    if want_lse:
        lse = torch.empty((batch, heads, seq) if cu_seqlens is None else (heads, rows), dtype=F32, device=q.device)
    p = L.AttnFwdParams(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), ld_qkv=q.stride(0), o=o.data_ptr(), ld_o=o.stride(0),
                        lse=_ptr(lse), lengths=_ptr(lengths), batch=batch, seq=seq, heads=heads, scale=scale,
                        cu_seqlens=_ptr(cu_seqlens), total_rows=rows if cu_seqlens is not None else 0)
    L.check(L.load().jl_attn_fwd(C.byref(p), _stream()))
    return o, lse


def attn_bwd(q, k, v, o, d_o, lse, lengths, batch: int, seq: int, heads: int, scale: float, cu_seqlens: Optional[torch.Tensor] = None):
    """Returns dqkv [B*seq, 3*heads*64] bf16 (dq | dk | dv column blocks); [total, 3*heads*64] in the packed layout."""
    hd = heads * 64
    rows = q.shape[0]
    dqkv = torch.empty((rows, 3 * hd), dtype=BF16, device=q.device)
    delta = torch.empty((batch, heads, seq) if cu_seqlens is None else (heads, rows), dtype=F32, device=q.device)
    for t, nm in ((o, "o"), (d_o, "d_o")):
        _need(t, BF16, nm)
        _rows2d(t, nm)
    if o.stride(0) != d_o.stride(0):
        raise ValueError("attn_bwd: o and d_o must share one row stride")
    p = L.AttnBwdParams(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), ld_qkv=q.stride(0), o=o.data_ptr(), d_o=d_o.data_ptr(),
                        ld_o=o.stride(0), lse=lse.data_ptr(), dq=dqkv.data_ptr(), dk=dqkv[:, hd:].data_ptr(),
                        dv=dqkv[:, 2 * hd:].data_ptr(), ld_dqkv=dqkv.stride(0), delta=delta.data_ptr(), lengths=_ptr(lengths),
                        batch=batch, seq=seq, heads=heads, scale=scale, cu_seqlens=_ptr(cu_seqlens),
                        total_rows=rows if cu_seqlens is not None else 0)
    L.check(L.load().jl_attn_bwd(C.byref(p), _stream()))
    return dqkv


# ----------------------------------------------------------------------------------------------- CTC
def ctc_loss(logits: torch.Tensor, labels: torch.Tensor, input_lengths: torch.Tensor, blank: int = 0, reduction: str = "sum",
             zero_infinity: bool = False, want_grad: bool = False, grad_dtype=BF16, cu_seqlens: Optional[torch.Tensor] = None,
             max_len: int = 0):
    """logits [B, T, V] (fp32 | bf16, unit inner stride), labels [B, S] int32 (negative = pad), input_lengths [B] int32.
    Returns (loss [1] fp32, nll [B] fp32, grad [B, T, V] | None).  Packed layout: logits [total, V] with ``cu_seqlens`` [B + 1]
    and ``max_len`` >= every length; the gradient is then [total, V] too."""
    if logits.dtype not in (F32, BF16):
        raise TypeError("ctc_loss: logits must be fp32 or bf16")
    _need(labels, I32, "labels", 2)
    _need(input_lengths, I32, "input_lengths", 1)
    if reduction not in ("sum", "mean"):
        raise ValueError(f"ctc_loss: unsupported reduction {reduction!r}")
    if cu_seqlens is None:
        _need(logits, logits.dtype, "logits", 3)
        b, t, v = logits.shape
        if logits.stride(2) != 1 or logits.stride(0) != t * logits.stride(1):
            raise ValueError("ctc_loss: logits must be row-contiguous [B*T, V]")
        ld, gshape = logits.stride(1), (b, t, v)
    else:
        _need(logits, logits.dtype, "logits", 2)
        _need(cu_seqlens, I32, "cu_seqlens", 1)
        _rows2d(logits, "logits")
        b, t, v = input_lengths.numel(), int(max_len), logits.shape[1]
        if t <= 0 or cu_seqlens.numel() != b + 1:
            raise ValueError("ctc_loss: the packed layout needs max_len > 0 and cu_seqlens of B + 1 entries")
        ld, gshape = logits.stride(0), (logits.shape[0], v)
    labels = labels.contiguous()
    dev = logits.device
    nll = torch.empty((b,), dtype=F32, device=dev)
    loss = torch.empty((1,), dtype=F32, device=dev)
    grad = torch.empty(gshape, dtype=grad_dtype, device=dev) if want_grad else None
    p = L.CtcParams(logits=logits.data_ptr(), ld_logits=ld, logits_dtype=L.JL_DT_F32 if logits.dtype == F32 else L.JL_DT_BF16,
                    labels=labels.data_ptr(), max_label_len=labels.shape[1], input_lengths=input_lengths.data_ptr(), batch=b, seq=t,
                    vocab=v, blank=blank, reduction=L.JL_CTC_SUM if reduction == "sum" else L.JL_CTC_MEAN,
                    zero_infinity=1 if zero_infinity else 0, nll=nll.data_ptr(), loss=loss.data_ptr(), grad=_ptr(grad),
                    ld_grad=v, grad_dtype=L.JL_DT_F32 if grad_dtype == F32 else L.JL_DT_BF16, cu_seqlens=_ptr(cu_seqlens))
    lib = L.load()
    nbytes = C.c_size_t(0)
    L.check(lib.jl_ctc_workspace_bytes(C.byref(p), C.byref(nbytes)))
    ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=dev)
    L.check(lib.jl_ctc_fwd(C.byref(p), ws.data_ptr(), _stream()))
    return loss, nll, grad


def ctc_greedy(logits: torch.Tensor, input_lengths: torch.Tensor, blank: int = 0, cu_seqlens: Optional[torch.Tensor] = None,
               max_len: int = 0):
    """→ (out_ids [B, T] int32 with -1 tail, out_lengths [B] int32, frame_ids [B, T] int32).  Packed layout: logits [total, V] with
    ``cu_seqlens`` [B + 1] and T = ``max_len``."""
    if logits.dtype not in (F32, BF16):
        raise TypeError("ctc_greedy: logits must be fp32 or bf16")
    _need(input_lengths, I32, "input_lengths", 1)
    if cu_seqlens is None:
        _need(logits, logits.dtype, "logits", 3)
        b, t, v = logits.shape
        if logits.stride(2) != 1 or logits.stride(0) != t * logits.stride(1):
            raise ValueError("ctc_greedy: logits must be row-contiguous [B*T, V]")
        ld = logits.stride(1)
    else:
        _need(logits, logits.dtype, "logits", 2)
        _need(cu_seqlens, I32, "cu_seqlens", 1)
        _rows2d(logits, "logits")
        b, t, v = input_lengths.numel(), int(max_len), logits.shape[1]
        if t <= 0 or cu_seqlens.numel() != b + 1:
            raise ValueError("ctc_greedy: the packed layout needs max_len > 0 and cu_seqlens of B + 1 entries")
        ld = logits.stride(0)
    dev = logits.device
    frame_ids = torch.empty((b, t), dtype=I32, device=dev)
    out_ids = torch.empty((b, t), dtype=I32, device=dev)
    out_len = torch.empty((b,), dtype=I32, device=dev)
    p = L.CtcGreedyParams(logits=logits.data_ptr(), ld_logits=ld,
                          logits_dtype=L.JL_DT_F32 if logits.dtype == F32 else L.JL_DT_BF16, input_lengths=input_lengths.data_ptr(),
                          batch=b, seq=t, vocab=v, blank=blank, frame_ids=frame_ids.data_ptr(), out_ids=out_ids.data_ptr(),
                          out_lengths=out_len.data_ptr(), cu_seqlens=_ptr(cu_seqlens))
    L.check(L.load().jl_ctc_greedy(C.byref(p), _stream()))
    return out_ids, out_len, frame_ids


# ----------------------------------------------------------------------------------------------- helpers
def im2col_k5s2(x: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """[B, T, C] bf16 contiguous → ([B*T_out, 5*C] bf16, T_out) for Conv1d(k=5, s=2, p=2)."""
    _need(x, BF16, "x", 3)
    if not x.is_contiguous():
        raise ValueError("im2col: x must be contiguous")
    b, t, c = x.shape
    t_out = (t - 1) // 2 + 1
    out = torch.empty((b * t_out, 5 * c), dtype=BF16, device=x.device)
    L.check(L.load().jl_im2col_k5s2(x.data_ptr(), out.data_ptr(), b, t, c, t_out, _stream()))
    return out, t_out


def wave_stats(wave: torch.Tensor, num_samples: torch.Tensor) -> torch.Tensor:
    """wave [B, N] fp32, num_samples [B] int32 → stats [B, 2] fp32 = (mean, 1 / sqrt(var + 1e-7)) over the valid samples."""
    _need(wave, F32, "wave", 2)
    _need(num_samples, I32, "num_samples", 1)
    stats = torch.empty((wave.shape[0], 2), dtype=F32, device=wave.device)
    L.check(L.load().jl_wave_stats(wave.data_ptr(), wave.stride(0), num_samples.data_ptr(), wave.shape[0], wave.shape[1], stats.data_ptr(),
                                   _stream()))
    return stats


def wave_im2col(wave: torch.Tensor, num_samples: torch.Tensor, stats: torch.Tensor, t_out: int, kernel: int, stride: int,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Normalised waveform windows for Conv1d(1 → C, kernel, stride): → bf16 [B · t_out, 16]."""
    _need(wave, F32, "wave", 2)
    _need(stats, F32, "stats", 2)
    b = wave.shape[0]
    if out is None:
        out = torch.empty((b * t_out, 16), dtype=BF16, device=wave.device)
    L.check(L.load().jl_wave_im2col(wave.data_ptr(), wave.stride(0), num_samples.data_ptr(), b, wave.shape[1], stats.data_ptr(),
                                    out.data_ptr(), t_out, kernel, stride, _stream()))
    return out


def im2col_1d(x: torch.Tensor, t_out: int, kernel: int, stride: int, pad: int = 0, c0: int = 0, cg: Optional[int] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [B, T_in, C] bf16 → [B · t_out, kernel · cg] bf16 (tap-major) for Conv1d(kernel, stride, pad) over channels [c0, c0 + cg)."""
    _need(x, BF16, "x", 3)
    if not x.is_contiguous():
        raise ValueError("im2col_1d: x must be contiguous")
    b, t_in, c = x.shape
    cg = c if cg is None else cg
    if out is None:
        out = torch.empty((b * t_out, kernel * cg), dtype=BF16, device=x.device)
    L.check(L.load().jl_im2col_1d(x.data_ptr(), out.data_ptr(), b, t_in, c, t_out, kernel, stride, pad, c0, cg, _stream()))
    return out


def embed_positions_(h: torch.Tensor, scale: float, pos_table: torch.Tensor, lengths: torch.Tensor, batch: int, seq: int) -> torch.Tensor:
    _need(h, BF16, "h", 2)
    _need(pos_table, F32, "pos_table", 2)
    _need(lengths, I32, "lengths", 1)
    if not h.is_contiguous() or pos_table.shape[0] < seq + 2 or pos_table.shape[1] != h.shape[1] or not pos_table.is_contiguous():
        raise ValueError("embed_positions: bad shapes")
    L.check(L.load().jl_embed_positions(h.data_ptr(), scale, pos_table.data_ptr(), lengths.data_ptr(), batch, seq, h.shape[1], _stream()))
    return h


def embed_positions_packed(h: torch.Tensor, scale: float, pos_table: torch.Tensor, cu_seqlens: torch.Tensor, batch: int, seq: int,
                           total: int) -> torch.Tensor:
    """h [B·seq, d] bf16 (padded rows) → [total, d] bf16 packed rows: × scale + sinusoid row (t + 2); padded frames are dropped."""
    _need(h, BF16, "h", 2)
    _need(pos_table, F32, "pos_table", 2)
    _need(cu_seqlens, I32, "cu_seqlens", 1)
    if not h.is_contiguous() or h.shape[0] != batch * seq or pos_table.shape[0] < seq + 2 or pos_table.shape[1] != h.shape[1] or not pos_table.is_contiguous():
        raise ValueError("embed_positions_packed: bad shapes")
    out = torch.empty((total, h.shape[1]), dtype=BF16, device=h.device)
    L.check(L.load().jl_embed_positions_packed(h.data_ptr(), out.data_ptr(), scale, pos_table.data_ptr(), cu_seqlens.data_ptr(), batch, seq,
                                               h.shape[1], _stream()))
    return out


def transpose(x: torch.Tensor) -> torch.Tensor:
    _need(x, BF16, "x")
    _rows2d(x, "x")
    rows, cols = x.shape
    ld_out = (rows + 7) // 8 * 8
    out = torch.zeros((cols, ld_out), dtype=BF16, device=x.device) if ld_out != rows else torch.empty((cols, rows), dtype=BF16, device=x.device)
    L.check(L.load().jl_transpose_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols, _stream()))
    return out[:, :rows]


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need(x, BF16, "x")
    _rows2d(x, "x")
    rows, cols = x.shape
    out = torch.empty((cols,), dtype=F32, device=x.device) if out is None else out
    L.check(L.load().jl_colsum_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), rows, cols, None, _stream()))
    return out


def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need(x, F32, "x")
    if not x.is_contiguous():
        raise ValueError("cast_bf16: x must be contiguous")
    out = torch.empty(x.shape, dtype=BF16, device=x.device) if out is None else out
    if x.numel():
        L.check(L.load().jl_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()))
    return out


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need(a, BF16, "a")
    _need(b, BF16, "b")
    if not (a.is_contiguous() and b.is_contiguous()) or a.shape != b.shape:
        raise ValueError("add: operands must be contiguous and of equal shape")
    out = torch.empty_like(a) if out is None else out
    L.check(L.load().jl_add_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()))
    return out


def adamw_(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int, lr: float,
           beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.01, grad_scale: float = 1.0,
           param_bf16: Optional[torch.Tensor] = None, hyper_dev: Optional[torch.Tensor] = None) -> None:
    """Fused AdamW on a flat fp32 bucket (any contiguous slice of it).  ``hyper_dev``: fp32 CUDA tensor {lr, 1 - β1^t, sqrt(1 - β2^t)}
    read by the kernel instead of ``lr`` / ``step`` (graph-captured launches)."""
    for t, nm in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _need(t, F32, nm, 1)
        if not t.is_contiguous():
            raise ValueError(f"adamw: {nm} must be contiguous")
    p = L.AdamWParams(param=param.data_ptr(), grad=grad.data_ptr(), exp_avg=exp_avg.data_ptr(), exp_avg_sq=exp_avg_sq.data_ptr(),
                      param_bf16=_ptr(param_bf16), n=param.numel(), lr=lr, beta1=beta1, beta2=beta2, eps=eps,
                      weight_decay=weight_decay, grad_scale=grad_scale, step=step, hyper_dev=_ptr(hyper_dev))
    if hyper_dev is not None:
        _need(hyper_dev, F32, "hyper_dev", 1)
    L.check(L.load().jl_adamw_bucket(C.byref(p), _stream()))


def adamw_advance_(hyper: torch.Tensor) -> None:
    """hyper = {lr, bc1, bc2_sqrt, step, beta1, beta2} fp32 on the device: step += 1 and the bias corrections follow."""
    _need(hyper, F32, "hyper", 1)
    if hyper.numel() < 6 or not hyper.is_contiguous():
        raise ValueError("adamw_advance: hyper must be a contiguous fp32 tensor of 6 elements")
    L.check(L.load().jl_adamw_advance(hyper.data_ptr(), _stream()))
