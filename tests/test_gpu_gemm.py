"""GPU parity of the tcgen05 GEMM (through the C ABI) against fp32 matmul of the same bf16 operands: every tile width,
ragged M/N/K, both operand layouts, every epilogue."""
import math

import pytest
import torch

from helpers import pkg, rel_err

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


def _ops():
    return pkg().ops, pkg()._lib


def _mk(m, n, k, seed=0, a_mn=False, b_mn=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(m, k, device="cuda", generator=g) * 0.5).to(BF16)
    b = (torch.randn(n, k, device="cuda", generator=g) * 0.5).to(BF16)
    a_store = a.t().contiguous() if a_mn else a
    b_store = b.t().contiguous() if b_mn else b
    return a, b, a_store, b_store


SHAPES = [(128, 128, 64), (256, 256, 128), (128, 256, 768), (200, 136, 72), (1000, 768, 768), (777, 2304, 768),
          (300, 32, 768), (130, 256, 32), (64, 64, 8), (500, 5000, 768), (8000, 768, 3072), (250, 40, 128)]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_gemm_plain(m, n, k):
    ops, L = _ops()
    a, b, _, _ = _mk(m, n, k)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b, out_dtype=F32)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-3, (m, n, k)
    out16 = ops.gemm(a, b)
    assert rel_err(out16.float(), ref) < 1e-2


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("m,n,k", [(256, 256, 128), (200, 136, 72), (768, 32, 1000), (1000, 768, 5000), (5000, 768, 1000),
                                   (32, 256, 777), (777, 8, 128)])
def test_gemm_mn_major_operands(m, n, k, a_mn, b_mn):
    ops, L = _ops()
    if (a_mn and m % 8) or (b_mn and n % 8) or ((not a_mn or not b_mn) and k % 8):
        pytest.skip("row stride must be a multiple of 8 elements")
    a, b, a_s, b_s = _mk(m, n, k, seed=1, a_mn=a_mn, b_mn=b_mn)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a_s, b_s, out_dtype=F32, a_layout=int(a_mn), b_layout=int(b_mn))
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-3, (m, n, k, a_mn, b_mn)


@pytest.mark.parametrize("m,n,k,out_dtype", [(768, 64, 8000, F32), (192, 768, 8000, F32), (256, 32, 8000, BF16), (64, 192, 4104, F32)])
def test_gemm_split_k_weight_gradient_shapes(m, n, k, out_dtype):
    """dYᵀ·X shapes: few output tiles, long K → split-K with a fixed-order reduction (deterministic)."""
    ops, L = _ops()
    a, b, a_s, b_s = _mk(m, n, k, seed=11, a_mn=True, b_mn=True)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a_s, b_s, out_dtype=out_dtype, a_layout=1, b_layout=1)
    out2 = ops.gemm(a_s, b_s, out_dtype=out_dtype, a_layout=1, b_layout=1)
    torch.cuda.synchronize()
    assert rel_err(out.float(), ref) < (2e-3 if out_dtype == F32 else 1e-2)
    assert torch.equal(out, out2)


def test_gemm_matches_simt_reference_bitwise_epilogue():
    """Same epilogue code on a SIMT fp32-accumulate kernel: only the accumulation order differs."""
    ops, L = _ops()
    a, b, _, _ = _mk(300, 200, 136, seed=3)
    bias = torch.randn(200, device="cuda")
    x = ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GELU, out_dtype=F32)
    y = ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GELU, out_dtype=F32, reference=True)
    torch.cuda.synchronize()
    assert rel_err(x, y) < 1e-5


def test_gemm_epilogues():
    ops, L = _ops()
    m, n, k = 520, 264, 200
    a, b, _, _ = _mk(m, n, k, seed=5)
    g = torch.Generator(device="cuda").manual_seed(7)
    bias = torch.randn(n, device="cuda", generator=g)
    res = torch.randn(m, n, device="cuda", generator=g).to(BF16)
    aux = torch.randn(m, n, device="cuda", generator=g).to(BF16)
    acc = a.float() @ b.float().t()
    # bias + alpha
    out = ops.gemm(a, b, bias=bias, alpha=0.5, out_dtype=F32)
    assert rel_err(out, 0.5 * acc + bias) < 2e-3
    # GELU (erf) with saved pre-activation
    pre = torch.empty(m, n, dtype=BF16, device="cuda")
    out = ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GELU, aux_out=pre, out_dtype=F32)
    assert rel_err(out, torch.nn.functional.gelu(acc + bias)) < 2e-3
    assert rel_err(pre.float(), acc + bias) < 1e-2
    # ReLU + residual
    out = ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_RELU, residual=res, out_dtype=F32)
    assert rel_err(out, torch.relu(acc + bias) + res.float()) < 2e-3
    # GELU backward: acc * gelu'(aux)
    x = aux.float()
    gp = 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    out = ops.gemm(a, b, epilogue=L.JL_EPI_GELU_BWD, aux=aux, out_dtype=F32)
    assert rel_err(out, acc * gp) < 2e-3
    # training forward of the FFN input projection: GELU to C, gelu'(pre-activation) to aux_out; backward multiplies by it
    dact = torch.empty(m, n, dtype=BF16, device="cuda")
    out = ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GELU_DGELU, aux_out=dact, out_dtype=F32)
    xb = acc + bias
    assert rel_err(out, torch.nn.functional.gelu(xb)) < 2e-3
    gpb = 0.5 * (1 + torch.erf(xb / math.sqrt(2))) + xb * torch.exp(-0.5 * xb * xb) / math.sqrt(2 * math.pi)
    assert rel_err(dact.float(), gpb) < 5e-3
    assert float((dact.float() - gpb).abs().max()) < 1e-2          # bf16 rounding of a value in [-0.13, 1.13]
    out = ops.gemm(a, b, epilogue=L.JL_EPI_MUL_AUX, aux=aux, out_dtype=F32)
    assert rel_err(out, acc * x) < 2e-3
    with pytest.raises(L.JLError):
        ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GELU_DGELU)
    # ReLU backward
    out = ops.gemm(a, b, epilogue=L.JL_EPI_RELU_BWD, aux=aux, out_dtype=F32)
    assert rel_err(out, acc * (x > 0)) < 2e-3
    # GLU over interleaved (value, gate) columns
    out = ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GLU, out_dtype=F32)
    v = acc + bias
    assert out.shape == (m, n // 2)
    assert rel_err(out, v[:, 0::2] * torch.sigmoid(v[:, 1::2])) < 2e-3
    # row masking by utterance length
    rows_per_seq = 130
    lens = torch.tensor([130, 7, 0, 99], dtype=torch.int32, device="cuda")
    out = ops.gemm(a, b, bias=bias, residual=res, row_lengths=lens, rows_per_seq=rows_per_seq, out_dtype=F32)
    ref = acc + bias + res.float()
    t = torch.arange(m, device="cuda") % rows_per_seq
    valid = t < lens[torch.arange(m, device="cuda") // rows_per_seq]
    assert rel_err(out, ref * valid[:, None]) < 2e-3
    torch.cuda.synchronize()


def test_gemm_rejects_bad_arguments():
    ops, L = _ops()
    a = torch.zeros(16, 12, dtype=BF16, device="cuda")     # lda = 12, not a multiple of 8
    b = torch.zeros(16, 12, dtype=BF16, device="cuda")
    with pytest.raises(L.JLError):
        ops.gemm(a, b)
    with pytest.raises(ValueError):
        ops.gemm(torch.zeros(16, 16, dtype=BF16, device="cuda"), torch.zeros(16, 8, dtype=BF16, device="cuda"))


@pytest.fixture
def pair_mode():
    L = pkg()._lib
    L.load().jl_debug_set_gemm_mode(2)          # CTA-pair (cta_group::2) kernel wherever it is legal
    yield
    L.load().jl_debug_set_gemm_mode(0)


@pytest.mark.parametrize("m,n,k", [(256, 256, 128), (512, 128, 64), (8000, 768, 768), (8000, 2304, 768), (1000, 5000, 768),
                                   (777, 200, 136), (8000, 3072, 768), (300, 192, 3072)])
def test_gemm_cta_pair_kernel_plain(pair_mode, m, n, k):
    ops, L = _ops()
    a, b, _, _ = _mk(m, n, k, seed=21)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b, out_dtype=F32)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-3, (m, n, k)


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("m,n,k", [(512, 256, 128), (8000, 768, 3072), (5000, 768, 8000), (1000, 136, 200)])
def test_gemm_cta_pair_kernel_mn_major(pair_mode, m, n, k, a_mn, b_mn):
    ops, L = _ops()
    a, b, a_s, b_s = _mk(m, n, k, seed=22, a_mn=a_mn, b_mn=b_mn)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a_s, b_s, out_dtype=F32, a_layout=int(a_mn), b_layout=int(b_mn))
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-3, (m, n, k, a_mn, b_mn)


def test_gemm_cta_pair_kernel_epilogues(pair_mode):
    ops, L = _ops()
    m, n, k = 1100, 392, 200
    a, b, _, _ = _mk(m, n, k, seed=23)
    g = torch.Generator(device="cuda").manual_seed(7)
    bias = torch.randn(n, device="cuda", generator=g)
    res = torch.randn(m, n, device="cuda", generator=g).to(BF16)
    acc = a.float() @ b.float().t()
    pre = torch.empty(m, n, dtype=BF16, device="cuda")
    out = ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GELU, aux_out=pre, residual=res, out_dtype=F32)
    assert rel_err(out, torch.nn.functional.gelu(acc + bias) + res.float()) < 2e-3
    assert rel_err(pre.float(), acc + bias) < 1e-2
    out = ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GLU)
    v = acc + bias
    assert rel_err(out.float(), v[:, 0::2] * torch.sigmoid(v[:, 1::2])) < 1e-2
    lens = torch.tensor([275, 100, 0, 275], dtype=torch.int32, device="cuda")
    out = ops.gemm(a, b, row_lengths=lens, rows_per_seq=275, out_dtype=F32)
    t = torch.arange(m, device="cuda") % 275
    valid = t < lens[torch.arange(m, device="cuda") // 275]
    assert rel_err(out, acc * valid[:, None]) < 2e-3
    torch.cuda.synchronize()


def _tail_bytes(ops_mod, L, a, b, **kw):
    """Workspace the library asks for when the pair kernel would split its tail wave (0 = no split planned)."""
    import ctypes as C
    m, k = a.shape
    n = b.shape[0] if kw.get("b_layout", 0) == 0 else b.shape[1]
    out = torch.empty((m, n), dtype=BF16, device="cuda")
    bias, res = kw.get("bias"), kw.get("residual")
    p = L.GemmParams(a=a.data_ptr(), lda=a.stride(0), b=b.data_ptr(), ldb=b.stride(0), a_layout=0, b_layout=kw.get("b_layout", 0),
                     c=out.data_ptr(), ldc=n, bias=None if bias is None else bias.data_ptr(),
                     residual=None if res is None else res.data_ptr(), ldr=0 if res is None else res.stride(0),
                     m=m, n=n, k=k, epilogue=kw.get("epilogue", 0), out_dtype=L.JL_DT_BF16, alpha=1.0)
    nb, zb = C.c_size_t(0), C.c_size_t(0)
    L.check(L.load().jl_gemm_workspace_bytes(C.byref(p), C.byref(nb)))
    L.check(L.load().jl_gemm_workspace_zero_bytes(C.byref(p), C.byref(zb)))
    return nb.value, zb.value


@pytest.mark.parametrize("m,n,k,epi", [(8000, 768, 768, "bias_res"), (8000, 768, 3072, "bias_res"), (8000, 768, 2304, "plain"),
                                       (8000, 768, 768, "gelu"), (5000, 1024, 1024, "relu"), (2000, 768, 4096, "plain")])
def test_gemm_tail_split_matches_unsplit_and_reference(m, n, k, epi):
    """The partial last wave of the pair kernel is cut into K ranges (in-kernel fix-up by the last range to arrive): same
    result as the unsplit kernel up to fp32 summation order, equal to the fp32 reference within the usual tolerance,
    bit-identical from run to run, and the arrival counters are left zero."""
    ops, L = _ops()
    lib = L.load()
    a, b, _, _ = _mk(m, n, k, seed=3)
    g = torch.Generator(device="cuda").manual_seed(5)
    kw = {}
    if epi in ("bias_res", "gelu", "relu"):
        kw["bias"] = torch.randn(n, device="cuda", generator=g)
    if epi == "bias_res":
        kw["residual"] = torch.randn(m, n, device="cuda", generator=g).to(BF16)
    if epi == "gelu":
        kw["epilogue"] = L.JL_EPI_GELU
    if epi == "relu":
        kw["epilogue"] = L.JL_EPI_RELU
    lib.jl_debug_set_gemm_mode(2)          # pair kernel wherever legal (the automatic choice prefers 128 x 256 single-CTA tiles here)
    lib.jl_debug_set_gemm_tail(1)          # the tail split is off by default (measured slower)
    try:
        nbytes, zbytes = _tail_bytes(ops, L, a, b, **kw)
        assert nbytes > 0 and zbytes > 0, "this shape is expected to plan a tail split"
        out1 = ops.gemm(a, b, **kw).clone()
        out2 = ops.gemm(a, b, **kw).clone()
        torch.cuda.synchronize()
        assert torch.equal(out1, out2), "tail split must be bit-reproducible"
        ws = ops._TAIL_WS[(torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)]
        assert int(ws[:zbytes].to(torch.int32).abs().sum()) == 0, "arrival counters must be left zero"
        lib.jl_debug_set_gemm_tail(0)
        assert _tail_bytes(ops, L, a, b, **kw)[0] == 0
        unsplit = ops.gemm(a, b, **kw)
    finally:
        lib.jl_debug_set_gemm_tail(2)          # default: column slices in the single-CTA kernel only
        lib.jl_debug_set_gemm_mode(0)
    ref = ops.gemm(a, b, reference=True, **kw)
    torch.cuda.synchronize()
    assert rel_err(out1.float(), ref.float()) < 1e-2
    assert rel_err(out1.float(), unsplit.float()) < 4e-3
    z = a.float() @ b.float().t()
    if epi == "plain":
        assert rel_err(out1.float(), z) < 1e-2


@pytest.mark.parametrize("k,s,c,t", [(3, 2, 64, 400), (2, 2, 512, 1000), (3, 2, 512, 3001)])
def test_gemm_overlapping_a_rows_is_a_conv1d_without_im2col(k, s, c, t):
    """K-major A with lda < K: row r is the window of k consecutive frames starting at frame s·r of a [T, C] activation
    (rows overlap in memory) — the GEMM then equals Conv1d(C → N, k, stride s) over the frames."""
    ops, L = _ops()
    g = torch.Generator(device="cuda").manual_seed(11)
    x = (torch.randn(t + 8, c, device="cuda", generator=g) * 0.5).to(BF16)
    w = (torch.randn(96, k * c, device="cuda", generator=g) * 0.1).to(BF16)
    rows = (t - k) // s + 1
    a = torch.as_strided(x, (rows, k * c), (s * c, 1))
    out = ops.gemm(a, w, out_dtype=F32)
    ref = a.float() @ w.float().t()                      # as_strided materialises the windows on the torch side
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-3
    conv = torch.nn.functional.conv1d(x[:t].float().t()[None], w.float().view(96, k, c).permute(0, 2, 1), stride=s)[0].t()
    assert rel_err(out, conv) < 2e-3


@pytest.mark.parametrize("m,n,k,epi", [(8000, 768, 768, "bias_res"), (8000, 768, 3072, "plain"), (8000, 256, 512, "gelu"), (5000, 512, 1024, "relu"),
                                       (8000, 1000, 768, "plain")])
def test_gemm_tail_wave_as_column_slices(m, n, k, epi):
    """Single-CTA kernel: the partial last wave of 128 x BN tiles runs as 2 or 4 column slices per tile (independent work
    units, no partial sums): bit-identical to the unsliced schedule — every output element has the same K summation order."""
    ops, L = _ops()
    lib = L.load()
    a, b, _, _ = _mk(m, n, k, seed=4)
    g = torch.Generator(device="cuda").manual_seed(6)
    kw = {}
    if epi in ("bias_res", "gelu", "relu"):
        kw["bias"] = torch.randn(n, device="cuda", generator=g)
    if epi == "bias_res":
        kw["residual"] = torch.randn(m, n, device="cuda", generator=g).to(BF16)
    if epi == "gelu":
        kw["epilogue"] = L.JL_EPI_GELU
    if epi == "relu":
        kw["epilogue"] = L.JL_EPI_RELU
    lib.jl_debug_set_gemm_mode(1)              # single-CTA kernel
    try:
        for bn in (256, 128):
            lib.jl_debug_set_gemm_bn(bn)
            lib.jl_debug_set_gemm_tail(2)
            sliced = ops.gemm(a, b, **kw).clone()
            lib.jl_debug_set_gemm_tail(0)
            whole = ops.gemm(a, b, **kw).clone()
            torch.cuda.synchronize()
            assert torch.equal(sliced, whole), (bn, float((sliced.float() - whole.float()).abs().max()))
        ref = ops.gemm(a, b, reference=True, **kw)
        assert rel_err(sliced.float(), ref.float()) < 1e-2
    finally:
        lib.jl_debug_set_gemm_tail(2)
        lib.jl_debug_set_gemm_bn(0)
        lib.jl_debug_set_gemm_mode(0)
