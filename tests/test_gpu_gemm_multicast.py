"""GPU parity of the CTA-pair GEMM with B-operand TMA multicast across two pairs (clusters of four CTAs, gemm mode 3)."""
import pytest
import torch

from helpers import pkg, rel_err

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture
def mcast_mode():
    L = pkg()._lib
    L.load().jl_debug_set_gemm_mode(3)
    yield
    L.load().jl_debug_set_gemm_mode(0)


def _mk(m, n, k, seed, a_mn=False, b_mn=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(m, k, device="cuda", generator=g) * 0.5).to(BF16)
    b = (torch.randn(n, k, device="cuda", generator=g) * 0.5).to(BF16)
    return a, b, (a.t().contiguous() if a_mn else a), (b.t().contiguous() if b_mn else b)


@pytest.mark.parametrize("m,n,k", [(512, 256, 64), (1024, 256, 128), (8000, 2304, 768), (8000, 768, 3072), (8000, 3072, 768),
                                   (777, 200, 136), (1000, 5000, 768), (8000, 768, 64)])
def test_multicast_plain(mcast_mode, m, n, k):
    ops = pkg().ops
    a, b, _, _ = _mk(m, n, k, 31)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b, out_dtype=F32)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-3, (m, n, k)


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("m,n,k", [(1024, 256, 128), (8000, 768, 3072), (5000, 768, 8000)])
def test_multicast_mn_major(mcast_mode, m, n, k, a_mn, b_mn):
    ops = pkg().ops
    a, b, a_s, b_s = _mk(m, n, k, 32, a_mn, b_mn)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a_s, b_s, out_dtype=F32, a_layout=int(a_mn), b_layout=int(b_mn))
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 2e-3, (m, n, k, a_mn, b_mn)


def test_multicast_epilogue_and_speed():
    ops, L = pkg().ops, pkg()._lib
    m, n, k = 8000, 2304, 768
    a, b, _, _ = _mk(m, n, k, 33)
    bias = torch.randn(n, device="cuda")
    res = torch.randn(m, n, device="cuda").to(BF16)
    ref = a.float() @ b.float().t() + bias + res.float()
    times = {}
    for mode in (2, 3):
        L.load().jl_debug_set_gemm_mode(mode)
        out = ops.gemm(a, b, bias=bias, residual=res)
        torch.cuda.synchronize()
        assert rel_err(out.float(), ref) < 1e-2, mode
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm(a, b, bias=bias, residual=res, out=out)
        e1.record()
        torch.cuda.synchronize()
        times[mode] = e0.elapsed_time(e1) / 20
    L.load().jl_debug_set_gemm_mode(0)
    print("pair kernel %.1f us, multicast %.1f us" % (times[2] * 1e3, times[3] * 1e3))
