"""GPU tests of the tcgen05 tensor-core linear (csrc/tc_gemm.cu) against fp64: 3xTF32 must be fp32-class accurate."""
import pytest
import torch

from point_unet_b200 import ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("M,K,N", [(4096, 64, 64), (1000, 32, 32), (5000, 128, 128), (3000, 256, 512), (777, 1536, 512),
                                   (2000, 64, 32), (130, 96, 40), (128, 32, 128), (100000, 64, 64), (300, 44, 36)])
@pytest.mark.parametrize("mode", [3, 1])
def test_tc_linear_matches_fp64(M, K, N, mode):
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g) + 0.3
    w = torch.randn(K, N, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    want = x.double() @ w.double() + b.double()
    xg, wg, bg = x.cuda(), w.cuda(), b.cuda()
    ops.tc_error_flag(xg.device).zero_()
    y, mean, var = ops.linear_raw(xg, wg, bg, want_stats=True, tc_mode=mode)
    assert int(ops.tc_error_flag(xg.device).item()) == 0, "tcgen05 pipeline barrier timed out"
    tol = 2e-5 if mode == 3 else 3e-3  # 3xTF32: ~2e-6 at K<=256, grows ~sqrt(K) (8e-6 at K=1536)
    assert rel(y, want) < tol, (rel(y, want), tol)
    assert rel(mean, want.mean(0)) < max(tol, 1e-5) * 5
    assert rel(var, want.var(0, unbiased=False)) < max(tol, 1e-5) * 5
    # accumulate into an existing buffer, strided output (half of a concat buffer), K-major weight passed directly
    buf = torch.randn(M, 2 * N, generator=g).cuda()
    want2 = buf[:, N:].double().cpu() + x.double() @ w.double()
    ops.linear_raw(xg, None, None, out=buf[:, N:], accumulate=True, wt=wg.t().contiguous(), tc_mode=mode)
    assert rel(buf[:, N:], want2) < tol * 2
    # same numbers as the CUDA-core path within fp32 rounding when mode == 3
    if mode == 3:
        y0 = ops.linear_raw(xg, wg, bg, tc_mode=0)
        assert rel(y, y0) < 2e-5


def test_tc_used_by_autograd_linear_and_att_pool():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 256, 16, 64, generator=g)
    w = torch.randn(64, 64, generator=g) * 0.2
    dy = torch.randn(2, 256, 1, 64, generator=g)
    outs = []
    for mode in (0, 3):
        ops.TC_MODE = mode
        xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
        agg = ops.att_pool(xg, wg)
        (agg * dy.cuda()).sum().backward()
        outs.append((agg.detach(), xg.grad.clone(), wg.grad.clone()))
    ops.TC_MODE = 3
    for a, b in zip(outs[0], outs[1]):
        assert rel(b, a) < 2e-5


@pytest.mark.parametrize("M,K,N", [(8192, 64, 64), (50000, 32, 32), (20000, 128, 128), (9000, 256, 512), (5000, 1536, 512),
                                   (100000, 64, 32), (4100, 96, 40), (300000, 32, 64), (6000, 160, 32), (2812, 512, 1024), (1404, 1024, 1024),
                                   (703, 1536, 512), (600, 64, 64)])
@pytest.mark.parametrize("mode", [3, 1])
def test_tc_wgrad_matches_fp64(M, K, N, mode):
    g = torch.Generator().manual_seed(M + K + N + 1)
    x = torch.randn(M, K, generator=g) + 0.2
    dy = torch.randn(M, N, generator=g)
    want = x.double().t() @ dy.double()
    want_db = dy.double().sum(0)
    xg, dg = x.cuda(), dy.cuda()
    ops.tc_error_flag(xg.device).zero_()
    with_db = (K % 128) != 0
    dw, db = ops.wgrad_raw(xg, dg, want_db=with_db, tc_mode=mode)
    assert int(ops.tc_error_flag(xg.device).item()) == 0, "tcgen05 pipeline barrier timed out"
    tol = 2e-5 if mode == 3 else 5e-3
    assert rel(dw, want) < tol, (rel(dw, want), tol)
    if with_db:
        assert rel(db, want_db) < tol
    dw0, _ = ops.wgrad_raw(xg, dg, tc_mode=0)
    assert rel(dw, dw0) < (3e-5 if mode == 3 else 5e-3)
    # strided operands (halves of wider buffers) and bit-determinism
    wide = torch.zeros(M, 2 * K, device="cuda"); wide[:, K:] = xg
    dw2, _ = ops.wgrad_raw(wide[:, K:], dg, tc_mode=mode)
    assert torch.equal(dw2, dw)


@pytest.mark.gpu
def test_tc_linear_backed_up_pipeline_is_race_free():
    """Accumulate mode makes the epilogue the slowest stage, so the TMA ring runs full and every slot is refilled the moment
    its release barrier completes -- the regime in which a premature release shows up as stale 8-row groups."""
    g = torch.Generator().manual_seed(7)
    M, K, N = 100000, 64, 64
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(K, N, generator=g) / 8).cuda()
    wt = w.t().contiguous()
    prod = x.double() @ w.double()
    for trial in range(4):
        buf = torch.randn(M, 2 * N, generator=g).cuda()
        want = buf[:, N:].double() + prod
        ops.tc_error_flag(x.device).zero_()
        ops.linear_raw(x, None, None, out=buf[:, N:], accumulate=True, wt=wt, tc_mode=3)
        assert int(ops.tc_error_flag(x.device).item()) == 0
        err = (buf[:, N:].double() - want).abs().max().item()
        assert err < 1e-4, (trial, err)
