"""GPU parity against the REFERENCE'S OWN CUDA KERNELS: the .cu files under /root/reference compiled unmodified for sm_100 into
oracle/_ref/libsamplenet_ref_cuda.so (oracle/Makefile, oracle/ref_cuda_shim.cu) and launched on identical inputs next to this library's
kernels -- north_star: "outputs match the reference's own TF/CUDA ops on identical inputs (kNN indices and match assignments bit-exact,
distances/losses within a stated fp32 tolerance)".

The reference kernels run on the legacy default stream; every call here is on torch's default stream.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import samplenet_b200

    samplenet_b200._lib.lib()
    return samplenet_b200


@pytest.fixture(scope="module")
def refcu():
    from oracle import ref_cuda

    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libsamplenet_ref_cuda.so not built (needs /root/reference at build time)")
    return ref_cuda


def _clouds(seed, b, n, m, noise=0.02):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, n, 3, generator=g) - 0.5
    if m <= n:
        q = x[:, torch.randperm(n, generator=g)[:m]] + noise * torch.randn(b, m, 3, generator=g)
    else:
        q = torch.rand(b, m, 3, generator=g) - 0.5
    return x.cuda().contiguous(), q.cuda().contiguous()


@pytest.mark.parametrize("b,n,m", [(32, 64, 1024), (4, 37, 129), (2, 513, 511), (32, 1024, 1024), (50, 2048, 2048), (3, 5, 2000)])
def test_chamfer_forward_equals_reference_kernels(sb, refcu, b, n, m):
    """registration ChamferDistanceKernel and TF NmDistanceKernel (same algorithm, two files): indices AND squared distances bit-identical
    to this library's kernel in its default (FMA-contracted, what nvcc gives the reference) arithmetic."""
    x, q = _clouds(b + n + m, b, m, n)          # xyz1 = q (b, n, 3), xyz2 = x (b, m, 3)
    d1, i1, d2, i2 = sb.ops.nn_distance_forward(q, x)
    for fn in (refcu.chamfer_forward, refcu.nn_distance):
        r1, j1, r2, j2 = fn(q, x)
        torch.cuda.synchronize()
        assert torch.equal(i1, j1) and torch.equal(i2, j2)
        assert torch.equal(d1, r1) and torch.equal(d2, r2)


def test_chamfer_backward_vs_reference_kernels(sb, refcu):
    x, q = _clouds(3, 8, 1024, 64)
    d1, i1, d2, i2 = sb.ops.nn_distance_forward(q, x)
    g = torch.Generator().manual_seed(5)
    g1 = torch.rand(d1.shape, generator=g).cuda(); g2 = torch.rand(d2.shape, generator=g).cuda()
    gx1, gx2 = sb.ops.nn_distance_backward(q, x, g1, i1, g2, i2)
    rx1, rx2 = refcu.chamfer_backward(q, x, g1, i1, g2, i2)     # float atomics: order-dependent rounding
    torch.cuda.synchronize()
    np.testing.assert_allclose(gx1.cpu().numpy(), rx1.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gx2.cpu().numpy(), rx2.cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("b,n,m,k", [(32, 1024, 64, 8), (32, 1024, 32, 7), (4, 2048, 64, 16), (3, 200, 17, 3), (2, 1024, 1024, 7)])
def test_knn_equals_reference_selection_sort(sb, refcu, b, n, m, k):
    """tf_grouping.knn_point = TF distance matrix + the reference's selection-sort kernel (tf_grouping_g.cu:83-123): neighbour indices
    bit-exact (tie-free inputs), squared distances bit-exact in this library's unfused arithmetic mode (three roundings, the order TF's
    elementwise graph evaluates)."""
    x, q = _clouds(b * 7 + k, b, n, m)
    val, idx = refcu.knn_point(k, x, q)
    o = sb.ops.knn_soft_project_forward(x, q, k, "bnc", want=("idx", "val"), unfused=True)
    torch.cuda.synchronize()
    assert torch.equal(o["idx"], idx)
    assert torch.equal(o["val"], val)
    # group_point on those indices
    gp = refcu.group_point(x, idx)
    ours = sb.tf_ops.group_point(x, idx)
    torch.cuda.synchronize()
    assert torch.equal(gp, ours)


@pytest.mark.parametrize("n,m", [(64, 64), (96, 32), (300, 300), (2048, 2048)])
def test_emd_vs_reference_kernels(sb, refcu, oracle, n, m):
    """approxmatch / matchcost / matchcostgrad of tf_approxmatch_g.cu (float, __expf, 512-thread tree reductions) on identical inputs.
    The reference GPU kernel is itself only an approximation of its CPU twin (its self-test flags |diff| > 1e-2, approxmatch.cpp:222);
    this library's fast kernel and its exact mode are both compared, and cost / gradients on IDENTICAL match."""
    b = 2 if n < 2048 else 1
    g = torch.Generator().manual_seed(n + m)
    a = torch.rand(b, n, 3, generator=g).cuda(); c = torch.rand(b, m, 3, generator=g).cuda()
    rm = refcu.approx_match(a, c)
    fast = sb.tf_ops.approx_match(a, c)
    torch.cuda.synchronize()
    assert float((fast - rm).abs().max()) < 5e-3
    if n <= 300:
        exact = sb.tf_ops.approx_match(a, c, exact=True)
        assert float((exact - rm).abs().max()) < 5e-3
        # assignments: equal wherever the reference's own top-2 gap exceeds that tolerance
        am, ar = exact.argmax(dim=2), rm.argmax(dim=2)
        gap = torch.gather(rm, 2, ar[..., None])[..., 0] - torch.gather(rm, 2, am[..., None])[..., 0]
        assert bool((gap < 5e-3).all())
    rc = refcu.match_cost(a, c, rm)
    oc = sb.ops.match_cost_forward(a, c, rm)
    torch.cuda.synchronize()
    np.testing.assert_allclose(oc.cpu().numpy(), rc.cpu().numpy(), rtol=2e-5)
    rg1, rg2 = refcu.match_cost_grad(a, c, rm)
    og1, og2 = sb.ops.match_cost_grad(a, c, rm)
    torch.cuda.synchronize()
    np.testing.assert_allclose(og1.cpu().numpy(), rg1.cpu().numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(og2.cpu().numpy(), rg2.cpu().numpy(), rtol=2e-4, atol=2e-5)
