"""Trainer loss compositions (samplenet_b200/trainers.py) against the formulas of the reference trainers written out with basic ops."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_argument_checks_need_no_gpu():
    from samplenet_b200 import trainers
    with pytest.raises(ValueError):
        trainers.autoencoder_loss(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), loss="l2")
    with pytest.raises(ValueError):
        trainers.registration_samplenet_loss(None, None, None, 64, 0.01, 0.01, num_sampled_clouds=3)


@pytest.mark.gpu
def test_trainer_losses_match_reference_formulas():
    import __graft_entry__ as ge
    ge.build()
    import samplenet_b200 as sb
    from samplenet_b200 import trainers, tf_ops
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    p0 = (torch.rand(8, 256, 3, generator=g) - 0.5).cuda()
    p1 = (torch.rand(8, 256, 3, generator=g) - 0.5).cuda()
    cd = sb.ChamferDistance()

    def simp_loss(ref, samp, m, gamma, delta):   # registration/src/samplenet.py:171-181
        c12, c21 = cd(samp, ref)
        return c12.mean() + c12.max(dim=1)[0].mean() + (gamma + delta * m) * c21.mean()

    for nsc in (1, 2):
        net = sb.SampleNet(16, 128, group_size=4, input_shape="bnc", output_shape="bnc").cuda().train()
        ref_net = sb.SampleNet(16, 128, group_size=4, input_shape="bnc", output_shape="bnc").cuda().train()
        ref_net.load_state_dict(net.state_dict())
        loss, (q0, q1), info = trainers.registration_samplenet_loss(net, p0, p1, 16, 0.01, 0.1, gamma=1, delta=0.5, num_sampled_clouds=nsc)
        s1, pr1 = ref_net(p1)
        want = simp_loss(p1, s1, 16, 1, 0.5)
        if nsc == 2:
            s0, pr0 = ref_net(p0)
            want = 0.5 * (want + simp_loss(p0, s0, 16, 1, 0.5))
            np.testing.assert_allclose(q0.detach().cpu().numpy(), pr0.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
        else:
            assert q0 is p0
        np.testing.assert_allclose(float(info["simplification_loss"]), float(want), rtol=2e-5)
        np.testing.assert_allclose(float(loss), 0.01 * float(want) + 0.1 * float(ref_net.get_projection_loss()), rtol=2e-5)
        np.testing.assert_allclose(q1.detach().cpu().numpy(), pr1.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
        loss.backward()
        assert net.fc4.weight.grad is not None and float(net.fc4.weight.grad.abs().sum()) > 0

    samp = p1[:, :32].contiguous() + 0.01
    tot, ls = trainers.classification_total_loss(torch.tensor(2.0, device="cuda"), p1, samp, 32, 30.0, 1.0, 1, 0, torch.tensor(0.25, device="cuda"))
    np.testing.assert_allclose(float(ls), float(simp_loss(p1, samp, 32, 1, 0)), rtol=2e-5)
    np.testing.assert_allclose(float(tot), 2.0 + 30.0 * float(ls) + 0.25, rtol=1e-6)

    d1, _, d2, _ = tf_ops.nn_distance(samp, p1)
    np.testing.assert_allclose(float(trainers.autoencoder_loss(samp, p1, "chamfer")), float(d1.mean() + d2.mean()), rtol=1e-6)
    emd = trainers.autoencoder_loss(p0[:, :64].contiguous(), p1[:, :64].contiguous(), "emd")
    m = tf_ops.approx_match(p0[:, :64].contiguous(), p1[:, :64].contiguous())
    np.testing.assert_allclose(float(emd), float(tf_ops.match_cost(p0[:, :64].contiguous(), p1[:, :64].contiguous(), m).mean()), rtol=1e-6)
    for den in (False, True):
        l, dist, idx, dist2, per = trainers.autoencoder_simplification_loss(p1, samp, 32, den)
        w = 32 / 64.0 * (2 if den else 1)
        np.testing.assert_allclose(float(l), float(d1.mean() + d1.max(dim=1)[0].mean() + w * d2.mean()), rtol=2e-6)
        assert per.shape == (8, 1) and idx.dtype == torch.int32 and torch.equal(dist, d1) and torch.equal(dist2, d2)
    sizes = [8, 16, 32]
    prog = trainers.progressive_simplification_loss(p1, samp, sizes, 1, 0)
    np.testing.assert_allclose(float(prog), sum(float(simp_loss(p1, samp[:, :s].contiguous(), s, 1, 0)) for s in sizes), rtol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,m,sizes", [(4, 1024, 1024, [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]), (3, 500, 300, [7, 50, 299, 300]), (32, 1024, 1024, [8, 16, 32, 64, 128, 256, 512, 1024])])
def test_progressive_one_pass_equals_per_prefix(b, n, m, sizes):
    """csrc/progressive.cu: every prefix's Chamfer distances AND indices bit-identical to a stand-alone nn_distance on the sliced samples
    (the reference's graph: one NnDistance per prefix, train_samplenet_progressive.py:196-220); loss and gradients within fp32 tolerance."""
    import numpy as np
    import samplenet_b200 as sb
    from samplenet_b200 import ops, trainers

    g = torch.Generator().manual_seed(b + n + m)
    x = (torch.rand(b, n, 3, generator=g) - 0.5).cuda()
    s = (torch.rand(b, m, 3, generator=g) - 0.5).cuda()
    dist1, idx1, dist2, idx2, terms = ops.progressive_loss_forward(x, s, sizes, [1.0 + 0.01 * v for v in sizes])
    for p, sz in enumerate(sizes):
        d1, i1, d2, i2 = ops.nn_distance_forward(s[:, :sz].contiguous(), x)
        assert torch.equal(dist1[:, :sz], d1) and torch.equal(idx1[:, :sz], i1)
        assert torch.equal(dist2[:, p], d2) and torch.equal(idx2[:, p], i2)
    s1 = s.clone().requires_grad_(True); x1 = x.clone().requires_grad_(True)
    s2 = s.clone().requires_grad_(True); x2 = x.clone().requires_grad_(True)
    l1 = trainers.progressive_simplification_loss(x1, s1, sizes, gamma=1, delta=0.01, one_pass=True)
    l2 = trainers.progressive_simplification_loss(x2, s2, sizes, gamma=1, delta=0.01, one_pass=False)
    np.testing.assert_allclose(float(l1), float(l2), rtol=2e-6)
    np.testing.assert_allclose(float(terms[-1]), float(l2), rtol=2e-6)
    l1.backward(); l2.backward()
    np.testing.assert_allclose(s1.grad.cpu().numpy(), s2.grad.cpu().numpy(), rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(x1.grad.cpu().numpy(), x2.grad.cpu().numpy(), rtol=1e-4, atol=1e-8)
