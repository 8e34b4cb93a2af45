"""Generate golden fixtures by running the REFERENCE's own Python classes, imported unmodified from
/root/reference/registration/src, on CPU in this (GPU-less) container.

    python tests/golden/make_golden.py        # writes tests/golden/*.npz

The reference imports three un-vendored packages (two CUDA-only) at module import time; they are replaced in
sys.modules by exact CPU stand-ins that supply ONLY the boundary functions:
  knn_cuda.KNN(k, transpose_mode=False)(ref (B,3,N), query (B,3,M)) -> (dist (B,k,M), idx (B,k,M) int64)
      brute force, sorted by (squared distance, index)            [registration/src/soft_projection.py:11-14]
  pointnet2.utils.pointnet2_utils.grouping_operation(features (B,C,N), idx (B,M,k) int32) -> (B,C,M,k)
      torch.gather (autograd gives the scatter-add backward)      [registration/src/soft_projection.py:86-88]
The reference Chamfer extension (chamfer_distance.{cpp,cu}) is JIT-built unmodified by the reference's own
chamfer_distance.py (torch.utils.cpp_extension.load) and its CPU entry points are used.
Everything else -- SampleNet, SoftProjection, ChamferDistance, the losses, autograd -- is reference code.

The fixtures cannot be regenerated on the GPU box (no /root/reference there); they are committed.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/registration"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    knn_cuda = types.ModuleType("knn_cuda")

    class KNN:
        def __init__(self, k, transpose_mode=False):
            self.k = k
            self.transpose_mode = transpose_mode

        def __call__(self, ref, query):
            assert not self.transpose_mode
            r = ref.detach().permute(0, 2, 1)  # B,N,3
            q = query.detach().permute(0, 2, 1)  # B,M,3
            diff = r[:, None, :, :] - q[:, :, None, :]  # B,M,N,3
            d = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
            dn = d.numpy()
            # stable argsort == sorted by (distance, index)
            order = np.argsort(dn, axis=2, kind="stable")[:, :, : self.k]
            idx = torch.from_numpy(order.astype(np.int64))
            dist = torch.gather(d, 2, idx).sqrt()
            return dist.permute(0, 2, 1).contiguous(), idx.permute(0, 2, 1).contiguous()

    knn_cuda.KNN = KNN
    sys.modules["knn_cuda"] = knn_cuda

    pn2 = types.ModuleType("pointnet2")
    pn2u = types.ModuleType("pointnet2.utils")
    pn2uu = types.ModuleType("pointnet2.utils.pointnet2_utils")

    def grouping_operation(features, idx):
        B, C, N = features.shape
        _, M, K = idx.shape
        ii = idx.long().reshape(B, 1, M * K).expand(B, C, M * K)
        return torch.gather(features, 2, ii).reshape(B, C, M, K)

    pn2uu.grouping_operation = grouping_operation

    def _not_on_this_path(*a, **k):  # imported by src/fps.py:4-5, src/random_sampling.py:4; never called here
        raise NotImplementedError

    pn2uu.furthest_point_sample = _not_on_this_path
    pn2uu.gather_operation = _not_on_this_path
    pn2.utils = pn2u
    pn2u.pointnet2_utils = pn2uu
    sys.modules["pointnet2"] = pn2
    sys.modules["pointnet2.utils"] = pn2u
    sys.modules["pointnet2.utils.pointnet2_utils"] = pn2uu

    # kornia (unpinned, registration/Dockerfile:6) is imported by src/qdataset.py:4-5; the only function on the registration step's path
    # is conversions.quaternion_to_rotation_matrix((x, y, z, w)) (qdataset.py:74-75): restated here as kornia <= 0.4 defines it
    # (normalise, then the standard unit-quaternion rotation matrix).  PARITY UNPINNED for this one function (source not in the tree).
    for name in ("kornia", "kornia.geometry", "kornia.geometry.conversions", "kornia.geometry.linalg"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["kornia"].geometry = sys.modules["kornia.geometry"]
    sys.modules["kornia.geometry"].conversions = sys.modules["kornia.geometry.conversions"]
    sys.modules["kornia.geometry"].linalg = sys.modules["kornia.geometry.linalg"]

    def quaternion_to_rotation_matrix(quaternion):
        q = torch.nn.functional.normalize(quaternion, p=2, dim=-1, eps=1e-12)
        x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
        tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
        twx, twy, twz = tx * w, ty * w, tz * w
        txx, txy, txz = tx * x, ty * x, tz * x
        tyy, tyz, tzz = ty * y, tz * y, tz * z
        one = torch.ones_like(x)
        m = torch.stack([one - (tyy + tzz), txy - twz, txz + twy,
                         txy + twz, one - (txx + tzz), tyz - twx,
                         txz - twy, tyz + twx, one - (txx + tyy)], dim=-1)
        return m.view(quaternion.shape[:-1] + (3, 3))

    sys.modules["kornia.geometry.conversions"].quaternion_to_rotation_matrix = quaternion_to_rotation_matrix
    # registration/main.py also imports h5py (through data/modelnet_loader_torch.py:12; never called here)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))


def unit_cube(x):
    """OnUnitCube.method2 (registration/src/pctransforms.py:162-166) per cloud: divide by the largest
    axis extent, then subtract the mean."""
    c = x.max(dim=1)[0] - x.min(dim=1)[0]  # B,3
    s = c.max(dim=1)[0].view(-1, 1, 1)
    v = x / s
    return v - v.mean(dim=1, keepdim=True)


def main():
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/samplenet_ref_torch_ext")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    install_stubs()
    sys.path.insert(0, REF)
    from src.samplenet import SampleNet  # noqa: E402  (reference, unmodified)
    from src.soft_projection import SoftProjection  # noqa: E402
    from src.chamfer_distance import ChamferDistance  # noqa: E402

    torch.set_num_threads(4)

    # ---------------- fixture 1: config 0 — SampleNet fwd + soft-proj + both losses, B=2, N=1024->64, k=8
    torch.manual_seed(0)
    net = SampleNet(64, 128, group_size=8, initial_temperature=1.0, input_shape="bnc", output_shape="bnc")
    # make BN affine params and temperature non-trivial so that parity exercises them
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.startswith("bn") and name.endswith("weight"):
                p.copy_(1.0 + 0.25 * torch.randn(p.shape, generator=g))
            if name.startswith("bn") and name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
        net.project._temperature.fill_(0.35)
    state0 = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    gx = torch.Generator().manual_seed(0)
    x = unit_cube(torch.rand(2, 1024, 3, generator=gx) - 0.5)
    net.train()
    simp, proj = net(x)
    simp.retain_grad()
    loss_s = net.get_simplification_loss(x, simp, 64, 1, 0)
    loss_p = net.get_projection_loss()
    rw = torch.randn(proj.shape, generator=torch.Generator().manual_seed(2))
    total = 0.01 * loss_s + 0.01 * loss_p + (proj * rw).sum()
    total.backward()
    grads = {
        "grad_fc4_bias": net.fc4.bias.grad.numpy(),
        "grad_fc4_weight_row0": net.fc4.weight.grad[0].numpy(),
        "grad_conv1_weight": net.conv1.weight.grad.numpy(),
        "grad_conv5_bias": net.conv5.bias.grad.numpy(),
        "grad_bn3_weight": net.bn3.weight.grad.numpy(),
        "grad_temperature": net.project._temperature.grad.numpy(),
        "grad_simp": simp.grad.numpy(),
    }
    # fp64 evaluation of the same generator (reference module cast to double): the yardstick for fp32 noise.  With B=2 the
    # BatchNorm over the batch in the FC head is ill-conditioned, fp32 results scatter ~3e-4 around this.
    import copy
    net64 = copy.deepcopy(net).double()
    net64.load_state_dict({k: (torch.from_numpy(v).double() if torch.from_numpy(v).is_floating_point() else torch.from_numpy(v)) for k, v in state0.items()})
    net64.train()
    net64.skip_projection = True
    with torch.no_grad():
        simp64, _ = net64(x.double())
    state1 = net.state_dict()
    run_stats = {("after_" + k): v.detach().numpy() for k, v in state1.items() if "running" in k or "num_batches" in k}
    np.savez_compressed(
        os.path.join(OUT, "samplenet_reg_b2.npz"),
        x=x.numpy(), simp_fp64=simp64.numpy(), simp=simp.detach().numpy(), proj=proj.detach().numpy(), rw=rw.numpy(),
        loss_simplification=loss_s.detach().numpy(), loss_projection=loss_p.detach().numpy(),
        **{("sd_" + k): v for k, v in state0.items()}, **grads, **run_stats,
    )

    # eval-mode forward of the same net would call .cuda() (samplenet.py:141); restate its steps up to there
    # (same reference functions: KNN(1) + sputils.nn_matching) so that the matched output is pinned too.
    from src import sputils  # noqa: E402
    net.eval()
    with torch.no_grad():
        xb = x.permute(0, 2, 1)
        import torch.nn.functional as F
        y = F.relu(net.bn1(net.conv1(xb))); y = F.relu(net.bn2(net.conv2(y))); y = F.relu(net.bn3(net.conv3(y)))
        y = F.relu(net.bn4(net.conv4(y))); y = F.relu(net.bn5(net.conv5(y)))
        y = torch.max(y, 2)[0]
        y = F.relu(net.bn_fc1(net.fc1(y))); y = F.relu(net.bn_fc2(net.fc2(y))); y = F.relu(net.bn_fc3(net.fc3(y)))
        y = net.fc4(y).view(-1, 3, 64)
        _, idx = sys.modules["knn_cuda"].KNN(1, transpose_mode=False)(xb.contiguous(), y.contiguous())
        idx = np.squeeze(idx.numpy(), axis=1)
        z = sputils.nn_matching(x.numpy(), idx, 64, complete_fps=True)
    np.savez_compressed(os.path.join(OUT, "samplenet_reg_b2_eval.npz"), simp_eval=y.permute(0, 2, 1).numpy(),
                        nn_idx=idx.astype(np.int32), match=z)

    # ---------------- fixture 2: SoftProjection project / propagate / project_and_propagate + grads
    torch.manual_seed(3)
    B, N, M, K, Fd = 3, 200, 17, 8, 5
    pc = torch.randn(B, 3, N, requires_grad=True)
    qc = (pc.detach()[:, :, torch.randperm(N)[:M]] + 0.05 * torch.randn(B, 3, M)).requires_grad_(True)
    feats = torch.randn(B, Fd, N, requires_grad=True)
    sp = SoftProjection(K, initial_temperature=0.7, is_temperature_trainable=True, min_sigma=1e-4)
    pp, pf = sp(pc, qc, feats, action="project_and_propagate")
    r1 = torch.randn(pp.shape); r2 = torch.randn(pf.shape)
    ((pp * r1).sum() + (pf * r2).sum()).backward()
    only_proj = sp(pc.detach(), qc.detach())
    only_prop = sp(pc.detach(), qc.detach(), feats.detach(), action="propagate")
    np.savez_compressed(
        os.path.join(OUT, "softproj_reg.npz"),
        point_cloud=pc.detach().numpy(), query_cloud=qc.detach().numpy(), feats=feats.detach().numpy(),
        temperature=np.float32(0.7), min_sigma=np.float32(1e-4), k=np.int32(K),
        proj=pp.detach().numpy(), prop=pf.detach().numpy(), only_proj=only_proj.detach().numpy(),
        only_prop=only_prop.detach().numpy(), r1=r1.numpy(), r2=r2.numpy(),
        grad_point_cloud=pc.grad.numpy(), grad_query_cloud=qc.grad.numpy(), grad_feats=feats.grad.numpy(),
        grad_temperature=sp._temperature.grad.numpy(),
    )

    # ---------------- fixture 3: ChamferDistance fwd/bwd through the reference autograd Function
    torch.manual_seed(4)
    a = torch.randn(4, 37, 3, requires_grad=True)
    b = torch.randn(4, 129, 3, requires_grad=True)
    d1, d2 = ChamferDistance()(a, b)
    w1 = torch.rand(d1.shape); w2 = torch.rand(d2.shape)
    ((d1 * w1).sum() + (d2 * w2).sum()).backward()
    np.savez_compressed(os.path.join(OUT, "chamfer_reg.npz"), xyz1=a.detach().numpy(), xyz2=b.detach().numpy(),
                        dist1=d1.detach().numpy(), dist2=d2.detach().numpy(), w1=w1.numpy(), w2=w2.numpy(),
                        grad_xyz1=a.grad.numpy(), grad_xyz2=b.grad.numpy())
    headline_and_registration(SampleNet, state0)
    print("golden fixtures written to", OUT)


def perturbed_samplenet(SampleNet, m=64, temperature=0.35):
    """The SampleNet of fixture 1: default init under torch.manual_seed(0), BN affine parameters perturbed with Generator(seed 1)."""
    torch.manual_seed(0)
    net = SampleNet(m, 128, group_size=8, initial_temperature=1.0, input_shape="bnc", output_shape="bnc")
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.startswith("bn") and name.endswith("weight"):
                p.copy_(1.0 + 0.25 * torch.randn(p.shape, generator=g))
            if name.startswith("bn") and name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
        net.project._temperature.fill_(temperature)
    return net


def headline_and_registration(SampleNet, state0):
    import copy

    # ---------------- fixture 4: the HEADLINE size -- B=32, N=1024 -> 64, k=8: forward, both losses, every parameter gradient.
    # The weights are those of fixture 1 (same seeds; asserted), so only the batch and the outputs are stored.
    net = perturbed_samplenet(SampleNet)
    for k, v in net.state_dict().items():
        assert np.array_equal(v.detach().numpy(), state0[k]), k
    x = unit_cube(torch.rand(32, 1024, 3, generator=torch.Generator().manual_seed(5)) - 0.5)
    net.train()
    simp, proj = net(x)
    simp.retain_grad()
    loss_s = net.get_simplification_loss(x, simp, 64, 1, 0)
    loss_p = net.get_projection_loss()
    rw = torch.randn(proj.shape, generator=torch.Generator().manual_seed(6))
    (0.01 * loss_s + 0.01 * loss_p + (proj * rw).sum()).backward()
    gnorm = {("gnorm_" + k): np.float64(p.grad.double().norm().item()) for k, p in net.named_parameters()}
    net64 = copy.deepcopy(net).double()
    net64.load_state_dict({k: (torch.from_numpy(v).double() if torch.from_numpy(v).is_floating_point() else torch.from_numpy(v)) for k, v in state0.items()})
    net64.train()
    with torch.no_grad():
        simp64, proj64 = net64(x.double())
        d64 = ((simp64[:, :, None, :] - x.double()[:, None, :, :]) ** 2).sum(-1)   # (the reference Chamfer extension is float-only)
        c12, c21 = d64.min(dim=2)[0], d64.min(dim=1)[0]
        loss64 = c12.mean() + c12.max(dim=1)[0].mean() + c21.mean()
    np.savez_compressed(
        os.path.join(OUT, "samplenet_reg_b32.npz"),
        x=x.numpy(), simp=simp.detach().numpy(), proj=proj.detach().numpy(), rw=rw.numpy(), simp_fp64=simp64.numpy(), proj_fp64=proj64.numpy(),
        loss_simplification=loss_s.detach().numpy(), loss_projection=loss_p.detach().numpy(), loss_simplification_fp64=loss64.numpy(),
        grad_fc4_bias=net.fc4.bias.grad.numpy(), grad_conv1_weight=net.conv1.weight.grad.numpy(), grad_conv5_bias=net.conv5.bias.grad.numpy(),
        grad_bn3_weight=net.bn3.weight.grad.numpy(), grad_fc2_weight_rows=net.fc2.weight.grad[:4].numpy(), grad_conv4_weight_rows=net.conv4.weight.grad[:4].numpy(),
        grad_temperature=net.project._temperature.grad.numpy(), grad_simp=simp.grad.numpy(),
        after_bn5_running_mean=net.bn5.running_mean.numpy(), after_bn5_running_var=net.bn5.running_var.numpy(),
        after_bn_fc3_running_var=net.bn_fc3.running_var.numpy(), **gnorm,
    )

    # ---------------- fixture 5: one registration training step's loss assembly through the reference's own `Action`
    # (registration/main.py:221-247, 500-598): compute_samplenet_loss + compute_pcrnet_loss on a synthetic (template, source, igt) batch.
    # PCRNet (4.4 M parameters) is default-initialised under torch.manual_seed(11); a consumer rebuilds it from the same seed.
    sys.modules.setdefault("data", types.ModuleType("data"))
    mlt = types.ModuleType("data.modelnet_loader_torch")
    mlt.ModelNetCls = object
    sys.modules["data.modelnet_loader_torch"] = mlt
    argv_keep = sys.argv
    sys.argv = ["main.py"]
    import main as refmain  # noqa: E402  (registration/main.py, unmodified)
    sys.argv = argv_keep
    from src import sputils as ref_sputils  # noqa: E402
    from src.qdataset import QuaternionTransform  # noqa: E402
    import src.quaternion as Q  # noqa: E402

    args = refmain.options(["-o", "/tmp/x", "--datafolder", "none", "--sampler", "samplenet", "--train-samplenet", "--num-sampled-clouds", "2",
                            "--device", "cpu"], parser=ref_sputils.get_parser())
    for ncl in (1, 2):
        args.num_sampled_clouds = ncl
        act = refmain.Action(args)
        torch.manual_seed(11)
        model = act.create_model()
        model.sampler.load_state_dict({k: torch.from_numpy(v) for k, v in state0.items()})
        model.sampler.train()
        gb = torch.Generator().manual_seed(12)
        B = 4
        p0 = unit_cube(torch.rand(B, 1024, 3, generator=gb) - 0.5)
        rot = (torch.rand(B, 3, generator=gb) - 0.5) * (np.pi / 2)
        quat = torch.from_numpy(Q.euler_to_quaternion(rot.numpy(), "xyz").astype(np.float32))
        vec = torch.cat([quat, torch.zeros(B, 3)], dim=1)
        igt = {"vec": vec, "inversion": torch.tensor([False])}
        p1 = QuaternionTransform(vec).rotate(p0)
        data = (p0, p1, igt)
        sl, sampled, info = act.compute_samplenet_loss(model, data, "cpu")
        pl, pinfo = act.compute_pcrnet_loss(model, sampled, "cpu", 0)
        cons = act.compute_sampling_consistency(sampled, "cpu")
        if ncl == 2:   # the whole step of train_1 (main.py:340-352): task loss + sampling losses, backward into the sampler
            model.zero_grad()
            (pl + sl).backward()
            gn = {("gnorm_" + k): np.float64(p.grad.double().norm().item()) for k, p in model.sampler.named_parameters()}
        else:
            gn = {}
        np.savez_compressed(
            os.path.join(OUT, "registration_step_c%d.npz" % ncl), p0=p0.numpy(), p1=p1.detach().numpy(), igt_vec=vec.numpy(),
            samplenet_loss=sl.detach().numpy(), simplification_loss=info["simplification_loss"].detach().numpy(),
            projection_loss=info["projection_loss"].detach().numpy(), p0_out=sampled[0].detach().numpy(), p1_out=sampled[1].detach().numpy(),
            pcrnet_loss=pl.detach().numpy(), chamfer_loss=pinfo["chamfer_loss"].detach().numpy(), qnorm_loss=pinfo["qnorm_loss"].detach().numpy(),
            rot_err=np.float32(pinfo["rot_err"].detach().numpy()), norm_err=pinfo["norm_err"].detach().numpy(), trans_err=pinfo["trans_err"].detach().numpy(),
            twist=pinfo["est_transform"].vec.detach().numpy(), consistency=cons.detach().numpy(), alpha=np.float32(args.alpha), lmbda=np.float32(args.lmbda), **gn,
        )


if __name__ == "__main__":
    main()
