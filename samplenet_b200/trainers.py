"""Loss composition of the three reference trainers, restated over this package's API (the callers).

Only the sampling-related step logic is here; the task networks (PCRNet, the PointNet classifier, the auto-encoder), their
optimisers and the data pipelines stay the reference's.  Every function takes the task-side quantities as arguments and returns
what the reference's step returns, so a maintainer can swap the body of the corresponding method for one call.

    registration    registration/main.py:500-538         compute_samplenet_loss
    classification  classification/train_samplenet.py:163-180
    reconstruction  reconstruction/src/pointnet_ae.py:110-124 (AE loss), samplenet_pointnet_ae.py:165-189 (simplification loss)
"""
import torch

from . import tf_ops


def registration_samplenet_loss(sampler, p0, p1, num_out_points, alpha, lmbda, gamma=1, delta=0, num_sampled_clouds=1):
    """`Action.compute_samplenet_loss` (registration/main.py:500-538).  p0: template, p1: source, both in `sampler.input_shape`.
    Returns (samplenet_loss, (p0_out, p1_projected), info) with info = {"simplification_loss", "projection_loss"}."""
    if num_sampled_clouds not in (1, 2):
        raise ValueError("num_sampled_clouds must be 1 or 2")
    to_bnc = (lambda t: t) if sampler.input_shape == "bnc" else (lambda t: t.permute(0, 2, 1).contiguous())
    out_bnc = (lambda t: t) if sampler.output_shape == "bnc" else (lambda t: t.permute(0, 2, 1).contiguous())
    p1_simplified, p1_projected = sampler(p1)
    simplification_loss = sampler.get_simplification_loss(to_bnc(p1), out_bnc(p1_simplified), num_out_points, gamma, delta)
    p0_out = p0
    if num_sampled_clouds == 2:   # sample the template as well
        p0_simplified, p0_projected = sampler(p0)
        p0_loss = sampler.get_simplification_loss(to_bnc(p0), out_bnc(p0_simplified), num_out_points, gamma, delta)
        simplification_loss = 0.5 * (simplification_loss + p0_loss)
        p0_out = p0_projected
    projection_loss = sampler.get_projection_loss()
    samplenet_loss = alpha * simplification_loss + lmbda * projection_loss
    return samplenet_loss, (p0_out, p1_projected), {"simplification_loss": simplification_loss, "projection_loss": projection_loss}


def classification_total_loss(loss_classifier, point_clouds, simplified_points, num_out_points, alpha, lmbda, gamma, delta, loss_projection):
    """classification/train_samplenet.py:173-180: `loss_classifier + ALPHA * loss_simplification + LMBDA * loss_projection` with the
    TF model's `get_simplification_loss` (samplenet_model.py:176-188) on (B,N,3) clouds.  Returns (loss, loss_simplification)."""
    loss_simplification = tf_ops.get_simplification_loss(point_clouds, simplified_points, num_out_points, gamma, delta)
    return loss_classifier + alpha * loss_simplification + lmbda * loss_projection, loss_simplification


def autoencoder_loss(x_reconstr, gt, loss="chamfer"):
    """`PointNetAutoEncoder._create_loss` (reconstruction/src/pointnet_ae.py:113-124): Chamfer = mean(d12) + mean(d21) over the batch;
    EMD = mean over the batch of match_cost(x, gt, approx_match(x, gt))."""
    if loss == "chamfer":
        cost_p1_p2, _, cost_p2_p1, _ = tf_ops.nn_distance(x_reconstr, gt)
        return cost_p1_p2.mean() + cost_p2_p1.mean()
    if loss == "emd":
        match = tf_ops.approx_match(x_reconstr, gt)
        return tf_ops.match_cost(x_reconstr, gt, match).mean()
    raise ValueError("loss must be 'chamfer' or 'emd'")


def autoencoder_simplification_loss(ref_pc, samp_pc, pc_size, is_denoising=False):
    """`SampleNetPointNetAE._get_simplification_loss` (reconstruction/src/samplenet_pointnet_ae.py:165-189): weight w = pc_size / 64
    (doubled when denoising) on the input->sample term.  Returns (loss, dist, idx, dist2, nn_distance_per_cloud (B,1))."""
    cost_p1_p2, idx, cost_p2_p1, _ = tf_ops.nn_distance(samp_pc, ref_pc)
    max_cost = cost_p1_p2.max(dim=1)[0].mean()
    per_cloud = cost_p1_p2.mean(dim=1, keepdim=True) + cost_p2_p1.mean(dim=1, keepdim=True)
    w = pc_size / 64.0
    loss = cost_p1_p2.mean() + max_cost + (2 * w if is_denoising else w) * cost_p2_p1.mean()
    return loss, cost_p1_p2, idx, cost_p2_p1, per_cloud


def progressive_simplification_loss(ref_pc, ordered_samples, sizes, gamma=1, delta=0, one_pass=True):
    """SampleNetProgressive (classification/train_samplenet_progressive.py:196-220): the simplification loss summed over the
    prefixes `ordered_samples[:, :s]` for s in `sizes`.  one_pass=True (default): ONE launch evaluates every prefix (the sample -> input
    distances of a prefix are a slice, the input -> sample distances a running prefix minimum: csrc/progressive.cu); one_pass=False: the
    reference's structure, one Chamfer evaluation per prefix."""
    sizes = [int(s) for s in sizes]
    if one_pass and len(sizes) <= 16 and ordered_samples.shape[1] <= 4096 and sorted(set(sizes)) == sizes:
        from . import ops

        total, _ = ops.ProgressiveLossFunction.apply(ordered_samples, ref_pc, sizes, [gamma + delta * s for s in sizes])
        return total
    total = torch.zeros((), device=ref_pc.device)
    for s in sizes:
        total = total + tf_ops.get_simplification_loss(ref_pc, ordered_samples[:, :s].contiguous(), s, gamma, delta)
    return total


# ----------------------------------------------------------------------------------------------------- whole steps with task networks
class ClassificationStep:
    """One training step of classification/train_samplenet.py:154-199: sampler (generator + TF-flavoured soft projection, sigma = T^2) in
    front of a FROZEN PointNet classifier; loss = loss_classifier + ALPHA * loss_simplification + LMBDA * loss_projection.
    `sampler` is a SampleNet-like module returning (simplified, projected) on (B,N,3) input; `classifier` a tasknets.PointNetCls."""

    def __init__(self, sampler, classifier, num_out_points, alpha=30.0, lmbda=1.0, gamma=1.0, delta=0.0):
        self.sampler, self.classifier = sampler, classifier
        self.M, self.alpha, self.lmbda, self.gamma, self.delta = num_out_points, alpha, lmbda, gamma, delta
        classifier.requires_grad_(False)
        classifier.eval()

    def loss(self, point_clouds, labels):
        simplified, projected = self.sampler(point_clouds)
        pred, end_points = self.classifier(projected)
        loss_classifier = self.classifier.get_loss(pred, labels, end_points)
        loss_simplification = self.sampler.get_simplification_loss(point_clouds, simplified, self.M, self.gamma, self.delta)
        loss_projection = self.sampler.get_projection_loss()
        total = loss_classifier + self.alpha * loss_simplification + self.lmbda * loss_projection
        return total, {"loss_classifier": loss_classifier, "loss_simplification": loss_simplification, "loss_projection": loss_projection, "pred": pred}


class ProgressiveClassificationStep(ClassificationStep):
    """classification/train_samplenet_progressive.py:156-230: ONE generator pass emits MAX ordered points; the classifier and the
    simplification loss are evaluated on every power-of-two prefix and summed."""

    def __init__(self, sampler, classifier, min_points, max_points, alpha=30.0, lmbda=1.0, gamma=1.0, delta=0.0):
        super().__init__(sampler, classifier, max_points, alpha, lmbda, gamma, delta)
        self.sizes = []
        s = min_points
        while s <= max_points:
            self.sizes.append(s)
            s *= 2

    def loss(self, point_clouds, labels):
        simplified, projected = self.sampler(point_clouds)
        loss_classifier = 0.0
        for s in self.sizes:
            pred, end_points = self.classifier(projected[:, :s].contiguous())
            loss_classifier = loss_classifier + self.classifier.get_loss(pred, labels, end_points)
        loss_simplification = progressive_simplification_loss(point_clouds, simplified, self.sizes, self.gamma, self.delta)
        loss_projection = self.sampler.get_projection_loss()
        total = loss_classifier + self.alpha * loss_simplification + self.lmbda * loss_projection
        return total, {"loss_classifier": loss_classifier, "loss_simplification": loss_simplification, "loss_projection": loss_projection}


class ReconstructionStep:
    """One training step of reconstruction/src/samplenet_pointnet_ae.py:46-189: sampler (rec widths, sigma = max(T, 1e-2)^2) in front of a
    FROZEN auto-encoder; loss = AE loss(reconstruction of the projected points, input cloud) [Chamfer or EMD] + ALPHA * simplification
    loss (weight pc_size / 64 on the input -> sample term) + LMBDA * projection loss."""

    def __init__(self, sampler, ae, num_out_points, alpha=0.01, lmbda=1e-4, ae_loss="chamfer"):
        self.sampler, self.ae, self.M, self.alpha, self.lmbda, self.ae_loss = sampler, ae, num_out_points, alpha, lmbda, ae_loss
        ae.requires_grad_(False)
        ae.eval()

    def loss(self, point_clouds):
        simplified, projected = self.sampler(point_clouds)
        x_reconstr = self.ae(projected)
        loss_ae = autoencoder_loss(x_reconstr, point_clouds, self.ae_loss)
        loss_simplification, _, _, _, _ = autoencoder_simplification_loss(point_clouds, simplified, self.M)
        loss_projection = self.sampler.get_projection_loss()
        total = loss_ae + self.alpha * loss_simplification + self.lmbda * loss_projection
        return total, {"loss_ae": loss_ae, "loss_simplification": loss_simplification, "loss_projection": loss_projection}
