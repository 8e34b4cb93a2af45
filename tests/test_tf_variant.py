"""TF-variant generator (classification/models/samplenet_model.py): name/layout translation of TF variables (CPU) and, on the GPU,
the kernels against a numpy restatement of the TF graph evaluated directly in the TF layout."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_tf_variables(seed=0, m=32, bottleneck=128, ema_style="nested"):
    """Variables named and shaped as the reference's TF graph creates them (tf_util.py:150-177, 361-379, 491-519)."""
    r = np.random.default_rng(seed)
    v = {}
    widths = [None, 64, 64, 64, 128, bottleneck]
    for i in range(1, 6):
        shape = [1, 3, 1, 64] if i == 1 else [1, 1, widths[i - 1], widths[i]]
        sc = "sampler/conv%d" % i
        v[sc + "/weights:0"] = (r.standard_normal(shape) * 0.3).astype(np.float32)
        v[sc + "/biases:0"] = (r.standard_normal(widths[i]) * 0.1).astype(np.float32)
        _bn(v, r, sc, widths[i], ema_style)
    fcw = [bottleneck, 256, 256, 256, 3 * m]
    for i in range(4):
        sc = "sampler/fc1%db" % (i + 1)
        v[sc + "/weights:0"] = (r.standard_normal([fcw[i], fcw[i + 1]]) * (1.0 / np.sqrt(fcw[i]))).astype(np.float32)
        v[sc + "/biases:0"] = (r.standard_normal(fcw[i + 1]) * 0.1).astype(np.float32)
        _bn(v, r, sc, fcw[i + 1], ema_style)
    v["sampler/temperature:0"] = np.float32(1.0)
    v["classifier/conv1/weights:0"] = np.zeros([1, 3, 1, 64], np.float32)   # another scope in the same checkpoint: must be ignored
    return v


def _bn(v, r, sc, c, ema_style):
    v[sc + "/bn/gamma:0"] = (1.0 + 0.2 * r.standard_normal(c)).astype(np.float32)
    v[sc + "/bn/beta:0"] = (0.1 * r.standard_normal(c)).astype(np.float32)
    mid = (sc + "/bn/") if ema_style == "nested" else ""
    v[sc + "/bn/" + mid + "moments/Squeeze/ExponentialMovingAverage:0"] = (0.1 * r.standard_normal(c)).astype(np.float32)
    v[sc + "/bn/" + mid + "moments/Squeeze_1/ExponentialMovingAverage:0"] = (0.5 + r.random(c)).astype(np.float32)


def tf_graph_numpy(v, x, training):
    """The TF graph of samplenet_model.get_model restated in numpy IN THE TF LAYOUT (float64): conv2d on the (B,N,3,1) image with the
    [1,3] kernel, 1x1 convs as channel contractions on (B,N,1,C), max over N, matmul with [Cin,Cout] weights, batch_norm_template."""
    def bn(y, sc, axes):
        g, be = v[sc + "/bn/gamma:0"].astype(np.float64), v[sc + "/bn/beta:0"].astype(np.float64)
        if training:
            mu, va = y.mean(axis=axes), y.var(axis=axes)
        else:
            mu = [a for k, a in v.items() if k.startswith(sc + "/bn/") and k.endswith("moments/Squeeze/ExponentialMovingAverage:0")][0].astype(np.float64)
            va = [a for k, a in v.items() if k.startswith(sc + "/bn/") and k.endswith("moments/Squeeze_1/ExponentialMovingAverage:0")][0].astype(np.float64)
        return (y - mu) / np.sqrt(va + 1e-3) * g + be
    img = x.astype(np.float64)[..., None]                                  # (B, N, 3, 1)
    k1 = v["sampler/conv1/weights:0"].astype(np.float64)                   # [1, 3, 1, 64]
    net = np.einsum("bnwi,wio->bno", img, k1[0])[:, :, None, :]            # VALID conv with a [1,3] kernel -> (B, N, 1, 64)
    net = np.maximum(bn(net + v["sampler/conv1/biases:0"], "sampler/conv1", (0, 1, 2)), 0)
    for i in range(2, 6):
        sc = "sampler/conv%d" % i
        net = np.einsum("bnwi,io->bnwo", net, v[sc + "/weights:0"].astype(np.float64)[0, 0]) + v[sc + "/biases:0"]
        net = np.maximum(bn(net, sc, (0, 1, 2)), 0)
    net = net.max(axis=1).reshape(x.shape[0], -1)
    for i in range(4):
        sc = "sampler/fc1%db" % (i + 1)
        net = bn(net @ v[sc + "/weights:0"].astype(np.float64) + v[sc + "/biases:0"], sc, (0,))
        if i < 3:
            net = np.maximum(net, 0)
    return net.reshape(x.shape[0], -1, 3)


def torch_specs_eval(conv, fc, x, training):
    """The translated layer tables evaluated with stock torch ops on the CPU (float64): what the kernels are asked to compute."""
    import torch.nn.functional as F
    y = torch.from_numpy(x).double().reshape(-1, 3)
    b = x.shape[0]
    for i, d in enumerate(conv + fc):
        if i == len(conv):
            y = y.view(b, -1, y.shape[1]).max(dim=1)[0]
        y = F.linear(y, torch.from_numpy(d["weight"]).double(), torch.from_numpy(d["bias"]).double())
        if "gamma" in d:
            y = F.batch_norm(y, None if training else torch.from_numpy(d["mean"]).double(), None if training else torch.from_numpy(d["var"]).double(),
                             torch.from_numpy(d["gamma"]).double(), torch.from_numpy(d["beta"]).double(), training, 0.0, 1e-3)
        if d["relu"]:
            y = F.relu(y)
    return y.view(b, -1, 3).numpy()


@pytest.mark.parametrize("ema_style", ["nested", "flat"])
@pytest.mark.parametrize("training", [True, False])
def test_tf_variable_translation_matches_tf_graph_semantics(ema_style, training):
    from samplenet_b200.tf_variant import layer_tables_from_tf
    v = make_tf_variables(1, m=32, ema_style=ema_style)
    conv, fc = layer_tables_from_tf(v, "sampler")
    assert [d["weight"].shape for d in conv] == [(64, 3), (64, 64), (64, 64), (128, 64), (128, 128)]
    assert [d["weight"].shape for d in fc] == [(256, 128), (256, 256), (256, 256), (96, 256)]
    assert all("gamma" in d for d in conv + fc) and [d["relu"] for d in fc] == [True, True, True, False]
    x = (np.random.default_rng(2).random((4, 200, 3)) - 0.5).astype(np.float32)
    np.testing.assert_allclose(torch_specs_eval(conv, fc, x, training), tf_graph_numpy(v, x, training), rtol=1e-9, atol=1e-10)


def test_tf_variable_translation_errors():
    from samplenet_b200.tf_variant import layer_tables_from_tf
    v = make_tf_variables(3)
    bad = dict(v); del bad["sampler/conv3/weights:0"]
    with pytest.raises(KeyError):
        layer_tables_from_tf(bad)
    bad = dict(v); bad["sampler/fc12b/weights:0"] = np.zeros((100, 256), np.float32)
    with pytest.raises(ValueError):
        layer_tables_from_tf(bad)
    bad = dict(v); bad["sampler/conv2/bn/extra/moments/Squeeze/ExponentialMovingAverage:0"] = np.zeros(64, np.float32)
    with pytest.raises(KeyError):
        layer_tables_from_tf(bad)
    conv, fc = layer_tables_from_tf({k.replace("sampler/", "", 1): a for k, a in v.items() if k.startswith("sampler/")}, scope="")
    assert len(conv) == 5 and len(fc) == 4


@pytest.mark.gpu
@pytest.mark.parametrize("training", [True, False])
def test_tf_variant_generator_on_gpu_matches_tf_graph(training):
    import __graft_entry__ as ge
    ge.build()
    from samplenet_b200.tf_variant import TFSampleNetGenerator
    v = make_tf_variables(4, m=32)
    gen = TFSampleNetGenerator.from_tf_variables(v, "sampler", bn_decay=0.5).cuda()
    gen.train(training)
    x = (np.random.default_rng(5).random((32, 1024, 3)) - 0.5).astype(np.float32)
    want = tf_graph_numpy(v, x, training)
    with torch.no_grad():
        got = gen(torch.from_numpy(x).cuda()).cpu().numpy()
    assert got.shape == (32, 32, 3)
    np.testing.assert_allclose(got, want, rtol=3e-4, atol=1e-4)   # fp32 kernels (3xTF32) vs the float64 restatement, O(1) outputs
    if training:   # moving averages advanced with momentum 1 - bn_decay (mean: exactly TF's rule)
        mu0 = v["sampler/conv2/bn/sampler/conv2/bn/moments/Squeeze/ExponentialMovingAverage:0"]
        assert not np.allclose(gen.l1_mean.cpu().numpy(), mu0)
