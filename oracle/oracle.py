"""ctypes/numpy binding of the CPU oracle (oracle/samplenet_oracle.c) and of the reference's own CPU code
(oracle/_ref/libsamplenet_ref.so, built by oracle/Makefile from the sources under /root/reference).

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; samplenet_b200/ never does.

All arrays are numpy, C-contiguous, float32 / int32, layouts as in the reference's TF ops (BNC).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)
_d = ctypes.POINTER(ctypes.c_double)


def build(quiet=True):
    """(Re)build liboracle.so and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=quiet)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libsamplenet_ref.so"))


def _ref():
    global _REF
    if _REF is None:
        _REF = ctypes.CDLL(os.path.join(_HERE, "_ref", "libsamplenet_ref.so"))
    return _REF


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


# ----------------------------------------------------------------------------- Chamfer / nn_distance
def nn_distance(xyz1, xyz2, contract=False):
    """(dist1, idx1, dist2, idx2) like tf_nndistance.py:12-23.  contract=False == reference CPU arithmetic,
    contract=True == the reference CUDA kernels' FMA-contracted arithmetic."""
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.empty((b, n), np.float32); i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32); i2 = np.empty((b, m), np.int32)
    _lib().orc_nn_distance(b, n, m, _fp(xyz1), _fp(xyz2), _fp(d1), _ip(i1), _fp(d2), _ip(i2), int(contract))
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, g1, idx1, g2, idx2):
    xyz1, xyz2, g1, g2 = _c(xyz1), _c(xyz2), _c(g1), _c(g2)
    idx1, idx2 = _c(idx1, np.int32), _c(idx2, np.int32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = np.empty_like(xyz1); gx2 = np.empty_like(xyz2)
    _lib().orc_nn_distance_grad(b, n, m, _fp(xyz1), _fp(xyz2), _fp(g1), _ip(idx1), _fp(g2), _ip(idx2), _fp(gx1), _fp(gx2))
    return gx1, gx2


def ref_chamfer_forward(xyz1, xyz2):
    """The reference's own chamfer_distance_forward (chamfer_distance.cpp:90-111), compiled unmodified."""
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((b, n), np.float32); i1 = np.zeros((b, n), np.int32)
    d2 = np.zeros((b, m), np.float32); i2 = np.zeros((b, m), np.int32)
    _ref().ref_chamfer_forward(b, n, m, _fp(xyz1), _fp(xyz2), _fp(d1), _fp(d2), _ip(i1), _ip(i2))
    return d1, i1, d2, i2


def ref_chamfer_backward(xyz1, xyz2, g1, idx1, g2, idx2):
    xyz1, xyz2, g1, g2 = _c(xyz1), _c(xyz2), _c(g1), _c(g2)
    idx1, idx2 = _c(idx1, np.int32), _c(idx2, np.int32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = np.zeros_like(xyz1); gx2 = np.zeros_like(xyz2)
    _ref().ref_chamfer_backward(b, n, m, _fp(xyz1), _fp(xyz2), _fp(gx1), _fp(gx2), _fp(g1), _fp(g2), _ip(idx1), _ip(idx2))
    return gx1, gx2


# ----------------------------------------------------------------------------- kNN / grouping / projection
def knn_point(k, xyz1, xyz2, contract=False, tie_mode=0):
    """(val (b,m,k), idx (b,m,k)) like tf_grouping.py:64-91; xyz1 dataset, xyz2 queries."""
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    val = np.empty((b, m, k), np.float32); idx = np.empty((b, m, k), np.int32)
    rc = _lib().orc_knn_point(b, n, m, k, _fp(xyz1), _fp(xyz2), _fp(val), _ip(idx), int(contract), int(tie_mode))
    if rc != 0:
        raise ValueError("orc_knn_point failed rc=%d (k > n?)" % rc)
    return val, idx


def group_point(points, idx):
    points, idx = _c(points), _c(idx, np.int32)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    _lib().orc_group_point(b, n, c, m, ns, _fp(points), _ip(idx), _fp(out))
    return out


def group_point_grad(points_shape, idx, grad_out):
    idx, grad_out = _c(idx, np.int32), _c(grad_out)
    b, n, c = points_shape
    _, m, ns = idx.shape
    gp = np.empty((b, n, c), np.float32)
    _lib().orc_group_point_grad(b, n, c, m, ns, _fp(grad_out), _ip(idx), _fp(gp))
    return gp


def soft_project(points, query, idx, sigma, hard=False, feats=None):
    """points (b,n,3), query (b,m,3), idx (b,m,k) -> proj (b,m,3), weights (b,m,k), dist (b,m,k)[, prop (b,m,f)]."""
    points, query, idx = _c(points), _c(query), _c(idx, np.int32)
    b, n, _ = points.shape
    _, m, k = idx.shape
    proj = np.empty((b, m, 3), np.float32); w = np.empty((b, m, k), np.float32); d = np.empty((b, m, k), np.float32)
    if feats is not None:
        feats = _c(feats)
        f = feats.shape[2]
        prop = np.empty((b, m, f), np.float32)
        _lib().orc_soft_project(b, n, m, k, _fp(points), _fp(query), _ip(idx), ctypes.c_float(sigma), int(hard),
                                _fp(feats), f, _fp(proj), _fp(w), _fp(d), _fp(prop))
        return proj, w, d, prop
    _lib().orc_soft_project(b, n, m, k, _fp(points), _fp(query), _ip(idx), ctypes.c_float(sigma), int(hard),
                            None, 0, _fp(proj), _fp(w), _fp(d), None)
    return proj, w, d


# ----------------------------------------------------------------------------- EMD
def approx_match(xyz1, xyz2):
    """match (b, m, n) with GPU-kernel semantics (tf_approxmatch_g.cu:1-179)."""
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.empty((b, m, n), np.float32)
    rc = _lib().orc_approxmatch(b, n, m, _fp(xyz1), _fp(xyz2), _fp(match))
    if rc != 0:
        raise MemoryError("orc_approxmatch rc=%d" % rc)
    return match


def match_cost(xyz1, xyz2, match):
    xyz1, xyz2, match = _c(xyz1), _c(xyz2), _c(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost = np.empty((b,), np.float32)
    _lib().orc_matchcost(b, n, m, _fp(xyz1), _fp(xyz2), _fp(match), _fp(cost))
    return cost


def match_cost_grad(xyz1, xyz2, match):
    xyz1, xyz2, match = _c(xyz1), _c(xyz2), _c(match)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.empty((b, n, 3), np.float32); g2 = np.empty((b, m, 3), np.float32)
    _lib().orc_matchcostgrad(b, n, m, _fp(xyz1), _fp(xyz2), _fp(match), _fp(g1), _fp(g2))
    return g1, g2


def ref_approxmatch_cpu(xyz1, xyz2):
    """Reference approxmatch_cpu (approxmatch.cpp:17-76): match laid out (b, n, m), double accumulators."""
    xyz1, xyz2 = _c(xyz1), _c(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.zeros((b, n, m), np.float32)
    _ref().ref_approxmatch_cpu(b, n, m, _fp(xyz1), _fp(xyz2), _fp(match))
    return match


def ref_matchcost_cpu(xyz1, xyz2, match_nm):
    xyz1, xyz2, match_nm = _c(xyz1), _c(xyz2), _c(match_nm)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost = np.zeros((b,), np.float32)
    _ref().ref_matchcost_cpu(b, n, m, _fp(xyz1), _fp(xyz2), _fp(match_nm), _fp(cost))
    return cost


def ref_matchcostgrad_cpu(xyz1, xyz2, match_nm):
    xyz1, xyz2, match_nm = _c(xyz1), _c(xyz2), _c(match_nm)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g2 = np.zeros((b, m, 3), np.float32)
    _ref().ref_matchcostgrad_cpu(b, n, m, _fp(xyz1), _fp(xyz2), _fp(match_nm), _fp(g2))
    return g2


# ----------------------------------------------------------------------------- inference matching
def nn_matching(full_pc, idx, k, complete_fps=True):
    """sputils.nn_matching (registration/src/sputils.py:31-41): full_pc (B,N,3) float32, idx (B,M) -> (B,k,3) float64."""
    full_pc = _c(full_pc)
    idx = _c(idx, np.int32)
    bsz, n, _ = full_pc.shape
    out = np.zeros((bsz, k, 3), np.float64)
    for ii in range(bsz):
        _lib().orc_nn_matching_one(n, k, idx.shape[1], _fp(full_pc[ii]), _ip(idx[ii]), int(complete_fps),
                                   out[ii].ctypes.data_as(_d))
    return out


# ----------------------------------------------------------------------------- loss assembly (numpy)
def simplification_loss(ref_pc, samp_pc, pc_size, gamma=1.0, delta=0.0, contract=False):
    """registration/src/samplenet.py:171-181 / classification/models/samplenet_model.py:176-188, in float32 numpy."""
    c12, _, c21, _ = nn_distance(samp_pc, ref_pc, contract)
    max_cost = np.float32(np.mean(np.max(c12, axis=1), dtype=np.float32))
    m12 = np.float32(np.mean(c12, dtype=np.float32))
    m21 = np.float32(np.mean(c21, dtype=np.float32))
    return np.float32(m12 + max_cost + np.float32(gamma + delta * pc_size) * m21)
