"""ctypes loader for libsamplenet_b200.so (the C-ABI CUDA library, include/samplenet_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.  The product path never
touches oracle/ or any CPU implementation.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsamplenet_b200.so")

BNC, BCN = 0, 1
DIST_FMA, DIST_UNFUSED = 0, 1
GEN_EXACT_FP32 = 1
GEN_WORKSPACE_PRIMED = 32
EMD_EXACT = 1
SIGMA_VALUE, SIGMA_FROM_T_REG, SIGMA_FROM_T_CLS, SIGMA_FROM_T_REC = 0, 1, 2, 3

_c_float_p = ctypes.c_void_p  # raw device pointers travel as integers
_vp = ctypes.c_void_p
_int = ctypes.c_int
_size = ctypes.c_size_t
_float = ctypes.c_float


class Layer(ctypes.Structure):
    """Mirror of `snb200_layer`."""

    _fields_ = [
        ("c_in", _int), ("c_out", _int),
        ("weight", _vp), ("bias", _vp), ("bn_weight", _vp), ("bn_bias", _vp),
        ("bn_running_mean", _vp), ("bn_running_var", _vp), ("bn_num_batches_tracked", _vp),
        ("bn_eps", _float), ("bn_momentum", _float), ("relu", _int),
    ]


class LayerGrad(ctypes.Structure):
    """Mirror of `snb200_layer_grad`."""

    _fields_ = [("weight", _vp), ("bias", _vp), ("bn_weight", _vp), ("bn_bias", _vp)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "snb200_last_error": (ctypes.c_char_p, []),
    "snb200_version": (_int, []),
    "snb200_launch_count": (ctypes.c_ulonglong, []),
    "snb200_nn_distance_forward": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "snb200_nn_distance_backward": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "snb200_simplification_loss_workspace_bytes": (_size, [_int, _int, _int]),
    "snb200_simplification_loss_forward": (_int, [_int, _int, _vp, _int, _vp, _float, _vp, _vp, _vp, _vp, _vp, _vp, _size, _int, _vp]),
    "snb200_knn_soft_project_forward": (_int, [_int, _int, _int, _int, _int, _vp, _vp, _vp, _int, _float, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "snb200_soft_project_backward_workspace_bytes": (_size, [_int, _int, _int, _int, _int]),
    "snb200_soft_project_backward": (_int, [_int, _int, _int, _int, _int, _vp, _vp, _vp, _int, _float, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _size, _vp]),
    "snb200_project_and_loss_workspace_bytes": (_size, [_int, _int, _int]),
    "snb200_project_and_loss_forward": (_int, [_int, _int, _int, _int, _vp, _vp, _vp, _int, _float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _float, _vp, _vp, _size, _vp, _int, _vp]),
    "snb200_group_point": (_int, [_int, _int, _int, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "snb200_group_point_grad": (_int, [_int, _int, _int, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "snb200_encoder_workspace_bytes": (_size, [_int, _int, _int, ctypes.POINTER(Layer)]),
    "snb200_encoder_forward": (_int, [_int, _int, _int, _vp, _int, ctypes.POINTER(Layer), _int, _vp, _vp, _size, _vp]),
    "snb200_generator_workspace_bytes": (_size, [_int, _int, _int, ctypes.POINTER(Layer), _int, ctypes.POINTER(Layer)]),
    "snb200_generator_forward": (_int, [_int, _int, _int, _vp, _int, ctypes.POINTER(Layer), _int, ctypes.POINTER(Layer), _int, _vp, _int, _vp, _int, _vp, _size, _vp]),
    "snb200_generator_backward_supported": (_int, [_int, _int, _int, ctypes.POINTER(Layer), _int, ctypes.POINTER(Layer)]),
    "snb200_generator_train_forward": (_int, [_int, _int, _int, _vp, _int, ctypes.POINTER(Layer), _int, ctypes.POINTER(Layer), _vp, _int, _vp,
                                              ctypes.POINTER(_vp), _int, _vp, _size, _vp]),
    "snb200_generator_backward_workspace_bytes": (_size, [_int, _int, _int, ctypes.POINTER(Layer), _int, ctypes.POINTER(Layer)]),
    "snb200_generator_backward": (_int, [_int, _int, _int, _vp, _int, ctypes.POINTER(Layer), _int, ctypes.POINTER(Layer), ctypes.POINTER(_vp), _vp, _vp, _int,
                                         ctypes.POINTER(LayerGrad), ctypes.POINTER(LayerGrad), _vp, _size, _vp]),
    "snb200_debug_head_timestamps": (_int, [_vp]),
    "snb200_debug_conv_stack_timestamps": (_int, [_vp]),
    "snb200_debug_tc_gemm": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, ctypes.c_uint, _int, _int, _vp]),
    "snb200_fc_head_workspace_bytes": (_size, [_int, _int, ctypes.POINTER(Layer)]),
    "snb200_fc_head_forward": (_int, [_int, _vp, _int, ctypes.POINTER(Layer), _int, _vp, _int, _vp, _size, _vp]),
    "snb200_progressive_loss_workspace_bytes": (_size, [_int, _int, _int, _int]),
    "snb200_progressive_loss_forward": (_int, [_int, _int, _int, _vp, _vp, _int, ctypes.POINTER(_int), ctypes.POINTER(_float), _vp, _vp, _vp, _vp, _vp, _vp, _size,
                                               _vp, _int, _vp]),
    "snb200_approxmatch_workspace_bytes": (_size, [_int, _int, _int]),
    "snb200_approxmatch": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _size, _vp]),
    "snb200_approxmatch_mode": (_int, [_int, _int, _int, _vp, _vp, _vp, _int, _vp, _size, _vp]),
    "snb200_matchcost_workspace_bytes": (_size, [_int]),
    "snb200_matchcost": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _size, _vp]),
    "snb200_matchcostgrad": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "snb200_nn_matching": (_int, [_int, _int, _int, _int, _vp, _vp, _int, _vp, _vp, _vp]),
}

_lib = None


class SampleNetB200Error(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle.  Raises ImportError loudly if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "samplenet_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `python samplenet_b200/csrc/build.py`). There is no CPU fallback." % LIB_PATH
            )
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        msg = lib().snb200_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("%s: %s" % (what, msg))
        raise SampleNetB200Error("%s failed (rc=%d): %s" % (what, rc, msg))


def launch_count():
    return int(lib().snb200_launch_count())
