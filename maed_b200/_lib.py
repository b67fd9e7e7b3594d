"""ctypes binding of libmaed_b200.so (C ABI declared in include/maed_b200.h).

The library is built in-tree by ``maed_b200/build.py`` (nvcc, sm_100a).  There is NO fallback: if the
shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmaed_b200.so")

c_void_pp = C.POINTER(C.c_void_p)
c_float_p = C.c_void_p          # device pointers are passed as integers
_P = C.c_void_p
_I = C.c_int
_L = C.c_longlong
_F = C.c_float
_D = C.c_double
_Z = C.c_size_t


class MaedConfig(C.Structure):
    _fields_ = [("num_blocks", _I), ("num_heads", _I), ("mode", _I), ("decoder", _I), ("hidden_dim", _I),
                ("nsplit", _I), ("temp_frames", _I), ("encoder", _I)]


class MaedOutputs(C.Structure):
    _fields_ = [("feat", _P), ("pose6d", _P), ("shape", _P), ("cam", _P), ("rotmat", _P), ("theta", _P),
                ("kp2d", _P), ("kp3d", _P), ("n_joints", _I)]


class MaedTrainOutputs(C.Structure):
    _fields_ = [("feat", _P), ("pose6d", _P), ("shape", _P), ("cam", _P)]


class MaedSmplAssets(C.Structure):
    _fields_ = [("v_template", _P), ("shapedirs", _P), ("posedirs", _P), ("J_template", _P), ("J_shapedirs", _P),
                ("lbs_weights", _P), ("J_regressor_extra", _P), ("parents", _P), ("extra_vertex_ids", _P), ("joint_map", _P)]


class MaedLossWeights(C.Structure):
    _fields_ = [("kp2d", C.c_float), ("kp3d", C.c_float), ("pose", C.c_float), ("shape", C.c_float), ("norm", C.c_float),
                ("accl", C.c_float)]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int)     # maed_exchange_fn
PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int)     # maed_progress_fn
_U = C.c_ulonglong
MODES = {"vanilla": 0, "parallel": 1, "series": 2, "coupling": 3, "temporal": 4}
DECODERS = {"ktd": 0, "iterative": 1}
ENCODERS = {"ste": 0, "cnn": 1}
TAP_NAMES = ["stem", "stage0", "stage1", "stage2", "embed"] + ["block%d" % i for i in range(8)]

# name -> (restype, argtypes); every symbol declared in include/maed_b200.h
SIGNATURES = {
    "maed_last_error": (C.c_char_p, []),
    "maed_version": (_I, []),
    "maed_launch_count": (_L, []),
    "maed_op_gemm": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _L, _I, _I, _P]),
    "maed_op_gemm_bottleneck": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _P, _P, _L, _I, _I, _P, _L, _I, _P]),
    "maed_op_conv_gemm": (_I, [_P, _L, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _I, _P]),
    "maed_op_fold_bn": (_I, [_P, _I, _L, _P, _P, _P, _P, _F, _P, _P, _P]),
    "maed_op_maxpool3x3s2": (_I, [_P, _I, _I, _I, _I, _P, _P, _L, _P]),
    "maed_op_split_f32": (_I, [_P, _P, _L, _L, _P]),
    "maed_op_prep_conv_weight": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    "maed_op_im2col_stem": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    "maed_op_im2col_nhwc": (_I, [_P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    "maed_op_conv_gn": (_I, [_P, _L, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _F, _I, _P, _L, _P, _L, _P, _P]),
    "maed_op_stem_conv": (_I, [_P, _I, _P, _L, _I, _I, _P, _P, _P]),
    "maed_op_groupnorm": (_I, [_P, _I, _I, _I, _P, _P, _F, _I, _P, _L, _P, _L, _P, _P]),
    "maed_op_groupnorm_train": (_I, [_P, _I, _I, _I, _P, _P, _F, _I, _P, _L, _P, _L, _P, _P]),
    "maed_op_groupnorm_maxpool": (_I, [_P, _I, _I, _I, _I, _P, _P, _F, _P, _L, _P, _P]),
    "maed_op_layernorm": (_I, [_P, _L, _P, _P, _I, _I, _F, _P, _L, _P]),
    "maed_op_attention": (_I, [_I, _P, _L, _I, _I, _I, _I, _F, _I, _P, _P, _L, _P]),
    "maed_op_linear_f32": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _P]),
    "maed_op_decode_outputs": (_I, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "maed_engine_create": (_I, [C.POINTER(MaedConfig), C.POINTER(_P)]),
    "maed_engine_destroy": (None, [_P]),
    "maed_engine_num_params": (_I, [_P]),
    "maed_engine_param_name": (C.c_char_p, [_P, _I]),
    "maed_engine_param_numel": (_L, [_P, _I]),
    "maed_engine_packed_bytes": (_Z, [_P]),
    "maed_engine_workspace_bytes": (_Z, [_P, _I]),
    "maed_engine_pack": (_I, [_P, c_void_pp, _P, _P]),
    "maed_engine_forward": (_I, [_P, c_void_pp, _P, _P, _I, _I, _P, _Z, C.POINTER(MaedOutputs), c_void_pp, _P]),
    # ---- training path
    "maed_train_set_exchange": (_I, [_P, EXCHANGE_FN, _P, _P, _I]),
    "maed_train_set_progress": (_I, [_P, PROGRESS_FN, _P]),
    "maed_train_pack_bytes": (_Z, [_P]),
    "maed_train_workspace_bytes": (_Z, [_P, _I]),
    "maed_train_pack": (_I, [_P, c_void_pp, _P, _P]),
    "maed_train_forward": (_I, [_P, c_void_pp, _P, _P, _I, _I, _P, _Z, _F, _U, C.POINTER(MaedTrainOutputs), _P]),
    "maed_train_backward": (_I, [_P, c_void_pp, _P, _P, _P, _I, _I, _P, _Z, _P, _P, _P, _F, _F, c_void_pp, _P]),
    "maed_adam_step": (_I, [_P, _P, _P, _P, _L, _D, _D, _D, _D, _D, _I, _F, _P]),
    "maed_bwd_transpose_planes": (_I, [_P, _L, _I, _I, _I, _P, _L, _I, _P]),
    "maed_bwd_colsum_chunks": (_I, []),
    "maed_bwd_colsum": (_I, [_P, _L, _I, _I, _F, _I, _P, _P, _P]),
    "maed_bwd_layernorm": (_I, [_P, _L, _P, _L, _P, _I, _I, _F, _P, _P, _L, _P, _P, _P, _P, _P]),
    "maed_bwd_layernorm_partial_rows": (_I, []),
    "maed_bwd_groupnorm": (_I, [_P, _P, _I, _I, _I, _P, _F, _P, _P, _P, _P, _L, _P, _I, _P]),
    "maed_bwd_batchnorm_scratch_doubles": (_Z, [_L, _I]),
    "maed_bwd_batchnorm": (_I, [_P, _L, _I, _P, _P, _F, _F, _P, _P, _I, _P, _L, _P, _L, _P, _P, _P, _F, _P, _P, _P, _L, _P, _P]),
    "maed_bwd_maxpool3x3s2": (_I, [_P, _I, _I, _I, _I, _P, _L, _P, _P, _P, _P]),
    "maed_bwd_wstd": (_I, [_P, _I, _P, _I, _I, _I, _I, _F, _F, _P, _P]),
    "maed_bwd_gelu": (_I, [_P, _P, _L, _P, _L, _P]),
    "maed_bwd_relu_mask": (_I, [_P, _P, _L, _P]),
    "maed_bwd_maxpool": (_I, [_P, _I, _I, _I, _I, _P, _P, _F, _P, _P, _L, _P, _P, _P, _P]),
    "maed_bwd_dilate2": (_I, [_P, _L, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    "maed_bwd_scatter_stride2": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "maed_bwd_blend": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "maed_bwd_sgemm": (_I, [_I, _I, _I, _I, _I, _F, _P, _I, _P, _I, _F, _P, _I, _P]),
    "maed_bwd_ktd_tree": (_I, [_P, _P, _P, _P, _P, _I, _F, _P, _P, _I, _P, _P]),
    "maed_bwd_attention": (_I, [_I, _P, _L, _P, _I, _I, _I, _I, _F, _I, _P, _P, _P]),
    "maed_bwd_wgrad_slab_floats": (_Z, [_I, _I, _I]),
    "maed_bwd_wgrad_splitk": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _F, _I, _P, _P, _I, _P]),
    "maed_bwd_wgrad_rows": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P, _I, _P]),
    "maed_bwd_wgrad_conv": (_I, [_P, _L, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P, _I, _P]),
    "maed_bwd_split_transposed": (_I, [_P, _I, _I, _P, _L, _P]),
    "maed_bwd_prep_conv_weight_dgrad": (_I, [_P, _I, _I, _I, _I, _I, _P, _L, _P]),
    "maed_bwd_dropout": (_I, [_P, _L, _F, _U, _P, _P, _P]),
    # ---- geometry tail (training)
    "maed_decode_pose_backward": (_I, [_P, _I, _P, _P, _I, _P, _P]),
    "maed_project_keypoints": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    # ---- fused loss
    "maed_loss_scratch_bytes": (_Z, [_I, _I]),
    "maed_loss_forward_backward": (_I, [_P, _P, _I, _I, _P, _P, _I, _I, _P, _P, _P, _I, C.POINTER(MaedLossWeights), _P, _P, _P, _P,
                                        _P, _Z, _P]),
    # ---- SMPL
    "maed_smpl_scratch_bytes": (_Z, [_I]),
    "maed_smpl_forward": (_I, [C.POINTER(MaedSmplAssets), _P, _P, _I, _P, _I, _P, _P, _P, _Z, _P]),
    "maed_smpl_backward_scratch_bytes": (_Z, [_I]),
    "maed_smpl_backward": (_I, [C.POINTER(MaedSmplAssets), _P, _P, _I, _P, _I, _P, _P, _P, _P, _P, _Z, _P]),
}

_lib = None


def load():
    """Loads the shared library (once) and installs the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "maed_b200: %s not found — build it with `python -m maed_b200.build` (or "
            "__graft_entry__.build()).  There is no CPU / PyTorch fallback for the hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().maed_last_error()
        raise RuntimeError("maed_b200 %s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))


def call(name, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
