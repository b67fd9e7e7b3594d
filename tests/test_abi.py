"""CPU: the C-ABI library builds, loads without a GPU driver, and exports every symbol include/maed_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "maed_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(maed_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_python_binding_covers_header(lib):
    from maed_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_engine_tables_without_gpu(lib):
    """Engine construction and its parameter table are host-only logic."""
    from maed_b200 import _lib
    cfg = _lib.MaedConfig(6, 12, _lib.MODES["parallel"], _lib.DECODERS["ktd"], 1024, 3, 16)
    h = ctypes.c_void_p()
    assert lib.maed_engine_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    n = lib.maed_engine_num_params(h)
    names = [lib.maed_engine_param_name(h, i).decode() for i in range(n)]
    assert n == 305 and len(set(names)) == n          # 305 parameter tensors (parallel + ktd)
    total = sum(lib.maed_engine_param_numel(h, i) for i in range(n))
    assert total == 72132153                           # 72 132 153 parameters
    assert lib.maed_engine_packed_bytes(h) > 0
    assert lib.maed_engine_workspace_bytes(h, 128) > lib.maed_engine_workspace_bytes(h, 8) > 0
    lib.maed_engine_destroy(h)
    bad = _lib.MaedConfig(6, 8, 1, 0, 1024, 3, 16)     # head_dim 96 unsupported -> error, not a crash
    assert lib.maed_engine_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"num_heads" in lib.maed_last_error()


def test_cnn_engine_tables_without_gpu(lib):
    """encoder='cnn': torchvision ResNet-50 parameter table (53 conv + BN pairs) + KTD decoder on 2048 features."""
    from maed_b200 import _lib
    cfg = _lib.MaedConfig(6, 12, _lib.MODES["vanilla"], _lib.DECODERS["ktd"], 1024, 3, 16, _lib.ENCODERS["cnn"])
    h = ctypes.c_void_p()
    assert lib.maed_engine_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    n = lib.maed_engine_num_params(h)
    names = [lib.maed_engine_param_name(h, i).decode() for i in range(n)]
    assert len(set(names)) == n and n == 53 * 5 + 4 + 48 + 4
    assert "encoder.layer4.2.bn3.running_var" in names and "encoder.layer2.0.downsample.0.weight" in names
    numel = dict(zip(names, (lib.maed_engine_param_numel(h, i) for i in range(n))))
    assert numel["decoder.fc1.weight"] == 1024 * 2048 and numel["encoder.conv1.weight"] == 64 * 3 * 49
    assert lib.maed_engine_workspace_bytes(h, 16) > lib.maed_engine_workspace_bytes(h, 2) > 0
    assert lib.maed_train_workspace_bytes(h, 4) > lib.maed_engine_workspace_bytes(h, 4)   # the tape of the training path
    lib.maed_engine_destroy(h)
    bad = _lib.MaedConfig(6, 12, 0, 0, 1024, 3, 16, 7)
    assert lib.maed_engine_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"encoder" in lib.maed_last_error()
