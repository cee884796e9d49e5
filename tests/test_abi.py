"""The C-ABI library loads and exports exactly what include/va_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "va_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(va_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    fns = header_functions()
    for must in ("va_create", "va_load_weights", "va_preprocess", "va_forward", "va_fuse", "va_last_error"):
        assert must in fns


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in va_b200.h but not exported"


def test_python_binding_covers_header(built_lib):
    from video_analytics_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    handle = _lib.load()
    assert handle.va_abi_version() == 1
    assert handle.va_launch_count() == 0 or handle.va_launch_count() > 0   # callable without a GPU


def test_no_cpu_fallback(built_lib):
    """Product ops refuse CPU tensors instead of silently computing with torch."""
    import torch
    from video_analytics_b200 import ops
    from video_analytics_b200._lib import VAError
    x = torch.zeros(1, 8, 16, 64, dtype=torch.bfloat16)
    w = torch.zeros(64, 64, 3, 3)
    with pytest.raises(VAError):
        ops.conv2d_nhwc(x, w, torch.zeros(64))
    if not torch.cuda.is_available():
        with pytest.raises(VAError):
            ops.StreamNet(0, 3)   # va_create needs an sm_100 device


def test_sass_is_blackwell_native(built_lib):
    """tcgen05 / TMA / TMEM instructions are present in the shipped cubin (UTCHMMA, UTMALDG, UTMASTG, LDTM)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(built_lib)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass     # no legacy mma.sync path


def test_training_primitives_reject_bad_arguments_without_touching_the_gpu(built_lib):
    """Argument validation of the K5 entry points happens before any CUDA call (status 1 = VA_ERR_INVALID)."""
    from video_analytics_b200 import _lib
    lib = _lib.load()
    assert lib.va_wgrad(None, None, 1, 4, 4, 64, 64, 64, 3, None, None) != 0
    assert b"NULL" in lib.va_last_error()
    assert lib.va_conv2d_dgrad(None, 1, 4, 4, 64, None, 64, None, None) != 0
    assert lib.va_relu_pool_bwd(None, None, 1, 4, 4, 64, 1, None, None, None) != 0
    assert lib.va_sgd_momentum(None, None, None, 10, 0.1, 0.9, 1, 1.0, None) != 0
    assert lib.va_ce_train(None, None, None, None, 1, 256, 101, None, None, None, None, None, None, None) != 0
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.va_wgrad(p, p, 1, 4, 4, 64, 64, 64, 5, p, None) != 0           # ks must be 1 or 3
    assert b"ks" in lib.va_last_error()
    assert lib.va_conv2d_dgrad(p, 1, 4, 4, 48, p, 64, p, None) != 0             # channels must be multiples of 64
    assert lib.va_maxpool2x2_nhwc(p, 1, 5, 4, 64, p, None, None) != 0           # odd height
    # va_allreduce_bf16: neither peer pointers nor a multicast pointer
    assert lib.va_allreduce_bf16(None, None, 2, 0, ctypes.c_longlong(64), 8, None) != 0
    # va_svm_fit (combinedModel.py:34-35): NULL arguments, one class only, too many features, bad regularisation
    d = ctypes.c_double
    assert lib.va_svm_fit(None, None, 4, 8, 2, d(1.0), d(1.0), d(1e-4), 10, None, None, None, None, None) != 0
    assert b"NULL" in lib.va_last_error()
    assert lib.va_svm_fit(p, p, 4, 8, 1, d(1.0), d(1.0), d(1e-4), 10, p, p, p, p, None) != 0
    assert b"2 classes" in lib.va_last_error()
    assert lib.va_svm_fit(p, p, 4, 2000, 3, d(1.0), d(1.0), d(1e-4), 10, p, p, p, p, None) != 0
    assert b"F <= 1024" in lib.va_last_error()
    assert lib.va_svm_fit(p, p, 4, 8, 3, d(0.0), d(1.0), d(1e-4), 10, p, p, p, p, None) != 0


def test_training_has_no_cpu_fallback(built_lib):
    import torch
    from video_analytics_b200 import train_ops as T
    from video_analytics_b200._lib import VAError
    with pytest.raises(VAError):
        T.maxpool2x2(torch.zeros(1, 4, 4, 8, dtype=torch.bfloat16))
    with pytest.raises(VAError):
        T.conv2d_wgrad(torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16), torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16), 64)
    if not torch.cuda.is_available():
        from video_analytics_b200.training import StreamTrainer
        with pytest.raises(VAError):
            StreamTrainer(torch.nn.Linear(2, 2))
