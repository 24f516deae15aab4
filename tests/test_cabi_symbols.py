"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/mst_b200.h declares (no compute)."""
import ctypes
import os
import re

from music_mixing_style_transfer_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mst_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    assert os.path.exists(built_lib)
    lib = ctypes.CDLL(built_lib)
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mst_b200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == names, "python binding table and header disagree"


def test_version_and_error_slot(built_lib):
    lib = _cabi.lib()
    assert lib.mst_version() == 100
    # argument validation happens on the host before any CUDA call
    cfg = _cabi.TcnConfig(n_blocks=14, n_inputs=2, n_outputs=2, channels=64, kernel_size=15, dilation_growth=2,
                          stack_size=15, cond_dim=2048)
    assert lib.mst_tcn_packed_bytes(ctypes.byref(cfg)) == 0
    assert "channel_width=128" in _cabi.last_error()
    cfg.channels = 128
    assert lib.mst_tcn_packed_bytes(ctypes.byref(cfg)) > 13 * 15 * 128 * 128 * 4
    assert lib.mst_fx_workspace_bytes(256, 262144) >= 256 * 16 * 8


def test_sass_uses_blackwell_tensor_and_tma_paths(built_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) and shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS: the TCN kernel is not on the tcgen05/TMA path"
    assert "HMMA.16816" not in sass  # no legacy mma.sync path
