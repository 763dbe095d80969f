"""CPU: the C-ABI library builds, loads and exports every symbol include/genconvit_b200.h declares.
No compute calls here (no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "genconvit_b200.h")).read()
    return sorted(set(re.findall(r"\b(gcv_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from genconvit_b200 import build, lib
    path = build.build()
    assert os.path.exists(path)
    dll = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(dll, s), f"{s} declared in the header but not exported"
    assert sorted(lib.EXPORTS) == syms
    dll.gcv_abi_version.restype = ctypes.c_int
    assert dll.gcv_abi_version() == 3


def test_epilogue_struct_matches_header_layout():
    from genconvit_b200 import lib
    # 8-byte pointers / int64 with natural alignment, as the C compiler lays out gcv_epilogue
    assert ctypes.sizeof(lib.Epilogue) == 120
    assert lib.Epilogue.ldd.offset == 80 and lib.Epilogue.mu_out.offset == 56 and lib.Epilogue.ln_stats.offset == 96
    assert lib.Epilogue.ln_eps.offset == 116


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import shutil
    import subprocess
    from genconvit_b200 import build
    if shutil.which("cuobjdump") is None:
        return
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing: the tcgen05/TMA path did not compile in"


def test_product_path_fails_loudly_without_gpu():
    import pytest
    import torch
    from genconvit_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.GcvError):
        lib.require_cuda(torch.zeros(1), "test")
