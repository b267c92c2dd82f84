"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol the header declares,
and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from multi_view_active_learning_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mval_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mval_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), "libmval_b200.so does not export %s" % name
        assert name in _lib.PROTOTYPES, "%s has no ctypes prototype" % name
    assert sorted(_lib.PROTOTYPES) == names
    assert lib.mval_version() == _lib.ABI_VERSION
    assert lib.mval_last_error() is not None


def test_library_is_plain_c_abi_without_torch():
    # the boundary must not depend on torch / python: check the dynamic dependencies of the .so
    import subprocess

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    rc = lib.mval_decode_argmax(None, 1, 2, 1, 64, 64, 4, None, None, None, None)
    assert rc == _lib.MVAL_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.mval_last_error()
    from multi_view_active_learning_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.decode_argmax(torch.zeros(1, 2, 1, 64, 64), 4)
    with pytest.raises(_lib.MvalError):
        _lib.check(rc)


def test_ransac_params_layout_matches_header():
    p = _lib.RansacParams()
    assert ctypes.sizeof(p) == 48
    assert _lib.RansacParams.epsilon.offset == 8 and _lib.RansacParams.pairs.offset == 32
    assert _lib.RansacParams.frame_keys.offset == 40


def test_header_is_plain_c(tmp_path):
    """include/mval_b200.h must be consumable from C (the boundary is a C ABI, not C++): compile a translation unit that
    includes it and takes the address of every declared entry point with gcc -std=c99 -pedantic."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    names = declared_symbols()
    src = tmp_path / "abi_check.c"
    body = "\n".join("  p[%d] = (void (*)(void))%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include "mval_b200.h"\nvoid (*p[%d])(void);\nint main(void) {\n%s\n  return MVAL_ABI_VERSION == %d ? 0 : 1;\n}\n'
                   % (len(names), body, _lib.ABI_VERSION))
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
