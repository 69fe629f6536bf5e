"""CPU: the C-ABI libraries load and export every symbol include/figbird_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "figbird_b200.h")
LIBS = {"product": os.path.join(ROOT, "figbird_b200", "_build", "libfigbird_b200.so"), "oracle": os.path.join(ROOT, "oracle", "_build", "libfb_oracle.so")}


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_two_levels():
    names = declared_functions()
    for need in ("fb_fillgaps_main", "fb_ctx_create", "fb_model_upload", "fb_batch_upload", "fb_em_run"):
        assert need in names


@pytest.mark.parametrize("which", sorted(LIBS))
def test_library_exports_every_declared_symbol(which):
    path = LIBS[which]
    assert os.path.exists(path), "%s missing: run __graft_entry__.build()" % path
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), "%s does not export %s" % (which, name)
    lib.fb_engine_name.restype = ctypes.c_char_p
    assert lib.fb_engine_name().decode() == ("cuda-sm100a" if which == "product" else "oracle-cpu")


def test_python_binding_covers_the_header():
    from figbird_b200 import capi
    assert sorted(capi.EXPORTS) == declared_functions()


def test_product_library_has_no_cpu_engine_symbols():
    """The oracle engine must not be linked into the product (no CPU fallback)."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", LIBS["product"]], stdout=subprocess.PIPE).stdout.decode()
    assert "oracle" not in out.lower()
    # ... and it does carry the sm_100a kernels
    sass = subprocess.run(["cuobjdump", "-lelf", LIBS["product"]], stdout=subprocess.PIPE, stderr=subprocess.STDOUT).stdout.decode()
    assert "sm_100a" in sass, sass[-500:]
    syms = subprocess.run(["nm", LIBS["product"]], stdout=subprocess.PIPE).stdout.decode()
    assert "fb_em_kernel" in syms and "fb_flank_kernel" in syms


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from figbird_b200 import capi
    with pytest.raises(RuntimeError):
        capi.Engine(0)
