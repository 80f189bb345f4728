"""CPU: the C-ABI library builds, loads and exports every symbol include/uforecon_b200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from uforecon_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "uforecon_b200.h")).read()
    declared = set(re.findall(r"\b(ufo_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    raw = C.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(raw, s), f"{s} not exported"
    assert lib.ufo_abi_version() == 2


def test_struct_layout_matches_header():
    # pointer-sized / int32 fields only, natural alignment: sizes derive from the header by construction
    assert C.sizeof(_lib.UfoSceneDesc) == 5 * 4 + 4 + 4 * 8 + 6 * 8 + 9 * 4 + 4 + 9 * 8
    assert C.sizeof(_lib.UfoLoftrLayer) == 10 * 8
    assert C.sizeof(_lib.UfoMlp3) == 6 * 8
    assert C.sizeof(_lib.UfoWeightsDesc) == 2 * 80 + 3 * 48 + 3 * 8 + 8
    assert C.sizeof(_lib.UfoDebugTaps) == 11 * 8
    assert C.sizeof(_lib.UfoRenderOut) == 6 * 8
    assert C.sizeof(_lib.UfoPixelwiseNet) == 11 * 8 + 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    sm = C.c_int32()
    rc = lib.ufo_device_info(C.byref(sm), None, None)
    assert rc == -2
    assert b"no CPU fallback" in lib.ufo_last_error()
    d = _lib.UfoWeightsDesc()
    h = C.c_void_p()
    assert lib.ufo_weights_create(C.byref(d), C.byref(h), None) != 0
    with pytest.raises(_lib.UfoError):
        _lib.check(lib.ufo_scene_create(C.byref(_lib.UfoSceneDesc()), C.byref(h), None))


def test_product_path_does_not_import_oracle():
    import subprocess
    import sys
    code = ("import sys; import uforecon_b200.renderer, uforecon_b200.costvolume, uforecon_b200.dist; "
            "assert not any(m.startswith('oracle') for m in sys.modules), 'product path imports oracle'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for f in os.listdir(os.path.join(ROOT, "uforecon_b200")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "uforecon_b200", f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f
