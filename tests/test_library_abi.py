"""CPU: libpndf.so loads and exports every symbol include/pndf.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "posendf_b200", "libpndf.so")):
        g.build()
    from posendf_b200 import _lib
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pndf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pndf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    from posendf_b200 import _lib
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pndf.h but not exported by libpndf.so"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in posendf_b200/_lib.py"


def test_config_validation_and_param_count(lib):
    from posendf_b200 import _lib
    n = C.c_size_t()
    assert lib.pndf_param_count(C.byref(_lib.make_config()), C.byref(n)) == 0 and n.value == 1365565
    assert lib.pndf_param_count(C.byref(_lib.make_config(use_enc=False)), C.byref(n)) == 0
    assert n.value == 1365565 - 3516 - 126 * 256 + 84 * 256
    assert lib.pndf_param_count(C.byref(_lib.make_config(dims=(256, 256))), C.byref(n)) != 0
    assert b"amass.yaml" in lib.pndf_last_error()
    assert lib.pndf_param_count(C.byref(_lib.make_config(use_enc=True, in_dim=84)), C.byref(n)) != 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from posendf_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(device=0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """no libpndf.so -> RuntimeError naming the build command, never a silent eager / CPU path"""
    from posendf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libpndf.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_host_side_layout_constants_match_the_kernel_header():
    """posendf_b200/train.py addresses the training exports by column: its constants must be the kernel's."""
    import re
    from posendf_b200 import train
    src = open(os.path.join(ROOT, "posendf_b200", "csrc", "pndf_kernel.cuh")).read()

    def const(name):
        m = re.search(r"constexpr int %s\s*=\s*([0-9]+)" % name, src)
        assert m, name
        return int(m.group(1))

    assert train.DUMP_ROWS == const("kDumpRows")
    assert train.ENC_FLOATS == const("kEncFloats")
    widths = [256, 512, 1024, 512, 256, 64]
    assert sum(widths) == const("kUnits")
    # layer inputs z_0..z_6 are contiguous from column 0, adjoints a_5..a_0 follow, then the z_0 adjoint
    off = 0
    for (c0, w), width in zip(train.Z_ROWS, [128] + widths):
        assert c0 == off
        assert w in (None, width)
        off += width
    assert off == train.Z_END
    for l in (5, 4, 3, 2, 1, 0):
        assert train.A_ROWS[l] == (off, widths[l])
        off += widths[l]
    assert off == train.A_END == train.G0_ROW and train.G0_ROW + 128 == train.DUMP_ROWS
    # the dump offsets the kernel passes to dump_rows()
    for c0 in [r[0] for r in train.Z_ROWS[1:]] + [r[0] for r in train.A_ROWS] + [train.G0_ROW]:
        assert re.search(r"dump_rows\(dbg_[ps], %d(\s|,|\+)" % c0, src), c0


def test_stateless_entry_points_reject_bad_arguments_before_touching_cuda():
    import ctypes as C
    from posendf_b200 import _lib
    lib = _lib.load()
    dummy = C.c_void_p(16)
    # fewer than 5 database poses / null pointers / unknown metric
    assert lib.pndf_knn_exact(0, dummy, 4, dummy, 3, 0, 0, dummy, dummy, None) != 0
    assert b"pndf_knn_exact" in lib.pndf_last_error()
    assert lib.pndf_knn_exact(0, None, 4, dummy, 100, 0, 0, dummy, dummy, None) != 0
    assert lib.pndf_knn_exact(0, dummy, 4, dummy, 100, 7, 0, dummy, dummy, None) != 0
    assert lib.pndf_knn_exact(0, dummy, 4, C.c_void_p(20), 100, 0, 0, dummy, dummy, None) != 0      # misaligned database
    assert lib.pndf_knn_rerank(0, dummy, 4, dummy, dummy, 3, 0, 0, dummy, dummy, None) != 0          # K < 5
    assert lib.pndf_softplus_adjoint(0, dummy, dummy, dummy, 5504, dummy, None, 100.0, 8, 6, dummy, None) != 0   # n % 4 != 0
    assert lib.pndf_axis_angle_to_quaternion(0, None, 3, dummy, None) != 0
    assert lib.pndf_quaternion_to_axis_angle(0, dummy, -1, dummy, None) != 0
    # empty work is a no-op, not an error
    assert lib.pndf_knn_exact(0, None, 0, None, 0, 0, 0, None, None, None) == 0
    assert lib.pndf_axis_angle_to_quaternion(0, None, 0, None, None) == 0
