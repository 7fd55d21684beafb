"""CPU checks of the C-ABI boundary: libodf.so builds for sm_100a, loads without a GPU, exports
every symbol include/odf.h declares (and the ctypes table mirrors the header), and refuses to
compute without a device instead of falling back."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "odf.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(odf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libodf.so does not export %s" % n


def test_ctypes_table_matches_header(lib):
    from odf import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "odf.h"\nint main(void){return ODF_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                           "-o", str(tmp_path / "t.o")])


def test_layout_helpers(lib):
    assert lib.odf_operand_pitch(1, 0) == 96 and lib.odf_operand_pitch(1024, 0) == 1088
    assert lib.odf_operand_pitch(1, 1) == 192 and lib.odf_operand_pitch(1024, 1) == 1152 and lib.odf_operand_pitch(1025, 1) == 1216
    assert lib.odf_operand_bytes(10, 1024, 1) == 10 * 1152 * 2 and lib.odf_operand_bytes(10, 1024, 0) == 10 * 1088 * 4
    assert lib.odf_default_kind() in (0, 1) and lib.odf_set_default_kind(7) != 0
    assert lib.odf_pad_rows(1) == 128 and lib.odf_pad_rows(128) == 128 and lib.odf_pad_rows(129) == 256
    assert lib.odf_tpad(1) == 16 and lib.odf_tpad(16) == 16 and lib.odf_tpad(17) == 32 and lib.odf_tpad(33) == -1
    assert lib.odf_version() >= 100
    for (n, m, d) in [(128, 128, 32), (1000000, 10000, 1024), (10000, 1000000, 1024), (300, 70000, 256)]:
        s = lib.odf_tile_splits(n, m, d, 1)
        tiles = (m + 127) // 128
        per = (tiles + s - 1) // s
        assert s >= 1 and (tiles + per - 1) // per == s        # no empty split
    # K panel: 3 bytes per (padded) kernel value (fp16 hi plane + one-byte residual plane); row ranges never empty; the CTA-pair tile wants >= 8192 rows
    assert lib.odf_panel16_bytes(131072, 10000) == 131072 * 10112 * 3 and lib.odf_panel16_bytes(100, 100) == 128 * 128 * 3
    for (n, m) in [(131072, 10000), (82496, 10000), (131072, 30000), (300, 200), (64, 5), (9000, 333)]:
        s = lib.odf_panel16_splits(n, m)
        stages = (n + 63) // 64
        per = (stages + s - 1) // s
        assert 1 <= s <= 32 and (stages + per - 1) // per == s
    assert lib.odf_tile_pair_eligible(8192) == 1 and lib.odf_tile_pair_eligible(8191) == 0
    from odf import _lib
    assert lib.odf_workspace_bytes(_lib.ODF_OP_MMV, 1000, 100, 64, 21) > 0
    assert lib.odf_cg_workspace_bytes(1000, 21) > 0


def test_sass_is_blackwell_native():
    """tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG must be in the binary."""
    so = os.path.join(ROOT, "online-detection_b200", "libodf.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    """Without a device the product path must fail loudly, not compute on the host."""
    from odf import ops, _lib
    x = torch.zeros(4, 8)
    with pytest.raises(ValueError, match="CUDA"):
        ops.Prepared(x)
    buf = (ctypes.c_float * 64)()
    rc = lib.odf_zscore(ctypes.cast(buf, ctypes.c_void_p), 2, 8, 8, None, 1.0, None)
    assert rc == -2 and b"" != lib.odf_last_error()
    from odf import GaussianKernel, InCoreFalkon
    m = InCoreFalkon(kernel=GaussianKernel(5.0), penalty=1e-3, M=4)
    with pytest.raises((ValueError, RuntimeError, _lib.OdfError)):
        m.fit(torch.randn(16, 8), torch.ones(16))


def test_missing_library_is_an_error(monkeypatch):
    from odf import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libodf.so")
    with pytest.raises(_lib.OdfError, match="no CPU/PyTorch fallback"):
        _lib.load()
