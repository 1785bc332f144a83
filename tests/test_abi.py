"""CPU: the C-ABI shared library builds, loads without a GPU and exports every symbol include/vitta_b200.h declares;
argument validation returns error codes instead of crashing.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vitta_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vitta_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from vitta_b200 import build
    build.build()
    from vitta_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    from vitta_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_lib.exported_symbols()) == names, "ctypes binding and header are out of sync"


def test_version_and_errors_without_gpu(lib):
    from vitta_b200 import _lib
    assert lib.vitta_version() >= 100
    ch = _lib.VittaChunking()
    assert lib.vitta_stats_chunking(0, 4, 1, 1, ctypes.byref(ch)) == -1      # VITTA_E_BADARG
    assert b"bad shape" in lib.vitta_last_error()
    assert lib.vitta_stats_chunking(128 * 196, 1024, 1, 128, ctypes.byref(ch)) == 0
    assert ch.frame_rows == 196 and ch.chunks_per_frame * ch.chunk_rows >= 196
    assert ch.n_entries == 128 * ch.chunks_per_frame
    assert lib.vitta_stats_chunking(128, 6, 49, 1, ctypes.byref(ch)) == 0
    assert ch.frame_rows == 128 * 49 and ch.n_entries >= 1
    # null pointers are rejected before any launch
    assert lib.vitta_tam_fwd(None, None, None, None, 1, 8, 49, 64, None) == -1


def test_chunk_counts_cover_every_row(lib):
    from vitta_b200 import _lib
    for rows, c, frames in [(25088, 256, 1), (25088, 256, 128), (6272, 2048, 128), (100352, 256, 128), (37, 8, 1),
                            (3136 * 4, 64, 4), (1, 4, 1)]:
        ch = _lib.chunking(rows, c, 1, frames)
        total = 0
        for e in range(ch.n_entries):
            j = e % ch.chunks_per_frame
            total += min(ch.chunk_rows, ch.frame_rows - j * ch.chunk_rows)
            assert min(ch.chunk_rows, ch.frame_rows - j * ch.chunk_rows) > 0
        assert total == rows, (rows, c, frames, total)


def test_chunking_fills_whole_waves(lib):
    """The chunk count of the streaming kernels (one chunk = one CTA = one statistics entry) is picked against whole
    waves of 148 SMs x 4 resident CTAs: the TANet layer shapes (128 frames) must not leave a half-empty last wave."""
    from vitta_b200 import _lib
    slots = 148 * 4
    for rows_per_frame, c in [(3136, 64), (3136, 256), (3136, 128), (784, 128), (784, 512), (784, 256), (49, 2048)]:
        ch = _lib.chunking(128 * rows_per_frame, c, 1, 128)
        lanes = min(32, 1 << ((c // 4).bit_length() - 1))
        ctas = ch.n_entries * ((c // 4 + lanes - 1) // lanes)
        waves = -(-ctas // slots)
        assert ctas / (waves * slots) >= 0.93, (rows_per_frame, c, ctas)
        assert ch.chunk_rows * ch.chunks_per_frame >= rows_per_frame > ch.chunk_rows * (ch.chunks_per_frame - 1)


def test_host_side_planners_without_gpu(lib):
    """Workspace / chunk planners are pure host functions: sane, positive and consistent without a device."""
    # TAM backward: at most 64 rows per chunk, small late-stage maps still yield >= 32 chunks per sample and channel tile
    for hw, c in [(3136, 64), (784, 128), (196, 256), (49, 512), (5, 8)]:
        n = lib.vitta_tam_num_chunks(hw, c)
        assert n >= 1 and -(-hw // n) <= 64
    assert lib.vitta_tam_num_chunks(196, 256) >= 30
    # weight-gradient workspace: splits x (Cout x KH*KW x Cin weight partials + Cout bias partials) floats, also in the
    # three-taps-per-item mode (Cin = 64, 3x3)
    for f, h, cin, cout, k, s in [(128, 56, 64, 64, 3, 1), (128, 56, 64, 256, 1, 1), (128, 14, 256, 256, 3, 1),
                                  (128, 28, 128, 128, 3, 2), (2, 9, 8, 24, 3, 1)]:
        n = lib.vitta_conv2d_wgrad_ws_floats(f, h, h, cin, cout, k, k, s, k // 2)
        assert n > 0 and n % (cout * k * k * cin + cout) == 0
    assert lib.vitta_conv2d_wgrad_ws_floats(0, 56, 56, 64, 64, 3, 3, 1, 1) == -1
    assert lib.vitta_bn_act_bwd_ws_floats(128, 3136, 64) > 0
    assert lib.vitta_gemm_set_operand_form(3) == -1 and lib.vitta_gemm_set_operand_form(0) == 0


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from vitta_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.VittaError):
        _lib.load()


def test_ctypes_signatures_match_the_header_prototypes():
    """Every binding in _lib._SIGNATURES has the arity and the argument classes (pointer / struct by value / int / int64 /
    float) of its prototype in include/vitta_b200.h -- a drifted ctypes signature would corrupt the call silently."""
    import ctypes as C
    import re
    from vitta_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "vitta_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)

    def klass_c(p):
        p = p.strip()
        if "*" in p:
            return "ptr"
        t = p.rsplit(" ", 1)[0].replace("const", "").strip()
        return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "float": "f32", "VittaBN": "VittaBN"}.get(t, t)

    def klass_py(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int: "i32", C.c_int32: "i32", C.c_int64: "i64", C.c_float: "f32"}.get(t, getattr(t, "__name__", str(t)))

    for name, (res, args) in _lib._SIGNATURES.items():
        m = re.search(r"^\s*(?:const\s+)?([\w]+\s*\*?)\s*" + name + r"\s*\(([^;{]*?)\)\s*;", hdr, re.S | re.M)
        assert m, "no prototype for " + name
        params = [p for p in m.group(2).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))
        for i, (pc, pa) in enumerate(zip(params, args)):
            assert klass_c(pc) == klass_py(pa), (name, i, pc.strip(), pa)
        ret = m.group(1).replace(" ", "")
        want = {"int": "i32", "int64_t": "i64", "char*": "ptr"}[ret]
        assert klass_py(res) == want, (name, ret, res)
