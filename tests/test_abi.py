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


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from vitta_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.VittaError):
        _lib.load()
