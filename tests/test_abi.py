"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/longbow_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "longbow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:lb|faiss_gpu)_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from longbow_b200 import _lib
    names = _declared()
    assert len(names) >= 40
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_six_faiss_symbols_present():
    # internal/gpu/faiss_gpu.go:16-21
    from longbow_b200 import _lib
    lib = _lib.load()
    for n in ("faiss_gpu_resources_new", "faiss_gpu_resources_free", "faiss_gpu_index_flat_l2_new",
              "faiss_gpu_index_flat_l2_free", "faiss_gpu_index_add", "faiss_gpu_index_search"):
        assert getattr(lib, n) is not None


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    from longbow_b200 import _lib, gpu
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.lb_index_create(0, 128, 0, 0, C.byref(h))
    assert rc == _lib.LB_ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.lb_last_error()
    assert lib.faiss_gpu_resources_new(0) is None
    with pytest.raises(RuntimeError, match="failed to initialize GPU resources"):
        gpu.NewIndexWithConfig(gpu.GPUConfig(0, 128))
    with pytest.raises(_lib.LongbowError):
        gpu.DenseIndex(128)


def test_argument_validation_before_device():
    from longbow_b200 import _lib, gpu
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.lb_index_create(0, -1, 0, 0, C.byref(h)) == _lib.LB_ERR_INVALID
    assert lib.lb_index_create(0, 16, 2, 1, C.byref(h)) == _lib.LB_ERR_UNSUPPORTED  # int8 cosine: no kernel
    assert lib.lb_index_create(0, 16, 9, 0, C.byref(h)) == _lib.LB_ERR_UNSUPPORTED
    assert lib.lb_pq_create(0, b"\0" * 4, 4, C.byref(h)) == _lib.LB_ERR_INVALID  # persistence.go:40
    with pytest.raises(ValueError, match="dimension must be positive"):  # gpu_test.go:49-55
        gpu.NewIndexWithConfig(gpu.GPUConfig(0, -1))
    lib.lb_index_free(None)  # no-op
    lib.faiss_gpu_index_flat_l2_free(None)
    lib.faiss_gpu_resources_free(None)


def test_pack_bitmap_layout():
    from longbow_b200.gpu import pack_bitmap
    m = np.zeros(130, bool)
    m[[0, 63, 64, 129]] = True
    w = pack_bitmap(m)
    assert w.dtype == np.uint64 and w.size == 3
    assert w[0] == (1 | (1 << 63)) and w[1] == 1 and w[2] == 2


def test_dense_words_from_id_sets():
    """The roaring -> dense conversion (internal/query/bitmap.go:95-100 feeding the allow / tombstone bitmaps)."""
    from longbow_b200.gpu import dense_words, pack_bitmap
    w = dense_words([0, 63, 64, 129, 129, 500], 130)          # a duplicate and an id past the index
    assert w.dtype == np.uint64 and w.tolist() == [1 | (1 << 63), 1, 2]
    assert dense_words([], 65).tolist() == [0, 0]
    rng = np.random.default_rng(0)
    mask = rng.random(1000) < 0.3
    assert np.array_equal(dense_words(np.flatnonzero(mask).astype(np.uint32), 1000), pack_bitmap(mask))


def test_pq_blob_roundtrip_host():
    # internal/pq/persistence.go:15-80 format; parsing errors surface before any device work
    import struct
    from longbow_b200 import pq
    with pytest.raises(ValueError, match="too short"):
        pq.PQEncoder.Deserialize(b"123")
    with pytest.raises(ValueError, match="invalid PQ parameters"):
        pq.PQEncoder.Deserialize(struct.pack("<III", 10, 3, 256) + b"\0" * 16)
    with pytest.raises(ValueError, match="size mismatch"):
        pq.PQEncoder.Deserialize(struct.pack("<III", 8, 2, 256) + b"\0" * 16)


def test_simd_mirror_validation():
    from longbow_b200 import simd
    q = np.zeros(4, np.float32)
    with pytest.raises(simd.SimdError, match="results length mismatch"):
        simd.EuclideanDistanceBatchFlat(q, np.zeros(8, np.float32), 2, 4, np.zeros(3, np.float32))
    with pytest.raises(simd.SimdError, match="flatVectors too small"):
        simd.EuclideanDistanceBatchFlat(q, np.zeros(4, np.float32), 2, 4, np.zeros(2, np.float32))
    with pytest.raises(simd.SimdError, match="query dimension mismatch"):
        simd.EuclideanDistanceBatchFlat(np.zeros(3, np.float32), np.zeros(8, np.float32), 2, 4, np.zeros(2, np.float32))
    with pytest.raises(simd.SimdError, match="invalid m"):
        simd.ADCDistanceBatch(np.zeros(256, np.float32), np.zeros(4, np.uint8), 0, np.zeros(1, np.float32))
    simd.EuclideanDistanceBatchFlat(q, np.zeros(0, np.float32), 0, 4, np.zeros(0, np.float32))  # n == 0: no-op


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 on its own (no C++ / CUDA / torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "longbow_b200.h")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])


def _split_top(args: str):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def _call_args(src: str, start: int):
    """Text between the parenthesis at src[start] and its match."""
    depth = 0
    for i in range(start, len(src)):
        if src[i] == "(":
            depth += 1
        elif src[i] == ")":
            depth -= 1
            if depth == 0:
                return src[start + 1:i]
    raise AssertionError("unbalanced call")


def test_go_shim_calls_match_the_header():
    """The Go toolchain is absent here, so the cgo shim (go/*.go) cannot be compiled; this keeps it from drifting:
    every C.<symbol>(...) call names a function the header declares and passes as many arguments as it takes."""
    hdr = open(os.path.join(ROOT, "include", "longbow_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    nparams = {}
    for m in re.finditer(r"\b((?:lb|faiss_gpu)_[a-z0-9_]+)\s*\(", hdr):
        args = _call_args(hdr, m.end() - 1).strip()
        nparams[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    calls = 0
    for fn in sorted(os.listdir(os.path.join(ROOT, "go"))):
        if not fn.endswith(".go"):
            continue
        src = open(os.path.join(ROOT, "go", fn)).read()
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\bC\.((?:lb|faiss_gpu)_[a-z0-9_]+)\s*\(", src):
            name = m.group(1)
            assert name in nparams, f"{fn}: C.{name} is not declared in include/longbow_b200.h"
            got = len(_split_top(_call_args(src, m.end() - 1)))
            assert got == nparams[name], f"{fn}: C.{name} called with {got} arguments, the header takes {nparams[name]}"
            calls += 1
    assert calls >= 15


def test_go_shim_calls_only_defined_functions():
    """A crude stand-in for `go vet`: every bare function call in go/*.go is a Go builtin / conversion or a function
    defined somewhere in the package (catches helpers that were renamed away, e.g. lastErr)."""
    builtins = {"len", "cap", "make", "append", "copy", "new", "panic", "delete", "min", "max", "func", "if", "for", "switch",
                "return", "import", "var", "const", "type", "defer", "go", "range", "select", "case", "else", "int", "int32", "int64", "uint32", "uint64", "float32", "float64", "byte", "string", "bool"}
    srcs = {}
    for fn in sorted(os.listdir(os.path.join(ROOT, "go"))):
        if fn.endswith(".go"):
            src = open(os.path.join(ROOT, "go", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)       # the cgo preamble
            src = re.sub(r"//[^\n]*", "", src)
            srcs[fn] = re.sub(r'"(?:[^"\\\n]|\\.)*"', '""', src)   # string literals
    defined = set()
    for src in srcs.values():
        defined |= set(re.findall(r"\bfunc\s+(?:\([^)]*\)\s*)?([A-Za-z_][A-Za-z0-9_]*)\s*\(", src))
        defined |= set(re.findall(r"\btype\s+([A-Za-z_][A-Za-z0-9_]*)\b", src))
        defined |= set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s+func\(", src))         # function-typed parameters
        defined |= set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*:=\s*func\(", src))    # local closures
    defined |= {"GPUConfig", "Index"}   # declared by the package this file joins (internal/gpu/interface.go:10-19,21-30)
    for fn, src in srcs.items():
        for m in re.finditer(r"(?<![.\w])([A-Za-z_][A-Za-z0-9_]*)\s*\(", src):
            name = m.group(1)
            if name in builtins or name in defined:
                continue
            line = src[:m.start()].count("\n") + 1
            raise AssertionError(f"go/{fn}:{line}: call of undefined function {name}()")
