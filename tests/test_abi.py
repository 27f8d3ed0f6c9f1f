"""CPU-only: the C-ABI library loads and exports every symbol include/rebert_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from robot_ebert_b200 import _native as nat

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(REPO, "include", "rebert_b200.h")).read()
    return sorted(set(re.findall(r"REBERT_API[^;(]*?\b(rebert_\w+)\s*\(", src)))


def test_header_declares_functions():
    names = _header_functions()
    assert len(names) >= 20 and "rebert_gemv_topk" in names and "rebert_gemm_topk" in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in _header_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_binding_matches_header():
    assert sorted(nat.exported_symbols()) == _header_functions()
    lib = nat.load()
    assert lib.rebert_abi_version() == nat.ABI_VERSION


def test_layout_and_candidates_host_logic():
    from robot_ebert_b200 import CatalogStore
    assert CatalogStore.layout(10, 1536, "bf16") == (1536, 10 * 1536 * 2)
    assert CatalogStore.layout(10, 1536, "fp32") == (1536, 10 * 1536 * 4)
    assert CatalogStore.layout(4, 32, "fp32")[0] == 32          # production collab shape: 128-byte rows
    assert CatalogStore.layout(4, 50, "bf16")[0] == 64          # padded to a power-of-two chunk count
    assert CatalogStore.layout(4, 2200, "bf16")[0] == 2304      # 9 chunks per lane
    lib = nat.load()
    assert [lib.rebert_candidates_for_k(k) for k in (1, 10, 16, 17, 48, 100, 112, 113, 240, 241)] == \
        [32, 32, 32, 64, 64, 128, 128, 256, 256, 0]
    with pytest.raises(ValueError):
        CatalogStore.layout(4, 0, "fp32")


def test_no_cpu_fallback():
    """Without a CUDA device the product path must raise, not compute on the CPU."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from robot_ebert_b200 import CatalogStore
    with pytest.raises(RuntimeError):
        CatalogStore.from_host(None, np.zeros((4, 8), np.float32))


def test_product_never_imports_oracle():
    """Product code must not import, call or execute anything under oracle/ (it is test infrastructure)."""
    pkg = os.path.join(REPO, "robot_ebert_b200")
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)|oracle/|oracle\.", re.M)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(root, f)).read()), f


def test_sorted_csr_validates_and_sorts():
    import numpy as np
    from robot_ebert_b200.catalog import sorted_csr
    ptr, col = sorted_csr([0, 3, 3, 5], [1, 5, 9, 2, 7])
    assert col.tolist() == [1, 5, 9, 2, 7]                       # already sorted per segment: untouched
    ptr, col = sorted_csr([0, 3, 3, 5], [9, 1, 5, 7, 2])
    assert col.tolist() == [1, 5, 9, 2, 7]
    for bad in ([1, 3], [0, 4, 2], [0, 2]):
        with pytest.raises(ValueError):
            sorted_csr(bad, [1, 2, 3])


def test_c_abi_argument_validation_without_a_gpu():
    """Every entry point validates its arguments before touching CUDA: status codes + thread-local message."""
    import ctypes as C
    lib = nat.load()
    cat = nat.Catalog(rows=0, inv_norm=0, norm64=0, n=10, row_base=0, d=32, ld=32, dtype=nat.F32, reserved=0)
    f = nat.Filter()
    assert lib.rebert_gemv_topk(C.byref(cat), None, C.byref(f), 32, None, 0, None, None) == nat.ERR_INVALID
    assert b"null" in lib.rebert_last_error()
    assert lib.rebert_finalize_topk(C.byref(cat), None, None, 32, 10, None, None, None, None, None) == nat.ERR_INVALID
    assert lib.rebert_catalog_layout(10, 32, 7, None, None) == nat.ERR_INVALID and b"dtype" in lib.rebert_last_error()
    assert lib.rebert_merge_topk(None, None, None, 1, 1, 1, 2, 1, 10, None, None, None, None) == nat.ERR_INVALID
    assert lib.rebert_exchange_buffer_bytes(0, 10, 0, 1) == 0 and lib.rebert_exchange_buffer_bytes(17, 10, 0, 1) == 0
    # per channel: results in LL form [2][world][2 * (2 k_max + 2)] + flags [2][world] + profiles [2][world][prof_len + 1] + flags [2][world]
    assert lib.rebert_exchange_buffer_bytes(8, 240, 0, 1) == (4 * 8 * 482 + 16 + 2 * 8 * 1 + 16) * 8
    assert lib.rebert_exchange_buffer_bytes(8, 240, 1536, 16) == 16 * (4 * 8 * 482 + 16 + 2 * 8 * 1537 + 16) * 8
    plan = nat.GemmPlan()
    assert lib.rebert_gemm_plan(1000, 16, 10, C.byref(plan)) == nat.ERR_UNSUPPORTED and b"too small" in lib.rebert_last_error()
    assert lib.rebert_gemm_plan(1_000_000, 4096, 100, C.byref(plan)) == nat.OK
    assert (plan.kc, plan.sample_rank, plan.cand_cap) == (128, 16, 4096) and plan.sample_rows % 256 == 0
    assert lib.rebert_gemm_plan(1_000_000, 64, 240, C.byref(plan)) == nat.OK and plan.kc == 320
    assert lib.rebert_gemm_plan(1_000_000, 64, 50, C.byref(plan)) == nat.OK and plan.kc == 96
    assert lib.rebert_gemm_plan(1_000_000, 4096, 5000, C.byref(plan)) == nat.ERR_UNSUPPORTED
    # int8 prefilter shadow: layout (1 byte per element), argument checks of the quantiser, kc = 256 only in the fast pass
    ld8, nbytes = C.c_int32(0), C.c_size_t(0)
    assert lib.rebert_catalog_layout(10, 1536, nat.I8, C.byref(ld8), C.byref(nbytes)) == nat.OK
    assert (ld8.value, nbytes.value) == (1536, 15360)
    assert lib.rebert_catalog_layout(10, 50, nat.I8, C.byref(ld8), None) == nat.OK and ld8.value == 64
    assert lib.rebert_catalog_quantize_i8(C.byref(cat), None, 32, None, None, None) == nat.ERR_INVALID
    cat8 = nat.Catalog(rows=128, inv_norm=128, norm64=128, n=10, row_base=0, d=32, ld=32, dtype=nat.I8, reserved=0)
    assert lib.rebert_catalog_quantize_i8(C.byref(cat8), 128, 32, 128, 128, None) == nat.ERR_INVALID and b"source dtype" in lib.rebert_last_error()
    assert lib.rebert_gemv_topk(C.byref(cat8), 128, C.byref(f), 32, 128, 1 << 20, 128, None) == nat.ERR_INVALID
    assert b"kc = 256" in lib.rebert_last_error()
    # host entry (single GPU and row shard alike): argument validation before any CUDA call
    x = nat.Exchange()
    x.world, x.rank, x.k_max, x.channels, x.seq = 2, 0, 240, 1, 1
    assert lib.rebert_recommend_host(C.byref(cat), None, None, None, 0, None, 0, None, 10, 32, 0, 0, None, 0, None, 0, None, C.byref(x),
                                     None, None, None, None, None) == nat.ERR_INVALID
    assert lib.rebert_recommend_device(C.byref(cat), None, None, None, None, 10, 32, None, 0, None, 0, None, None, None) == nat.ERR_INVALID
    assert lib.rebert_exchange_merge(C.byref(x), 10, 128, 128, 128, None) == nat.ERR_INVALID and b"null" in lib.rebert_last_error()
    assert lib.rebert_workspace_reset(None, 0, None) == nat.ERR_INVALID
    with pytest.raises(ValueError):
        nat.check(nat.ERR_INVALID)
    with pytest.raises(nat.NativeError):
        nat.check(nat.ERR_UNSUPPORTED)


def test_sorted_unique_i32_passes_sorted_input_through():
    from robot_ebert_b200.catalog import sorted_unique_i32
    a = np.array([1, 5, 9], dtype=np.int32)
    assert sorted_unique_i32(a) is a or np.shares_memory(sorted_unique_i32(a), a)          # no copy, no sort
    np.testing.assert_array_equal(sorted_unique_i32([9, 1, 5, 5, 1]), [1, 5, 9])
    np.testing.assert_array_equal(sorted_unique_i32(np.array([3, 3], dtype=np.int64)), [3])
    assert sorted_unique_i32(np.array([7], dtype=np.int64)).dtype == np.int32


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent CPU / PyTorch fallback: without the built .so the product import path raises."""
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", str(tmp_path / "librebert_b200.so"))
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        nat.load()
    monkeypatch.undo()
    assert nat.load().rebert_abi_version() == nat.ABI_VERSION


def test_header_is_plain_c_and_library_links_from_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (-pedantic) and a C program must link and run against the
    shared library without Python or C++ in the picture."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "c_embed")
    libdir = os.path.dirname(nat.LIB_PATH)
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"),
           os.path.join(REPO, "examples", "c_embed.c"), "-o", exe, "-L", libdir, "-lrebert_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "c_embed ok" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_pycall_module_validates_its_arguments_without_a_gpu():
    """csrc/pycall.c is the CPython side door into rebert_recommend_host (saves the ctypes call overhead per request).  Without a
    GPU the C entry refuses a NULL catalog before touching anything, which is enough to drive the module's own argument handling:
    buffers of the wrong item size, mismatched weights, negative addresses, wrong arity."""
    import ctypes as C
    from robot_ebert_b200.build import build_pycall
    build_pycall()                                                  # host compiler only; a no-op when the module is up to date
    nat._fast_call = False                                          # look it up (again) now that it exists
    fast = nat.fast_recommend_host()
    assert fast is not None, "robot_ebert_b200/_pycall*.so did not build or import"
    proof, info = nat.Proof(), nat.RequestInfo()
    q = np.zeros(32, dtype=np.float32)
    ex = np.arange(45, dtype=np.int32)
    lk = np.arange(5, dtype=np.int32)
    rows, scores = np.empty(16, dtype=np.int64), np.empty(16, dtype=np.float64)
    tail = (0, 10, 32, 64, 64, 1, 1, 1, 1, C.addressof(proof), 0, rows.ctypes.data, scores.ctypes.data, C.addressof(info), 0)
    rc, n = fast(0, q, None, None, ex, *tail)                       # cat = NULL: refused by the library, not by the module
    assert rc == nat.ERR_INVALID and n == 0 and b"null argument" in nat.load().rebert_last_error()
    assert fast(0, None, lk, np.ones(5, dtype=np.float32), None, *tail)[0] == nat.ERR_INVALID
    with pytest.raises(TypeError):
        fast(0, np.zeros(4, dtype=np.float64), None, None, None, *tail)      # 8-byte items
    with pytest.raises(ValueError):
        fast(0, None, lk, np.ones(3, dtype=np.float32), None, *tail)         # weights do not match liked rows
    with pytest.raises(TypeError):
        fast(0, "abc", None, None, None, *tail)
    with pytest.raises(OverflowError):
        fast(-1, q, None, None, None, *tail)
    with pytest.raises(TypeError):
        fast(0, q)
    with pytest.raises(ValueError):
        fast(0, q[::2], None, None, None, *tail)                              # not contiguous
