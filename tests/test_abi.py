"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/suitesparse_b200.h
declares, and the restated CHOLMOD struct layouts match the reference headers (when /root/reference is present)."""
import ctypes as C, os, re, subprocess, sys
import pytest
from conftest import REPO, B200_LIB

HEADER = os.path.join(REPO, "include", "suitesparse_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:cholmod_l_|ssb200_)[a-z_A-Z0-9]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(B200_LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(B200_LIB)
    syms = declared_symbols()
    assert {"cholmod_l_super_numeric", "cholmod_l_super_lsolve", "cholmod_l_super_ltsolve", "cholmod_l_gpu_probe",
            "ssb200_plan_create", "ssb200_factorize", "ssb200_solve"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_ctypes_struct_sizes_match_the_abi():
    from suitesparse_b200 import cholmod_host as H
    # sizes measured on the reference build
    assert (C.sizeof(H.Common), C.sizeof(H.Factor), C.sizeof(H.Sparse), C.sizeof(H.Dense)) == (2664, 208, 88, 56)


@pytest.mark.skipif(not os.path.isdir("/root/reference/CHOLMOD/Include"), reason="reference headers not on this box")
def test_struct_layout_against_reference_headers(tmp_path):
    fields = {"common": ["supernodal", "nrelax", "zrelax", "quick_return_if_not_posdef", "print", "error_handler", "nmethods", "method",
                         "postorder", "Flag", "Iwork", "itype", "status", "fl", "lnz", "malloc_count", "blas_ok", "useGPU", "gpuKernelTime",
                         "gpuNumKernelLaunches", "cublasHandle", "dev_mempool", "devBuffSize", "ibuffer", "syrkStart"],
              "factor": ["n", "minor", "Perm", "x", "nsuper", "ssize", "xsize", "maxcsize", "maxesize", "super", "pi", "px", "s", "is_ll", "is_super", "xtype", "useGPU"],
              "sparse": ["nrow", "p", "i", "nz", "x", "stype", "itype", "xtype", "sorted", "packed"],
              "dense": ["nrow", "ncol", "nzmax", "d", "x", "xtype", "dtype"]}
    body = ["#include <stdio.h>", "#include <stddef.h>", '#include "cholmod.h"', "#define SSB200_NO_DROPIN_PROTOTYPES",
            '#include "suitesparse_b200.h"', "int main(void){int bad=0;"]
    for t, fs in fields.items():
        body.append(f"if(sizeof(cholmod_{t})!=sizeof(ssb_cholmod_{t})){{printf(\"size {t}\\n\");bad=1;}}")
        for f in fs:
            body.append(f"if(offsetof(cholmod_{t},{f})!=offsetof(ssb_cholmod_{t},{f})){{printf(\"{t}.{f}\\n\");bad=1;}}")
    body.append("if(offsetof(cholmod_common,cholmod_gpu_potrf_calls)!=offsetof(ssb_cholmod_common,gpu_potrf_calls)){printf(\"tail\\n\");bad=1;}")
    body.append("return bad;}")
    src = tmp_path / "abi.c"; src.write_text("\n".join(body))
    inc = tmp_path / "inc"; inc.mkdir()
    (inc / "cholmod_config.h").write_text("#define NPARTITION\n")
    for n, m in (("cholmod_export.h", "CHOLMOD_EXPORT"), ("SuiteSparse_export.h", "SUITESPARSECONFIG_EXPORT")):
        (inc / n).write_text(f"#define {m}\n")
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", str(inc), "-I/root/reference/CHOLMOD/Include", "-I/root/reference/SuiteSparse_config",
                           "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout


def test_missing_library_fails_loudly(monkeypatch):
    from suitesparse_b200 import cholmod_host as H
    monkeypatch.setattr(H, "_b200_handle", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.load_b200("/nonexistent/libsuitesparse_b200.so")


def test_invalid_arguments_without_gpu():
    """Argument checks of the drop-in entry points run before any device work (cholmod_super_numeric.c:120-175,
    cholmod_super_solve.c:59-90), so their error codes can be checked on a CPU-only box."""
    from suitesparse_b200 import cholmod_host as H
    lib = C.CDLL(B200_LIB)
    cm = H.Common(); cm.itype = H.CHOLMOD_LONG; cm.dtype = 0; cm.status = 0
    L = H.Factor(); A = H.Sparse()
    f = lib.cholmod_l_super_numeric
    f.argtypes = [C.c_void_p] * 5
    assert f(None, None, None, None, None) == 0                      # NULL Common
    assert f(None, None, None, C.byref(L), C.byref(cm)) == 0 and cm.status == H.CHOLMOD_INVALID     # NULL A
    cm.status = 0; cm.itype = H.CHOLMOD_INT
    assert f(C.byref(A), None, None, C.byref(L), C.byref(cm)) == 0 and cm.status == H.CHOLMOD_INVALID   # wrong itype
    # upper-stored symmetric matrix is rejected
    import numpy as np
    x = np.ones(1); p = np.zeros(2, dtype=np.int64); i = np.zeros(1, dtype=np.int64)
    A.nrow = A.ncol = 1; A.nzmax = 1; A.p = p.ctypes.data; A.i = i.ctypes.data; A.x = x.ctypes.data
    A.stype = 1; A.itype = H.CHOLMOD_LONG; A.xtype = H.CHOLMOD_REAL; A.sorted = A.packed = 1
    L.n = 1; L.xtype = H.CHOLMOD_PATTERN; L.is_super = 1; L.itype = H.CHOLMOD_LONG
    cm.status = 0; cm.itype = H.CHOLMOD_LONG
    assert f(C.byref(A), None, None, C.byref(L), C.byref(cm)) == 0 and cm.status == H.CHOLMOD_INVALID
    # dimension mismatch
    A.stype = -1; L.n = 2; cm.status = 0
    assert f(C.byref(A), None, None, C.byref(L), C.byref(cm)) == 0 and cm.status == H.CHOLMOD_INVALID
    # L not supernodal
    L.n = 1; L.is_super = 0; cm.status = 0
    assert f(C.byref(A), None, None, C.byref(L), C.byref(cm)) == 0 and cm.status == H.CHOLMOD_INVALID
    # solve: too-small workspace E (Tcov/raw_factor.c:318-326)
    g = lib.cholmod_l_super_lsolve; g.argtypes = [C.c_void_p] * 4
    X = H.Dense(); E = H.Dense()
    L.is_super = 1; L.is_ll = 1; L.xtype = H.CHOLMOD_REAL; L.x = x.ctypes.data; L.maxesize = 4
    X.nrow = 1; X.ncol = 1; X.d = 1; X.nzmax = 1; X.x = x.ctypes.data; X.xtype = H.CHOLMOD_REAL
    E.nrow = 1; E.ncol = 1; E.d = 1; E.nzmax = 1; E.x = x.ctypes.data; E.xtype = H.CHOLMOD_REAL
    cm.status = 0
    assert g(C.byref(L), C.byref(X), C.byref(E), C.byref(cm)) == 0 and cm.status == H.CHOLMOD_INVALID
    # n-by-0 right-hand side returns TRUE immediately (Tcov/raw_factor.c:339-345)
    X.ncol = 0; E.nzmax = 0; cm.status = -1
    assert g(C.byref(L), C.byref(X), C.byref(E), C.byref(cm)) == 1 and cm.status == 0


def test_interposed_host_functions_reach_the_host_library():
    """cholmod_l_free_factor and cholmod_l_solve are interposed and hand over to the host library's own definitions.  With
    both libraries dlopen'ed (ctypes) RTLD_NEXT alone does not find them - the walk over the loaded objects must.
    (In a child process: loading the B200 library globally rebinds the host library's hot-path calls for the whole process.)"""
    from conftest import REF_LIB
    if not os.path.exists(REF_LIB):
        pytest.skip("reference build (host libcholmod) not present")
    code = """
import numpy as np, scipy.sparse as sp, sys
sys.path.insert(0, %r)
from suitesparse_b200 import gen, cholmod_host as H
ch = H.Cholmod(gpu=True)
A, p = gen.make_problem("lap7", 5)
S = ch.sparse(A, +1)
L = ch.analyze(S, p)
before = ch.cm.malloc_count
ch.free_factor(L)                                   # interposed -> host: the factor's arrays are really freed
assert ch.cm.malloc_count < before and ch.cm.status == 0, (before, ch.cm.malloc_count, ch.cm.status)
# a simplicial factor (CPU, host library) solved through the interposed cholmod_l_solve: not the fast path -> host's own
L2 = ch.analyze(S, p, supernodal=H.CHOLMOD_SIMPLICIAL)
assert ch.factorize(S, L2) == 1
x = ch.solve(L2, np.ones(A.shape[0]))
Af = A + sp.triu(A, 1).T
assert np.linalg.norm(Af @ x - 1.0) / np.sqrt(A.shape[0]) < 1e-12
ch.free_factor(L2)
print("OK")
""" % REPO
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_header_is_plain_c_and_example_links(tmp_path):
    """include/suitesparse_b200.h compiles as C99 and examples/plain_demo.c links against the library (no run: no GPU here)."""
    exe = tmp_path / "plain_demo"
    csrc = os.path.join(REPO, "suitesparse_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(REPO, "include"),
                           os.path.join(REPO, "examples", "plain_demo.c"), "-L", csrc, "-lsuitesparse_b200",
                           f"-Wl,-rpath,{csrc}", "-o", str(exe)])
    assert exe.exists()


@pytest.mark.gpu
def test_plain_c_example_runs(tmp_path):
    exe = tmp_path / "plain_demo"
    csrc = os.path.join(REPO, "suitesparse_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(REPO, "include"), os.path.join(REPO, "examples", "plain_demo.c"),
                           "-L", csrc, "-lsuitesparse_b200", f"-Wl,-rpath,{csrc}", "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    import numpy as np
    A = np.array([[4, 1, 0], [1, 5, 2], [0, 2, 6]], dtype=float)
    x = np.linalg.solve(A, np.array([1.0, 2.0, 3.0]))
    got = [float(v) for v in re.search(r"x = ([-0-9.e]+) ([-0-9.e]+) ([-0-9.e]+)", r.stdout).groups()]
    assert np.allclose(got, x, atol=1e-5) and "L(0,0)=2.000000" in r.stdout


def test_first_touch_threads_do_not_disturb_the_copies():
    """The threads that fault a fresh L->x in (atomic `or 0` on one word per page) run against the threaded copies of the
    staging ring; every byte must still arrive.  Host-only code of the library: no GPU needed."""
    import mmap
    import numpy as np
    lib = C.CDLL(B200_LIB)
    lib.ssb200_debug_first_touch_copy.restype = C.c_int
    lib.ssb200_debug_first_touch_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t]
    nbytes = 192 << 20
    rng = np.random.default_rng(5)
    src = rng.integers(0, 2**63, size=nbytes // 8, dtype=np.int64)
    os.environ["SSB200_FIRST_TOUCH_MIN_MB"] = "0"; os.environ["SSB200_FIRST_TOUCH_THREADS"] = "6"
    try:
        for rep in range(3):
            buf = mmap.mmap(-1, nbytes + 4096)                       # fresh, never touched pages every time
            dst = np.frombuffer(buf, dtype=np.int64, count=nbytes // 8, offset=8 * (1 + rep))   # not page aligned
            assert lib.ssb200_debug_first_touch_copy(dst.ctypes.data, src.ctypes.data, nbytes, 4, 8 << 20) == 0
            assert np.array_equal(dst, src)
            del dst; buf.close()
    finally:
        os.environ.pop("SSB200_FIRST_TOUCH_MIN_MB", None); os.environ.pop("SSB200_FIRST_TOUCH_THREADS", None)
