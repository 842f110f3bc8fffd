"""BASELINE config 1: the reference's own demo program (CHOLMOD/Demo/cholmod_l_demo.c, built unmodified by
`make -C oracle demo`) on a bundled Demo matrix — as is on the CPU (plumbing), and with the B200 library interposed by
LD_PRELOAD exactly as INTEGRATION.md describes for an unmodified C program."""
import os, re, subprocess
import pytest
from conftest import REPO, B200_LIB

DEMO = os.path.join(REPO, "oracle", "_ref", "cholmod_l_demo")
MAT = os.path.join(REPO, "oracle", "_ref", "matrix", "bcsstk02.tri")        # 66x66, flops/nnz(L) = 44 -> supernodal under CHOLMOD_AUTO
pytestmark = pytest.mark.skipif(not (os.path.exists(DEMO) and os.path.exists(MAT)), reason="reference demo not built (make -C oracle demo)")


def run_demo(preload=None):
    import tempfile
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = preload
        env["SSB200_VERBOSE"] = "1"
    with open(MAT) as f:
        r = subprocess.run([DEMO], stdin=f, capture_output=True, text=True, env=env, timeout=300,
                           cwd=tempfile.gettempdir())          # the demo appends to ./timelog.m
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"residual \(\|Ax-b\|/\(\|A\|\|x\|\+\|b\|\)\):\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)", r.stdout)
    assert m, r.stdout[-2000:]
    return r, float(m.group(1)), float(m.group(2))


def test_reference_demo_cpu_supernodal_path():
    r, res0, res1 = run_demo()
    assert "supernodal" in r.stdout.lower() or "super" in r.stdout.lower()
    assert res0 < 1e-12 and res1 < 1e-12
    assert "suitesparse_b200" not in r.stderr


@pytest.mark.gpu
def test_reference_demo_with_ld_preload():
    r, res0, res1 = run_demo(preload=B200_LIB)
    assert "[suitesparse_b200] cholmod_l_super_numeric" in r.stderr           # the factorization ran through our symbol
    assert "[suitesparse_b200] cholmod_l_super_lsolve" in r.stderr and "cholmod_l_super_ltsolve" in r.stderr
    assert res0 < 1e-10 and res1 < 1e-10
    r_cpu, c0, c1 = run_demo()
    assert abs(res0 - c0) < 1e-12
