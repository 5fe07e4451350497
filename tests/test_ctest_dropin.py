"""The reference's own acceptance programs, UNMODIFIED, linked against libopenblas_b200.so as a
drop-in (oracle/build_ref.py compiles ctest/c_?blat3c.c, c_?blas3.c, c_?3chke.c, c_xerbla.c from
/root/reference and links them first against our library): 27 783 cblas_?gemm calls per layout
for s/d/z (17 496 for c), the SYMM/HEMM, SYRK/HERK, SYR2K/HER2K, TRMM and TRSM sweeps, and the error-exit
checks of all of them, each judged by the reference's own checkers (DMMCH and friends).
Inputs: the reference's own ?in3 files, every routine enabled (oracle/build_ref.py puts them next
to the binaries under oracle/_ref/ctest/)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTEST = os.path.join(ROOT, "oracle", "_ref", "ctest")


@pytest.mark.parametrize("p", list("sdcz"))
def test_ctest_level3_gemm(p):
    exe = os.path.join(CTEST, f"x{p}cblat3")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ctest not built (needs /root/reference at build time)")
    with open(os.path.join(CTEST, f"{p}in3")) as f:      # the reference's own input file, copied by oracle/build_ref.py
        r = subprocess.run([exe], stdin=f, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    out = r.stdout
    print(out[-3000:])
    assert r.returncode == 0, out[-2000:]
    assert "FATAL" not in out and "FAIL" not in out.replace("FAILURES", ""), out[-2000:]
    import re
    assert len(re.findall(rf"cblas_{p}gemm\s+PASSED THE TESTS OF ERROR-EXITS", out)) == 1, out[-2000:]
    calls = 17496 if p == "c" else 27783
    assert re.search(rf"cblas_{p}gemm\s+PASSED THE COLUMN-MAJOR COMPUTATIONAL TESTS \(\s*{calls} CALLS\)", out), out[-2000:]
    assert re.search(rf"cblas_{p}gemm\s+PASSED THE ROW-MAJOR\s+COMPUTATIONAL TESTS \(\s*{calls} CALLS\)", out), out[-2000:]
    family = ["symm", "syrk", "syr2k", "trmm", "trsm"] + (["hemm", "herk", "her2k"] if p in "cz" else [])
    for r in family:
        assert len(re.findall(rf"cblas_{p}{r}\s+PASSED THE TESTS OF ERROR-EXITS", out)) == 1, (r, out[-3000:])
        assert re.search(rf"cblas_{p}{r}\s+PASSED THE COLUMN-MAJOR COMPUTATIONAL TESTS \(\s*\d+ CALLS\)", out), (r, out[-3000:])
        assert re.search(rf"cblas_{p}{r}\s+PASSED THE ROW-MAJOR\s+COMPUTATIONAL TESTS \(\s*\d+ CALLS\)", out), (r, out[-3000:])


@pytest.mark.parametrize("p", list("cz"))
def test_ctest_level3_gemm3m(p):
    """ctest/Makefile:168-176,210-225: x?cblat3_3m < ?in3_3m, the GEMM3M flavour of the complex drivers
    (c_?blat3c_3m.c, c_?blas3_3m.c, c_?3chke_3m.c), linked against this library alone.  The f2c'd driver the reference
    ships for builds without a Fortran compiler matches routine names in 12 characters and "cblas_?gemm3m" has 13, so
    it reports every routine as NOT TESTED -- against the reference itself as well: the transcript of the same
    binary linked against oracle/_ref/generic is committed (tests/golden/ctest_3m_reference_transcript_?.txt) and
    ours must equal it line for line.  The sweep the Fortran driver would run (c_?blat3_3m.f, ?CHK1) is restated in
    tests/test_gemm_gpu.py::test_gemm3m_over_the_ctest_grid."""
    exe = os.path.join(CTEST, f"x{p}cblat3_3m")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ctest not built (needs /root/reference at build time)")
    with open(os.path.join(CTEST, f"{p}in3_3m")) as f:
        r = subprocess.run([exe], stdin=f, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    out = r.stdout
    print(out[-3000:])
    assert r.returncode == 0, out[-2000:]
    assert "FATAL" not in out and "FAIL" not in out.replace("FAILURES", ""), out[-2000:]
    want = open(os.path.join(ROOT, "tests", "golden", f"ctest_3m_reference_transcript_{p}.txt")).read()
    assert out.splitlines() == want.splitlines()


def test_compare_sgemm_sbgemm():
    """test/compare_sgemm_sbgemm.c, unstubbed: the SBGEMM half against SGEMM, then the SBGEMV half (sbgemv_ of this
    library against an fp32 sgemv_ the harness brings along, and against the program's own bf16 dot products)."""
    exe = os.path.join(CTEST, "test_sbgemm")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ctest not built")
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert "FATAL ERROR" not in r.stdout, r.stdout[-1000:]
    assert r.returncode == 0, r.stdout[-1000:]
