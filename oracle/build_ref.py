#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- builds oracle/_ref/: the UNMODIFIED reference OpenBLAS GEMM path,
compiled with gcc straight from the sources where they lie under /root/reference (nothing is
copied into this repository, the reference's own build system is not run).

For each CPU target in oracle/ref_recipe/*.recipe (generic, haswell, skylakex,
sapphirerapids) the ~246 objects the GEMM entry points need -- interface/gemm.c,
driver/level3/gemm.c (level3.c, level3_thread.c), the target's kernel/ files, and the
runtime pieces under driver/others/ -- are compiled with the flags in the recipe against the
target's config.h (oracle/ref_config/<target>/) and linked into

    oracle/_ref/<target>/libopenblas_ref.so

plus, against the same sources, the reference's own harnesses:

    oracle/_ref/ctest/x{s,d,c,z}cblat3      ctest level-3 drivers linked against OUR library
    oracle/_ref/bench/{s,d,c,z,sb}gemm.b200  benchmark/gemm.c linked against OUR library
    oracle/_ref/bench/{s,d,c,z,sb}gemm.<target>  the same linked against the reference

oracle/ref_loader.py picks the best target the host CPU supports at run time (the GPU box's
CPU differs from the authoring container's).  Outputs are git-ignored but travel with gpurun.
"""
import concurrent.futures as cf
import os, shlex, shutil, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OPENBLAS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CC = os.environ.get("CC", "gcc")
TARGETS = ["generic", "haswell", "skylakex", "sapphirerapids"]


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(shlex.quote(c) for c in cmd) + "\n" + r.stdout + "\n")
        raise SystemExit(f"build_ref: command failed ({r.returncode})")
    return r.stdout


def read_recipe(target):
    common, arch, rows = [], [], []
    for line in open(os.path.join(HERE, "ref_recipe", f"{target}.recipe")):
        line = line.rstrip("\n")
        if not line or line.startswith("#"):
            continue
        if line.startswith("COMMON "):
            common = shlex.split(line[7:])
        elif line.startswith("ARCH "):
            arch = shlex.split(line[5:])
        else:
            obj, rel, extra = (line.split("\t") + ["", ""])[:3]
            rows.append((obj, rel, shlex.split(extra)))
    return common, arch, rows


def name_flags(obj):
    n = obj[:-2]
    return ["-UASMNAME", "-UASMFNAME", "-UNAME", "-UCNAME", "-UCHAR_NAME", "-UCHAR_CNAME",
            f"-DASMNAME={n}", f"-DASMFNAME={n}_", f"-DNAME={n}_", f"-DCNAME={n}",
            f'-DCHAR_NAME="{n}_"', f'-DCHAR_CNAME="{n}"', "-DNO_AFFINITY"]


def base_flags(target, common, arch):
    cfg = os.path.join(HERE, "ref_config", target)
    return common + ['-DVERSION="0.3.28.dev"'] + arch + [f"-I{cfg}", f"-I{REF}"]


def build_target(target, jobs):
    common, arch, rows = read_recipe(target)
    odir = os.path.join(OUT, target, "obj")
    os.makedirs(odir, exist_ok=True)
    lib = os.path.join(OUT, target, "libopenblas_ref.so")
    base = base_flags(target, common, arch)

    def compile_one(row):
        obj, rel, extra = row
        src = os.path.join(REF, rel)
        dst = os.path.join(odir, obj)
        if os.path.exists(dst) and os.path.getmtime(dst) > os.path.getmtime(src):
            return dst
        # headers are found relative to the source's own directory first, as in the reference
        run([CC] + base + name_flags(obj) + [f"-I{os.path.dirname(src)}"] + extra + ["-w", "-c", src, "-o", dst])
        return dst

    with cf.ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(compile_one, rows))
    run([CC, "-shared", "-o", lib] + objs + ["-lm", "-lpthread"])
    return lib, base


CTEST_STUBS = r'''
/* TEST INFRASTRUCTURE: nothing is stubbed any more -- the library under test implements every
 * level-3 routine the ctest drivers reference.  (File kept so the link line stays the same.) */
#include <stdio.h>
#include <stdlib.h>
#define STUB(n) void n(void) { fprintf(stderr, "ctest stub " #n " called\n"); abort(); }
'''
OTHER_L3 = []
HERM_L3 = []


def build_ctest(lib_dir, base):
    """ctest/c_?blat3c.c (f2c'd driver) + c_?blas3.c + c_?3chke.c + auxiliary.c + c_xerbla.c +
    constant.c, linked FIRST against libopenblas_b200.so (SURVEY section 4 drop-in recipe)."""
    cdir = os.path.join(OUT, "ctest")
    os.makedirs(cdir, exist_ok=True)
    src = os.path.join(REF, "ctest")
    flags = base + ["-DADD_", "-DCBLAS", "-w", f"-I{src}"]
    for p in "sdcz":
        stubs = CTEST_STUBS + "".join(f"STUB(cblas_{p}{r})\n" for r in OTHER_L3)
        if p in "cz":
            stubs += "".join(f"STUB(cblas_{p}{r})\n" for r in HERM_L3)
        stub_c = os.path.join(cdir, f"stubs_{p}.c")
        open(stub_c, "w").write(stubs)
        objs = []
        for f in [f"c_{p}blat3c.c", f"c_{p}blas3.c", f"c_{p}3chke.c", "auxiliary.c", "c_xerbla.c", "constant.c"]:
            o = os.path.join(cdir, f"{p}_{f[:-2]}.o")
            extra = ["-DDOUBLE"] if p in "dz" else []
            extra += ["-DCOMPLEX"] if p in "cz" else []
            run([CC] + flags + extra + ["-c", os.path.join(src, f), "-o", o])
            objs.append(o)
        # the driver's input file, unchanged, next to the binary (oracle/_ref/ is not part of the history)
        shutil.copyfile(os.path.join(src, f"{p}in3"), os.path.join(cdir, f"{p}in3"))
        exe = os.path.join(cdir, f"x{p}cblat3")
        run([CC, "-o", exe] + objs + [stub_c, f"-L{lib_dir}", "-lopenblas_b200",
                                       f"-Wl,-rpath,$ORIGIN/../../../openblas_b200/lib", "-lm", "-lpthread"])
    # the GEMM3M flavour of the complex drivers (ctest/Makefile:56,64,168-176,312-314,340-342): c_?blat3c_3m.c +
    # c_?blas3_3m.c + c_?3chke_3m.c, input files ?in3_3m
    for p in "cz":
        objs = []
        for f in [f"c_{p}blat3c_3m.c", f"c_{p}blas3_3m.c", f"c_{p}3chke_3m.c", "auxiliary.c", "c_xerbla.c", "constant.c"]:
            o = os.path.join(cdir, f"{p}3m_{f[:-2]}.o")
            extra = ["-DDOUBLE"] if p == "z" else []
            extra += ["-DCOMPLEX"]
            run([CC] + flags + extra + ["-c", os.path.join(src, f), "-o", o])
            objs.append(o)
        shutil.copyfile(os.path.join(src, f"{p}in3_3m"), os.path.join(cdir, f"{p}in3_3m"))
        exe = os.path.join(cdir, f"x{p}cblat3_3m")
        run([CC, "-o", exe] + objs + [f"-L{lib_dir}", "-lopenblas_b200", f"-Wl,-rpath,$ORIGIN/../../../openblas_b200/lib", "-lm", "-lpthread"])
    # the SBGEMM-vs-SGEMM comparison program of the reference (test/compare_sgemm_sbgemm.c).  Its second half
    # compares SBGEMV with SGEMV; SGEMV is not on the GEMM path and not in this library, so the harness gets a
    # plain netlib-semantics sgemv_ of its own (the program only uses it as the fp32 yardstick, tolerance 1.0).
    exe = os.path.join(cdir, "test_sbgemm")
    gemv_stub = os.path.join(cdir, "stubs_gemv.c")
    open(gemv_stub, "w").write(
        "/* TEST INFRASTRUCTURE: fp32 yardstick for the SBGEMV half of compare_sgemm_sbgemm.c (reference/sgemvf.f semantics,\n"
        " * positive increments only, as the program calls it).  sbgemv_ itself comes from libopenblas_b200.so. */\n"
        "void sgemv_(char *trans, int *m, int *n, float *alpha, float *a, int *lda, float *x, int *incx, float *beta, float *y, int *incy) {\n"
        "  int t = (*trans == 'T' || *trans == 't' || *trans == 'C' || *trans == 'c');\n"
        "  int leny = t ? *n : *m, lenx = t ? *m : *n;\n"
        "  for (int i = 0; i < leny; i++) {\n"
        "    float acc = 0.f;\n"
        "    for (int j = 0; j < lenx; j++) acc += (t ? a[j + (long)i * *lda] : a[i + (long)j * *lda]) * x[(long)j * *incx];\n"
        "    y[(long)i * *incy] = *alpha * acc + (*beta == 0.f ? 0.f : *beta * y[(long)i * *incy]);\n"
        "  }\n"
        "}\n")
    run([CC] + base + ["-DBFLOAT16", "-w", os.path.join(REF, "test", "compare_sgemm_sbgemm.c"), gemv_stub, "-o", exe,
                       f"-L{lib_dir}", "-lopenblas_b200", f"-Wl,-rpath,$ORIGIN/../../../openblas_b200/lib",
                       "-lm", "-lpthread"])


def build_bench(lib_dir, base_by_target):
    """benchmark/gemm.c, the reference's own timing harness (config 1 of BASELINE.json)."""
    bdir = os.path.join(OUT, "bench")
    os.makedirs(bdir, exist_ok=True)
    src = os.path.join(REF, "benchmark", "gemm.c")
    prec = {"s": [], "d": ["-DDOUBLE"], "c": ["-DCOMPLEX"], "z": ["-DCOMPLEX", "-DDOUBLE"], "sb": ["-DHALF"]}
    for p, defs in prec.items():
        any_base = next(iter(base_by_target.values()))
        run([CC] + any_base + defs + ["-w", src, "-o", os.path.join(bdir, f"{p}gemm.b200"),
                                      f"-L{lib_dir}", "-lopenblas_b200",
                                      f"-Wl,-rpath,$ORIGIN/../../../openblas_b200/lib", "-lm", "-lpthread"])
        for t, base in base_by_target.items():
            run([CC] + base + defs + ["-w", src, "-o", os.path.join(bdir, f"{p}gemm.{t}"),
                                      os.path.join(OUT, t, "libopenblas_ref.so"),
                                      f"-Wl,-rpath,$ORIGIN/../{t}", "-lm", "-lpthread"])


def main():
    if not os.path.isdir(REF):
        print(f"build_ref: {REF} not present; keeping whatever is already in {OUT}")
        return 0
    jobs = os.cpu_count() or 4
    bases = {}
    for t in TARGETS:
        lib, base = build_target(t, jobs)
        bases[t] = base
        print("built", os.path.relpath(lib, HERE))
    lib_dir = os.path.join(os.path.dirname(HERE), "openblas_b200", "lib")
    if os.path.exists(os.path.join(lib_dir, "libopenblas_b200.so")):
        build_ctest(lib_dir, bases["generic"])
        build_bench(lib_dir, bases)
        print("built ctest + benchmark harnesses against libopenblas_b200.so")
    else:
        print("build_ref: libopenblas_b200.so not built yet; skipped the drop-in harnesses")
    return 0


if __name__ == "__main__":
    sys.exit(main())
