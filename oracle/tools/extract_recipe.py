#!/usr/bin/env python3
"""One-off tool (TEST INFRASTRUCTURE): derive oracle/ref_recipe/<target>.recipe from the log of
a scratch build of the reference.

The reference's own build system is NOT part of this repository's build: oracle/build_ref.py
compiles the reference's GEMM-path sources directly with gcc, where they lie under
/root/reference, from the per-object table this tool wrote.  The table was obtained once by
running the reference's make in a scratch copy under /tmp (`make TARGET=<T> NO_LAPACK=1
NOFORTRAN=1 BUILD_BFLOAT16=1 NUM_THREADS=64 libs shared > <T>.log`), linking the GEMM entry
points against the resulting archive with a link map to learn which archive members the path
pulls in (246 objects from ~40 source files), and keeping, for exactly those objects, the
source file and the per-object -D flags.  Usage:

    extract_recipe.py <target> <make.log> <link.map> <scratch tree> > ref_recipe/<target>.recipe
"""
import os, re, shlex, sys

target, log, linkmap, tree = sys.argv[1:5]
members = sorted(set(re.findall(r"libopenblas[^(\s]*\.a\(([^)]+)\)", open(linkmap).read())))
want = set(members)

# flags every command shares: everything up to and including -DVERSION=...; then arch flags
rows, common, arch = {}, None, None
for line in open(log, errors="replace"):
    if " -o " not in line or " -c" not in line or "gcc" not in line.split(" ", 1)[0]:
        continue
    try:
        tok = shlex.split(line)
    except ValueError:
        continue
    if "-o" not in tok:
        continue
    obj = tok[tok.index("-o") + 1]
    if obj not in want or obj in rows:
        continue
    srcs = [t for t in tok[1:] if re.search(r"\.(c|S)$", t) and not t.startswith("-")]
    if len(srcs) != 1:
        continue
    src = srcs[0]
    depth2 = "-I../.." in tok
    if src.startswith("../kernel/"):
        rel = os.path.normpath(os.path.join("kernel", src[len("../kernel/"):]))
    elif depth2:
        rel = next(p for p in (f"driver/level3/{src}", f"driver/others/{src}")
                   if os.path.exists(os.path.join(tree, p)))
    else:   # -I.. : interface/ or a kernel/ sub-directory (kernel/Makefile.L3 names generic/ and $(ARCH)/ sources relatively)
        rel = next((p for p in (f"interface/{src}", f"kernel/{src}") if os.path.exists(os.path.join(tree, p))), f"interface/{src}")
    assert os.path.exists(os.path.join(tree, rel)), (obj, rel)
    i_ver = next(i for i, t in enumerate(tok) if t.startswith("-DVERSION="))
    i_un = tok.index("-UASMNAME")
    i_inc = next(i for i, t in enumerate(tok) if t in ("-I..", "-I../.."))
    if common is None:
        common = [t for t in tok[1:i_ver] if t != "-c"]
        arch = tok[i_ver + 1:i_un]
    pre = [t for t in tok[1:i_ver] if t not in common and t != "-c"]       # e.g. -DCBLAS
    extra = [t for t in tok[i_inc + 1:] if t not in ("-c", "-o", obj, src, "-I.", "-I..")]
    rows[obj] = (rel, pre + extra)

missing = want - set(rows)
assert not missing, f"no compile command found for: {sorted(missing)[:8]}"
print(f"# reference GEMM-path objects for TARGET={target}; written by oracle/tools/extract_recipe.py")
print("COMMON " + " ".join(common))
print("ARCH " + " ".join(arch))
for obj in sorted(rows):
    rel, extra = rows[obj]
    print(f"{obj}\t{rel}\t{' '.join(extra)}")
