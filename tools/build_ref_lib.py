"""Build libcbas_b200 from another git revision into cbas_b200/_ab/ for A/B timing on one GPU box.
usage: python tools/build_ref_lib.py <git-ref|WORKTREE> [nvcc flags]   ->   cbas_b200/_ab/libcbas_b200_<ref>[_flags].so
run with: CBAS_B200_LIB=cbas_b200/_ab/libcbas_b200_<ref>.so python bench.py ..."""
import os, subprocess, sys, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = sys.argv[1]          # a git revision, or WORKTREE for the files as they are now
extra = sys.argv[2:]       # extra nvcc flags, e.g. -DATC_POLY_PAIRS=3 (they also name the output)
out_dir = os.path.join(ROOT, "cbas_b200", "_ab")
os.makedirs(out_dir, exist_ok=True)
tmp = tempfile.mkdtemp(prefix="cbas_ab_")
try:
    if ref == "WORKTREE":
        shutil.copytree(os.path.join(ROOT, "cbas_b200", "csrc"), os.path.join(tmp, "cbas_b200", "csrc"),
                        ignore=shutil.ignore_patterns("build"))
        shutil.copytree(os.path.join(ROOT, "include"), os.path.join(tmp, "include"))
    else:
        tar = subprocess.run(["git", "-C", ROOT, "archive", ref, "cbas_b200/csrc", "include"], check=True, capture_output=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
    srcs = [os.path.join(tmp, "cbas_b200", "csrc", f) for f in ("api.cu", "gemm.cu", "encoder.cu", "head.cu")]
    tag = ref.replace('/', '_') + "".join("_" + f.lstrip("-").replace("=", "") for f in extra)
    out = os.path.join(out_dir, f"libcbas_b200_{tag}.so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O3",
           "--expt-relaxed-constexpr", *extra, "-I", os.path.join(tmp, "include"), "-shared", "-o", out, *srcs, "-cudart", "static"]
    subprocess.run(cmd, check=True)
    print(out)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
