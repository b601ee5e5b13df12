"""Builds libdegnorm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SRC = [os.path.join(CSRC, f) for f in ("abi.cu", "nmfoa_tiled.cu", "nmfoa_small_p4.cu", "nmfoa_small_p8.cu",
                                       "nmfoa_small_p12.cu", "nmfoa_mid_w8.cu", "nmfoa_mid_w4.cu", "nmfoa_mid_ws.cu", "nmfoa_wide.cu", "probes.cu")]
HDR = [os.path.join(CSRC, f) for f in ("common.cuh", "launch.h", "tma.cuh", "nmfoa_small.cuh", "nmfoa_mid.cuh")]
# tuning aid: DEGNORM_B200_VARIANT=name with DEGNORM_B200_NVCC_FLAGS="-DMID_FEAT=3" builds libdegnorm_b200.name.so
# beside the product library (loaded with DEGNORM_B200_LIB=<path>, see _lib.py)
VARIANT = os.environ.get("DEGNORM_B200_VARIANT", "")
OBJ = os.path.join(HERE, "csrc", "_build" + ("_" + VARIANT if VARIANT else ""))
OUT = os.path.join(HERE, "libdegnorm_b200%s.so" % ("." + VARIANT if VARIANT else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include")] + os.environ.get("DEGNORM_B200_NVCC_FLAGS", "").split()


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = SRC + HDR + [os.path.join(ROOT, "include", "degnorm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(args):
    src, obj, verbose = args
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    """One nvcc per translation unit, in parallel; objects whose source and headers are unchanged are re-used."""
    if not force and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in HDR + [os.path.join(ROOT, "include", "degnorm_b200.h")])
    jobs, objs = [], []
    for s in SRC:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            jobs.append((s, o, verbose))
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        for src, rc, log in ex.map(_compile, jobs):
            if verbose or rc != 0:
                sys.stderr.write("== %s\n%s" % (os.path.basename(src), log))
            if rc != 0:
                raise RuntimeError("nvcc failed on %s" % src)
    r = subprocess.run([NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking %s" % OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
