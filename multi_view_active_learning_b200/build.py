"""Builds libmval_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

Every csrc/*.cu is compiled to its own object (in parallel, only when it or a header is newer) and the objects are
linked into one shared library; there is no relocatable device code, each translation unit is self-contained."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_obj")
LIB = os.path.join(PKG, "libmval_b200.so")
SOURCES = ["capi.cu", "decode.cu", "peaks.cu", "xe.cu", "mapstream.cu", "triangulate.cu", "refine.cu", "fused.cu", "select.cu",
           "sal.cu", "kcenter.cu", "kcenter_tc.cu", "synth.cu", "pipeline.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; mval_b200 needs the CUDA toolkit to build (there is no CPU fallback)")


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(ROOT, "include", "mval_b200.h")]


def _sources():
    return [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources()] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    jobs, objs = [], []
    for s in _sources():
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            cmd = [nvcc] + FLAGS + (["-Xptxas=-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % cmd[-3])

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1) or 1) as ex:
        list(ex.map(run, jobs))
    subprocess.run([nvcc, "-arch=sm_100a", "-shared", "-o", LIB + ".tmp"] + objs, check=True, cwd=CSRC)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
