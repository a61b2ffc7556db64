"""Builds the sm_100a shared library in-tree (cbench_basic_b200/_lib/libbasic_b200.so) with nvcc.

The .so is git-ignored (built artefact) but NOT gpurun-ignored, so a library cross-compiled in the build
container travels to the GPU box with the snapshot.  nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libbasic_b200.so")
SOURCES = ["capi.cu", "tables.cu", "rans_compat.cu", "rans_lanes.cu", "rans_pair.cu", "gauss.cu", "ctx.cu", "ctx_scan.cu", "ctx_tc.cu", "tans.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-cudart", "static"] + \
             [f"-D{d}" for d in os.environ.get("BASIC_NVCC_DEFS", "").split(",") if d]   # debug builds (kernel timing switches)


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


STAMP = os.path.join(OUT_DIR, "libbasic_b200.sources.sha256")


def source_hash():
    """sha256 over the kernel sources, the header and the flags: what the library was built from (mtimes do not survive a
    checkout or the snapshot copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    files = sorted(os.path.join(SRC, f) for f in os.listdir(SRC)) + [os.path.join(HERE, "..", "include", "basic_b200.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fd:
            h.update(fd.read())
    return h.hexdigest()


def stale():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fd:
        return fd.read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for f in SOURCES:  # compile translation units in parallel
        obj = os.path.join(OUT_DIR, f.replace(".cu", ".o"))
        cmd = [_nvcc()] + [x for x in NVCC_FLAGS if x != "--shared"] + ["-Xptxas", "-v"] * int(verbose) + \
              ["-c", os.path.join(SRC, f), "-o", obj]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for f, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {f}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([_nvcc(), "--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", LIB] + objs + ["-Xlinker", "--export-dynamic"])
    with open(STAMP, "w") as fd:
        fd.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
