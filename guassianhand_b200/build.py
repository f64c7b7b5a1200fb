"""In-tree build of libghr.so (hand-written CUDA for sm_100a, C ABI in include/ghr.h).

    python -m guassianhand_b200.build [--force]

Plain nvcc, no torch headers: the library has no Python/torch dependency, PyTorch only owns the
device memory and the stream the host code passes in.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libghr.so")
OBJ = os.path.join(HERE, "csrc", "_obj")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "blend.cu", "preprocess_bwd.cu", "attributes.cu", "comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(out=None):
    out = out or OUT
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "ghr.h"))
    return any(os.path.getmtime(d) > t for d in deps)


# Build-time variants (never the shipped library): test / profiling instrumentation selected with -D flags,
# written next to libghr.so as libghr_<name>.so.  tests and tools load one by setting _native.LIB_PATH
# before the first _native.lib() call.
VARIANTS = {
    "exact": ["-DGHR_EXACT_EXP"],        # libdevice expf + IEEE divide in the blend kernels (parity counting test)
    "count": ["-DGHR_COUNT"],            # evaluated / contributing pair counters of the blend kernels (tools/cull_stats.py)
    "timeline": ["-DGHR_TIMELINE"],      # per-CTA start/stop clocks of the blend kernels (tools/blend_timeline.py)
    "bwd2": ["-DGHR_BWD_WARPS=2"],       # A/B: two half-tile CTAs per backward unit
    "bwdilp3": ["-DGHR_BWD_ILP=3"],      # A/B: instances per backward iteration
    "bwdilp4": ["-DGHR_BWD_ILP=4"],
    "bwdocc8": ["-DGHR_BWD_MINCTAS=8"],  # A/B: 64 registers, 8 CTAs per SM
    "bwdocc6": ["-DGHR_BWD_MINCTAS=6"],
    # A/B: forward blend occupancy (registers via min CTAs, shared memory via ring depth) against per-warp ILP
    "fwdilp8occ5": ["-DGHR_FWD_ILP=8", "-DGHR_FWD_MINCTAS=4", "-DGHR_FWD_STAGES=6"],   # the round-2 start point
    "fwdilp8occ8": ["-DGHR_FWD_ILP=8", "-DGHR_FWD_MINCTAS=8", "-DGHR_FWD_STAGES=4"],
    "nohalfq": ["-DGHR_BWD_NO_HALFQ"],    # A/B: backward blend with one survivor queue per warp (8x8 block)
    "nohalfq_count": ["-DGHR_BWD_NO_HALFQ", "-DGHR_COUNT"],
    "pbwd64": ["-DGHR_PBWD_THREADS=64"],  # A/B: preprocess_backward with 64-thread blocks
    "nored": ["-DGHR_NO_RED"],           # experiment: backward blend without its global reductions (wrong results)
}


def variant_path(name: str) -> str:
    return os.path.join(HERE, f"libghr_{name}.so")


def build(force: bool = False, verbose: bool = False, variant: str = None, defines=None) -> str:
    out, objdir, extra = OUT, OBJ, []
    if variant:
        out, objdir = variant_path(variant), os.path.join(OBJ, variant)
        extra = list(VARIANTS[variant] if defines is None else defines)
    if not force and not _stale(out):
        return out
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(one, SOURCES))
    log = "\n".join(r[1] for r in res)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write(log)
    if verbose:
        print(log)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *[r[0] for r in res], "-lcudart"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")]
    for v in names or [None]:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=v))
