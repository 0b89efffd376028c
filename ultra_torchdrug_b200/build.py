"""Builds `libultra_rspmm.so` (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The shared library is what travels to the GPU box (it is git-ignored, not gpurun-ignored).
`nvcc` cross-compiles without a GPU, so `build()` also runs in the authoring container.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libultra_rspmm.so")
SOURCES = ["rspmm_api.cu", "rspmm_index.cu", "rspmm_kernels.cu", "layer_epilogue.cu"]
HEADERS = ["rspmm_common.cuh", os.path.join(ROOT, "include", "ultra_rspmm.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc():
    for candidate in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if candidate and os.path.exists(candidate):
            return candidate
    raise RuntimeError("nvcc not found: the rspmm CUDA library cannot be built (there is no CPU fallback)")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    mtime = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mtime for d in deps)


def build(force=False, verbose=False):
    """Compile each .cu to an object (only when stale) and link the shared library."""
    nvcc = _nvcc()
    build_dir = os.path.join(HERE, "_build")
    os.makedirs(build_dir, exist_ok=True)
    headers = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objects, jobs = [], []
    for source in sources:
        src = os.path.join(CSRC, source)
        obj = os.path.join(build_dir, source.replace(".cu", ".o"))
        objects.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((source, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for source, job in jobs:
        out, _ = job.communicate()
        if job.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (source, out))
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed")
    if force or jobs or _stale(LIB_PATH, objects):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objects + ["-lcudart"])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
