"""Builds `libultra_rspmm.so` (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The shared library is what travels to the GPU box (it is git-ignored, not gpurun-ignored).
`nvcc` cross-compiles without a GPU, so `build()` also runs in the authoring container.
Staleness is decided by a content hash of the sources (file times do not survive every copy), and concurrent
callers (one process per GPU under torchrun) are serialised with a file lock.
"""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libultra_rspmm.so")
HASH_PATH = LIB_PATH + ".sha256"
SOURCES = ["rspmm_api.cu", "rspmm_index.cu", "rspmm_kernels.cu", "rspmm_staged.cu", "rspmm_narrow.cu", "rspmm_extend.cu", "probe.cu", "layer_epilogue.cu", "layer_linear.cu", "layer_linear_tc.cu", "layer_gemm_tc.cu"]
HEADERS = [os.path.join(CSRC, "rspmm_common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(ROOT, "include", "ultra_rspmm.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc():
    for candidate in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if candidate and os.path.exists(candidate):
            return candidate
    raise RuntimeError("nvcc not found: the rspmm CUDA library cannot be built (there is no CPU fallback)")


def _digest(paths):
    digest = hashlib.sha256(" ".join(NVCC_FLAGS[:8]).encode())
    for path in paths:
        with open(path, "rb") as handle:
            digest.update(os.path.basename(path).encode() + b"\0" + handle.read())
    return digest.hexdigest()


def source_hash():
    return _digest([os.path.join(CSRC, s) for s in SOURCES] + HEADERS)


def is_current():
    """True when the library on disk was built from the sources on disk."""
    try:
        with open(HASH_PATH) as handle:
            return os.path.exists(LIB_PATH) and handle.read().strip() == source_hash()
    except OSError:
        return False


def build(force=False, verbose=False):
    """Compile each .cu whose content (or a header) changed and link the shared library.  Returns its path."""
    if not force and is_current():
        return LIB_PATH
    build_dir = os.path.join(HERE, "_build")
    os.makedirs(build_dir, exist_ok=True)
    with open(os.path.join(build_dir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():      # another process built it while this one waited
                return LIB_PATH
            nvcc = _nvcc()
            objects, jobs = [], []
            for source in SOURCES:
                src = os.path.join(CSRC, source)
                obj = os.path.join(build_dir, source.replace(".cu", ".o"))
                stamp = obj + ".sha256"
                wanted = _digest([src] + HEADERS)
                objects.append(obj)
                try:
                    with open(stamp) as handle:
                        fresh = os.path.exists(obj) and handle.read().strip() == wanted
                except OSError:
                    fresh = False
                if force or not fresh:
                    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
                    jobs.append((source, stamp, wanted,
                                 subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            failed = False
            for source, stamp, wanted, job in jobs:
                out, _ = job.communicate()
                if job.returncode != 0:
                    failed = True
                    sys.stderr.write("nvcc failed on %s:\n%s\n" % (source, out))
                else:
                    with open(stamp, "w") as handle:
                        handle.write(wanted)
                    if verbose:
                        sys.stderr.write(out)
            if failed:
                raise RuntimeError("nvcc failed")
            temporary = LIB_PATH + ".tmp.%d" % os.getpid()
            subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", temporary] +
                                  objects + ["-lcudart"])
            os.replace(temporary, LIB_PATH)
            with open(HASH_PATH, "w") as handle:
                handle.write(source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
