"""Build librcbf_b200.so (the sm_100a kernels + C ABI) in-tree with nvcc.  No JIT cache: the .so sits next to this file
so it travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librcbf_b200.so")
SOURCES = ["rcbf_kernels.cu", "rcbf_safe_unicycle.cu", "rcbf_safe2_unicycle.cu", "rcbf_safe_cars.cu", "rcbf_cars2.cu", "rcbf_gp.cu", "rcbf_general.cu", "rcbf_replay.cu"]
HEADERS = ["rcbf_core.cuh", "rcbf_dynamics.cuh", "rcbf_backward.cuh", "rcbf_generic.cuh", "rcbf_safe_kernels.cuh", "rcbf_safe2.cuh", "rcbf_cars2.cuh", "rcbf_f2.cuh", "rcbf_tma.cuh",
           os.path.join("..", "..", "include", "rcbf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
NVCC_FLAGS += os.environ.get("RCBF_NVCC_EXTRA", "").split()  # A/B switches (-DRCBF_S2_...=...) for scripts/gpu_ab.sh


HASH_FILE = LIB + ".srchash"


def _source_hash():
    """Content hash of every source / header / flag: file mtimes do not survive a snapshot copy or a git checkout."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in sorted(SOURCES + HEADERS):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def needs_build():
    if not (os.path.exists(LIB) and os.path.exists(HASH_FILE)):
        return True
    with open(HASH_FILE) as fh:
        return fh.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile the translation units in parallel (nvcc -c), then link them into the shared library."""
    if not (force or needs_build()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose:
            print(out)
        objs.append(obj)
    r = subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a"] + objs +
                       ["-o", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(HASH_FILE, "w") as fh:
        fh.write(_source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
