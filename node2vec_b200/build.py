"""Compile node2vec_b200/csrc/*.cu into libn2v_b200.so for sm_100a (nvcc, in-tree)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libn2v_b200.so")
SOURCES = ["abi.cu", "peer_mem.cu", "csr_build.cu", "hash_build.cu", "alias_build.cu", "trim.cu", "index.cu", "walk.cu", "vocab.cu", "sgns.cu", "sgns_shared.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + [os.path.join(CSRC, "n2v_internal.cuh"), os.path.join(CSRC, "alias_core.cuh"),
                        os.path.join(HERE, "..", "include", "n2v_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _nvcc() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def have_nvcc() -> bool:
    return os.path.exists(_nvcc())


def build_library(force: bool = False, extra_flags=(), verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    nvcc = _nvcc()
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libn2v_b200.so")
    # one builder at a time (torchrun starts one process per GPU; all of them may find the .so stale):
    # an exclusive file lock, the build goes to a temporary name and is renamed into place, so no
    # process can ever dlopen a half-written library
    import fcntl
    with open(SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():          # somebody else built it while we waited
                return SO
            tmp = f"{SO}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-o", tmp] + sources()
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            try:
                subprocess.check_call(cmd, cwd=CSRC)
                os.replace(tmp, SO)
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO


EXAMPLE_SRC = os.path.join(HERE, "..", "examples", "c_consumer.c")
EXAMPLE_BIN = os.path.join(HERE, "..", "examples", "_build", "c_consumer")


def build_c_consumer(force: bool = False) -> str:
    """examples/c_consumer.c: a plain-C11 program that drives the hot path through the C ABI alone
    (gcc, not nvcc: the header must be consumable by a C compiler, as a cgo/JNI binding would)."""
    so = build_library()
    if not force and os.path.exists(EXAMPLE_BIN) and os.path.getmtime(EXAMPLE_BIN) >= max(
            os.path.getmtime(EXAMPLE_SRC), os.path.getmtime(so)):
        return EXAMPLE_BIN
    cuda = os.path.dirname(os.path.dirname(_nvcc()))
    os.makedirs(os.path.dirname(EXAMPLE_BIN), exist_ok=True)
    subprocess.check_call([
        "gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O2", EXAMPLE_SRC,
        "-I", os.path.join(HERE, "..", "include"), "-I", os.path.join(cuda, "include"),
        "-L", CSRC, "-ln2v_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
        "-Wl,-rpath,$ORIGIN/../../node2vec_b200/csrc", "-Wl,-rpath," + os.path.join(cuda, "lib64"),
        "-o", EXAMPLE_BIN])
    return EXAMPLE_BIN


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_c_consumer(force="--force" in sys.argv))
