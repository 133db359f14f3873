"""Builds libtic_b200.so (and the GPU self-test binary) from csrc/*.cu with nvcc for sm_100a, in-tree under _build/.

    python socialmedia-textimage-classification-auxlosses_b200/build.py [--force]

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libtic_b200.so")
SELFTEST = os.path.join(OUT, "selftest")
LIB_SOURCES = ["gemm.cu", "itc.cu", "heads.cu", "itm.cu", "fusion.cu", "ce.cu", "peer.cu", "attn_mma.cu", "eval.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default"]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    stale = force or not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest_source_mtime()
    if stale:
        objs = []
        procs = []
        for src in LIB_SOURCES:
            obj = os.path.join(OUT, src.replace(".cu", ".o"))
            objs.append(obj)
            procs.append((src, subprocess.Popen([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj],
                                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for src, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
            if verbose and out.strip():
                print(out)
        # the SHARED CUDA runtime (libcudart.so.12: torch's copy when torch is already loaded, else the toolkit's through the
        # rpath): the library carries no private runtime instance and none of the static runtime's symbol strings
        _run([nvcc, "-shared", "-cudart", "shared", "-o", LIB] + objs +
             ["-Xlinker", "-rpath,/usr/local/cuda/lib64", "-ldl", "-lrt", "-lpthread"])
        _run([nvcc] + NVCC_FLAGS + [os.path.join(CSRC, "selftest.cu"), "-o", SELFTEST, "-L" + OUT, "-ltic_b200",
                                    "-cudart", "shared", "-Xlinker", "-rpath," + "$ORIGIN", "-Xlinker", "-rpath,/usr/local/cuda/lib64"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
