"""Builds libmcsolver_b200.so in-tree with nvcc for sm_100a (no PyTorch).  The NVRTC-specialised modules the library compiles at
run time are cached as cubins under mcsolver_b200/build/jitcache (pre-filled by __graft_entry__.build()).

    python -m mcsolver_b200.build [--force] [--verbose]

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcsolver_b200.so")
SOURCES = ["engine.cu", "structured.cu", "pt.cu"]
NVCC_FLAGS = ["-std=c++20", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newest_source():
    t = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    t = max(t, os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "mcsolver_b200.h")))
    return t


def build(force=False, verbose=False):
    newest = _newest_source()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    headers_t = max([os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if not f.endswith(".cu")] +
                    [os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "mcsolver_b200.h"))])
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    logs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(sp), headers_t):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        logs.append("== %s ==\n%s" % (src, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
    for log in logs:
        name = log.split("==")[1].strip()
        with open(os.path.join(HERE, "build", "ptxas_%s.log" % name.replace(".cu", "")), "w") as f:
            f.write(log)
    if verbose:
        print("\n".join(logs))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
