"""Build libssdsb200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python -m ssds_pytorch_b200.build [--force] [--verbose]

The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libssdsb200.so")
STAMP = os.path.join(HERE, ".libssdsb200.stamp")
LOCK = os.path.join(HERE, ".libssdsb200.lock")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-I", INCLUDE]
# exact-arithmetic TUs: no FMA contraction so fp32 results match the reference's op-by-op order
EXACT = {"nms.cu", "decode.cu", "decode_large.cu", "anchors.cu", "match.cu", "loss.cu", "loss2.cu", "loss_step.cu"}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def digest():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for f in sorted(os.listdir(root)):
            p = os.path.join(root, f)
            if os.path.isfile(p):
                h.update(f.encode())
                h.update(open(p, "rb").read())
    h.update(open(__file__, "rb").read())
    return h.hexdigest()


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _fresh(d):
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == d


def build(force=False, verbose=False):
    d = digest()
    if not force and _fresh(d):
        return LIB
    # one builder at a time: under torchrun every rank imports the package at once; the others wait here
    # and then find the library up to date
    with open(LOCK, "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        if not force and _fresh(d):
            return LIB
        return _build_locked(d, verbose)


def _build_locked(d, verbose):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc()] + ARCH + COMMON + (["-fmad=false"] if src in EXACT else [])
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB + f".tmp.{os.getpid()}"
    subprocess.check_call([nvcc()] + ARCH + ["-shared", "-o", tmp] + objs)
    os.replace(tmp, LIB)                     # atomic: a concurrent loader never sees a half-written file
    open(STAMP, "w").write(d)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
