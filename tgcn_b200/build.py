"""Build recipe for libtgcn_b200.so (sm_100a only, in-tree so the .so travels to the GPU box)."""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtgcn_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint():
    h = hashlib.sha256()
    root = os.path.dirname(HERE)
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(root, "include", "tgcn_b200.h"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _up_to_date(fp):
    if os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            return fh.read().strip() == fp
    return False


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link the shared library. Returns its path.  Safe to call from
    several processes at once (torchrun ranks): the build runs under an exclusive file lock and the late comers find
    the library up to date."""
    import fcntl
    fp = _fingerprint()
    if not force and _up_to_date(fp):
        return LIB
    lock_path = os.path.join(HERE, "csrc", ".build_lock")
    with open(lock_path, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(fp):
                return LIB
            return _build_locked(fp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(fp, verbose):
    nvcc = nvcc_path()
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libtgcn_b200.so (and no prebuilt library is present)")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = list(FLAGS)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + ARCH + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    tmp = LIB + ".tmp.%d" % os.getpid()
    link = [nvcc] + ARCH + ["-shared", "-o", tmp] + objs + ["-cudart", "static"]
    subprocess.check_call(link)
    os.replace(tmp, LIB)               # atomic: a concurrent loader never maps a half-written library
    with open(STAMP, "w") as fh:
        fh.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
