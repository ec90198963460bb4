"""Build libvlpet.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.  No torch involved.

Every ``csrc/*.cu`` is compiled to its own object (in parallel, cached under ``build/`` by a digest of the source, the
shared headers and the flags) and the objects are linked into ``libvlpet.so``.  The library is written to a temporary
file and moved into place under a file lock, so concurrent ranks of a ``torchrun`` launch never load a half-written file
(ranks that lose the race wait on the lock and then find the stamp up to date)."""
import fcntl
import glob
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libvlpet.so")
STAMP = os.path.join(HERE, ".libvlpet.stamp")
LOCK = os.path.join(HERE, ".libvlpet.lock")
CC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
            "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
NVCC_FLAGS = CC_FLAGS + ["-shared"]   # (kept under this name: the stamp covers compile and link flags)


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libvlpet.so cannot be built (there is no CPU fallback)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "vlpet.h")]


def _sha(paths, extra=()):
    h = hashlib.sha256()
    for f in paths:
        with open(f, "rb") as fh:
            h.update(fh.read())
    for e in extra:
        h.update(e.encode())
    return h.hexdigest()


def _digest(defines=()):
    return _sha(sources() + _headers(), [" ".join(NVCC_FLAGS), " ".join(defines)])


def _compile_one(src, defines, log):
    key = _sha([src] + _headers(), [" ".join(CC_FLAGS), " ".join(defines)])[:20]
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + "." + key + ".o")
    if os.path.exists(obj):
        return obj
    for old in glob.glob(os.path.join(OBJ, os.path.basename(src)[:-3] + ".*.o")):
        if not defines:   # variant builds (extra -D) keep the default objects
            os.unlink(old)
    cmd = [_nvcc()] + CC_FLAGS + list(defines) + ["-c", "-o", obj + ".tmp", src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed on " + src)
    os.replace(obj + ".tmp", obj)
    return obj


def build(force=False, verbose=False, defines=(), out=None):
    """Default: libvlpet.so.  ``defines`` + ``out`` build a variant library (developer tools), no stamp."""
    out = out or LIB
    variant = out != LIB
    dig = _digest(defines)
    if not variant and not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with open(LOCK, "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if not variant and not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
                return LIB   # another rank built it while we waited
            log = []
            with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
                objs = list(ex.map(lambda s: _compile_one(s, defines, log), sources()))
            tmp = out + ".tmp.%d" % os.getpid()
            cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
            res = subprocess.run(cmd, capture_output=True, text=True)
            log.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if verbose or res.returncode != 0:
                sys.stderr.write("\n".join(log))
            if res.returncode != 0:
                raise RuntimeError("nvcc failed linking " + out)
            os.replace(tmp, out)
            if not variant:
                with open(os.path.join(OBJ, "build.log"), "w") as fh:
                    fh.write("\n".join(log))
                with open(STAMP, "w") as fh:
                    fh.write(dig)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return out


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
