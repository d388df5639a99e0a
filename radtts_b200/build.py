"""In-tree build of the C-ABI CUDA library (sm_100a only).

    python -m radtts_b200.build [--force] [--verbose]

Every csrc/*.cu is compiled with `nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` into
radtts_b200/lib/libradtts_b200.so.  nvcc cross-compiles without a GPU; the .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(LIBDIR, "libradtts_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-I", INCLUDE, "-I", CSRC,
] + os.environ.get("RADTTS_NVCC_EXTRA", "").split()      # e.g. RADTTS_NVCC_EXTRA=-DRB_LSTM_TIMELINE=1 (diagnostic builds)


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; radtts_b200 has no CPU fallback and cannot be built without CUDA")


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return sorted(hs)


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    srcs, hdrs = sources(), headers()
    hdr_digest = _digest(hdrs)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o")
        stamp = o + ".sha"
        want = _digest([s]) + hdr_digest
        objs.append(o)
        have = open(stamp).read() if os.path.exists(stamp) and os.path.exists(o) else ""
        if force or have != want:
            jobs.append((s, o, stamp, want))

    def compile_one(job):
        s, o, stamp, want = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        with open(stamp, "w") as f:
            f.write(want)
        return s, r.stderr

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s, err in ex.map(compile_one, jobs):
                if verbose:
                    print("== %s\n%s" % (os.path.basename(s), err))
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
