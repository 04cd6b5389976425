"""In-tree build of libatvs.so (nvcc, sm_100a only).  No JIT cache: the .so lives next to
the sources so that it travels with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libatvs.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
          "-Xcompiler", "-Wno-unused-function"]
# per-file extra flags: the geometry kernels must not contract a*b+c into an FMA (bit parity
# of the sample coordinates with the fp32 oracle)
SOURCES = {
    "api.cu": [],
    "geom.cu": ["--fmad=false"],
    "net_fp32.cu": [],
    "conv_tc.cu": [],
    "conv_ring.cu": [],
    "conv_ring_s2.cu": [],
    "conv_deconv.cu": [],
    "conv_deconv_ring.cu": [],
    "conv_attn_ring.cu": [],
    "fem2d.cu": [],
    "fusion.cu": ["--fmad=false"],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "atvs.h"))
    headers.append(os.path.abspath(__file__))
    nvcc = _nvcc()
    objs, jobs = [], []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            jobs.append(cmd)
    if jobs:     # one nvcc per translation unit, side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            for r in ex.map(lambda c: subprocess.run(c, check=False), jobs):
                if r.returncode != 0:
                    raise subprocess.CalledProcessError(r.returncode, r.args)
    if force or _stale(OUT, objs):
        cmd = [nvcc] + ARCH + ["--shared", "-o", OUT] + objs
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
