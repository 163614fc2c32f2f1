"""Builds picnic_b200/libpicnic_gpu.so (the C-ABI library) with nvcc for sm_100a.

In-tree build: the .so lands next to this file so that it travels with the
repository snapshot to the GPU box.  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpicnic_gpu.so")
SOURCES = ["pgpu_api.cu", "pgpu_push.cu", "pgpu_advance_cc1.cu", "pgpu_advance_cc1_1d.cu", "pgpu_bins.cu", "pgpu_collide.cu", "pgpu_exchange.cu", "pgpu_halo_p2p.cu", "pgpu_massmatrix.cu", "pgpu_suborbit.cu"]
# per-file flags: the mass-matrix deposit keeps every per-particle product an IEEE product (no contraction)
EXTRA_FLAGS = {"pgpu_massmatrix.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "picnic_gpu.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, tag="", defines=()):
    obj = os.path.join(CSRC, src.replace(".cu", tag + ".o"))
    cmd = ["nvcc"] + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + list(defines) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build(force=False, verbose=False):
    if not force and not _needs_build():
        return OUT
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = ["nvcc", "-shared", "-o", OUT] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("built", OUT)
    return OUT


def build_variant(tag, defines):
    """A/B experiments: the same sources with extra -D flags into libpicnic_gpu_<tag>.so, picked up by
    capi.load() when PGPU_LIB=<tag> (kernel tuning only; the product library is libpicnic_gpu.so)."""
    out = os.path.join(HERE, "libpicnic_gpu_%s.so" % tag)
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda f: _compile(f, "_" + tag, defines), SOURCES))
    r = subprocess.run(["nvcc", "-shared", "-o", out] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    print("built", out)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        build_variant(sys.argv[i + 1], sys.argv[i + 2:])
    else:
        build(force="--force" in sys.argv, verbose=True)
