"""Builds btl_bloomfilter_b200/libbtlbf_cuda.so (the sm_100a kernels + the C ABI of include/btlbf.h)
in-tree with nvcc.  nvcc cross-compiles without a GPU, so this runs in the build container too."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbtlbf_cuda.so")
SOURCES = ["kernels.cu", "capi.cu", "sort_bin_build_256.cu", "sort_bin_build_512.cu", "sort_bin_query_256.cu", "sort_bin_query_512.cu", "ingest.cu", "pack.cu", "apply2.cu", "seq_commit_cbf.cu", "seq_commit_bfchk.cu", "seq_ops_bloom.cu", "bin_legacy.cu"]
HEADERS = ["kernels.cuh", "tile_core.cuh", "sort_bin.cuh", "seq_kernel.cuh", "nthash_dev.cuh", "host_params.hpp", os.path.join("..", "..", "include", "btlbf.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libbtlbf_cuda.so")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    newest_header = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        # objects are reused when neither their source nor any header is newer (force rebuilds everything)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(
                newest_header, os.path.getmtime(os.path.join(CSRC, src))):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode(errors="replace")))
    cmd = [nvcc, "-shared", "-o", LIB + ".tmp"] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    os.replace(LIB + ".tmp", LIB)
    return LIB
