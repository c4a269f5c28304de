"""Build libmgnns_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

Usage: python -m mgnns_b200.csrc.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["gemm_ffma.cu", "spmm.cu", "text_gcn.cu", "attention.cu", "small_ops.cu", "tc_gemm.cu", "tc_linear.cu", "lstm.cu", "gcn_fused.cu", "pmi_sparse.cu", "count_ops.cu", "attention_tc.cu", "spmm_hub.cu", "optim.cu", "p2p_allreduce.cu", "text_bank_ops.cu"]
HEADERS = ["common.cuh", "tc_common.cuh", os.path.join("..", "..", "include", "mgnns_b200.h")]
LIB = os.path.join(HERE, "libmgnns_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _optional_sources():
    extra = []
    for name in sorted(os.listdir(HERE)):
        if name.endswith(".cu") and name not in SOURCES:
            extra.append(name)
    return extra


def _digest(srcs):
    h = hashlib.sha256()
    for rel in srcs + HEADERS:
        with open(os.path.join(HERE, rel), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    srcs = SOURCES + _optional_sources()
    digest = _digest(srcs)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in srcs:
        obj = os.path.join(HERE, src[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(HERE, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (src, out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log) + "\n")
    if failed:
        raise RuntimeError("nvcc failed; see mgnns_b200/csrc/build.log")
    cmd = [nvcc, "-shared", "-o", LIB] + objs  # static cudart (nvcc default); no libcuda link-time dependency
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
