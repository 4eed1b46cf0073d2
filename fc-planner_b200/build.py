"""Builds libfcvm.so (CUDA kernels + host manager + C ABI) for sm_100a with nvcc, in-tree.

    python fc-planner_b200/build.py [--force]

nvcc cross-compiles without a GPU.  -fmad=false keeps every FP64 expression un-contracted on the
device (the reference is an SSE2 build without FMA); -ffp-contract=off does the same for the host
side of the library.  -lineinfo lets ncu's source page map back to these files.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build", "libfcvm.so")
SOURCES = ["fcvm_kernels.cu", "fcvm_host.cu", "fcvm_cov.cu", "fcvm_map.cu", "fcvm_plan.cu", "fcvm_pose.cu", "fcvm_comm.cu"]
DEPS = SOURCES + ["fcvm_device.cuh", "fcvm_kernels.h", "fcvm_geom.h", "fcvm_hostutil.h", "fcvm_ddmath.h", "fcvm_ddmath_tables.h", "fcvm_comm.h", os.path.join("..", "..", "include", "fcvm.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-shared", "-Xcompiler", "-fPIC,-O3,-ffp-contract=off,-pthread", "-cudart", "static",
]


def kernel_build_id():
    """Hash of the sources k_eval is compiled from: stamps ncu captures (profiles/traffic.json) so that bench.py can tell
    whether the quoted traffic figures belong to the kernel it is timing."""
    import hashlib
    h = hashlib.sha1()
    for name in ("fcvm_kernels.cu", "fcvm_device.cuh", "fcvm_kernels.h"):
        h.update(open(os.path.join(CSRC, name), "rb").read())
    return h.hexdigest()[:12]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build() and not os.environ.get("FCVM_OUT"):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # tuning experiments: FCVM_DEFS="EVAL_UNROLL=2,EVAL_KEEP=16" FCVM_OUT=/path/variant.so (load with FCVM_LIB=...)
    defs = ["-D" + d for d in os.environ.get("FCVM_DEFS", "").split(",") if d]
    out = os.environ.get("FCVM_OUT", OUT)
    tmp = out + ".tmp%d" % os.getpid()     # the library appears atomically (a repo snapshot never sees a half-written file)
    cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout)
    os.replace(tmp, out)
    if verbose:
        print(r.stdout)
    return out


DBG = os.path.join(HERE, "_build", "libfcvm_dbg.so")


def build_debug(force=False):
    """The FCVM_DEBUG_BOUNDS build (every data-dependent index of k_eval checked and counted), used by
    tests/test_gpu_bounds.py in place of compute-sanitizer."""
    if not force and os.path.exists(DBG) and not any(os.path.getmtime(os.path.join(CSRC, d)) > os.path.getmtime(DBG) for d in DEPS):
        return DBG
    old = {k: os.environ.get(k) for k in ("FCVM_DEFS", "FCVM_OUT")}
    os.environ["FCVM_DEFS"], os.environ["FCVM_OUT"] = "FCVM_DEBUG_BOUNDS=1", DBG
    try:
        return build(force=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
