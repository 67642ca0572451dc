"""Build libshortseq_b200.so (sm_100a only) in-tree with nvcc.

    python -m shortseq_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libshortseq_b200.so")
SOURCES = ["ssq_ctx.cu", "ssq_scan.cu", "ssq_pack.cu", "ssq_counter.cu", "ssq_codec.cu", "ssq_host.cu", "ssq_fastq.cu", "ssq_slice.cu", "ssq_comm.cu", "ssq_one.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--expt-relaxed-constexpr"]


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "shortseq_b200.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: development variants (e.g. -DSSQ_LINE_KEYS=8 into variants/libssq_line8.so, loaded with
    SSQ_LIB=...); the product library is always built without defines."""
    lib = LIB if out is None else os.path.join(HERE, "variants", out)
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs, procs = [], []
    bdir = os.path.join(HERE, "build" if out is None else os.path.join("build", out))
    os.makedirs(bdir, exist_ok=True)
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd)))
        objs.append(obj)
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed for {src}")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lcudart", "-ldl"])
    return lib


HOSTEXT_SRC = os.path.join(HERE, "hostext", "_fastbox.c")


def hostext_path():
    import sysconfig
    return os.path.join(HERE, "_fastbox" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_hostext(force=False):
    """Compile the CPython helper module (gather / boxing loops of the drop-in ShortSeqCounter) with gcc."""
    import sysconfig
    out = hostext_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(HOSTEXT_SRC):
        return out
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-Wall", "-I", inc, HOSTEXT_SRC, "-o", out])
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
    if not outs:
        print(build_hostext(force="--force" in sys.argv))
