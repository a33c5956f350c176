"""Builds paropt_b200/libparopt_b200.so (sm_100a only) with nvcc, in-tree."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, os.environ.get("PCU_LIB_NAME", "libparopt_b200.so"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
    "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xptxas", "-v" if os.environ.get("PCU_PTXAS_V") else "-O3",
] + os.environ.get("PCU_EXTRA_FLAGS", "").split()


# per-file nvcc flags (none today)
PER_FILE_FLAGS = {}


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = glob.glob(os.path.join(CSRC, "*")) + [
        os.path.join(HERE, "..", "include", "paropt_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build", os.path.basename(LIB))
    os.makedirs(bdir, exist_ok=True)
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + FLAGS + PER_FILE_FLAGS.get(os.path.basename(src), []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-ldl"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
