"""Ahead-of-time build of the C-ABI library: nvcc -> lattice_net_b200/liblattice_b200.so (sm_100a only).

No JIT at run time (the reference JIT-compiles every <pos_dim, val_dim> instantiation through
jitify/NVRTC on first use, /root/reference/include/lattice_net/jitify_helper/jitify_helper.cuh:19-38).
"""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "liblattice_b200.so")
SOURCES = ["ln_api.cu", "ln_hash_splat.cu", "ln_neighbours.cu", "ln_slice.cu", "ln_conv.cu", "ln_conv_tc.cu", "ln_norm.cu", "ln_train.cu", "ln_pointnet.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-ftz=true",            # flush-to-zero like the reference's fast-math build; everything else stays IEEE
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.isfile(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, "..", "include", "lattice_b200.h")]
    return any(os.path.getmtime(d) > lib_m for d in deps if os.path.isfile(d))


def build(force=False, verbose=True, extra_flags=()):
    if not force and not needs_build():
        return LIB_PATH
    objs = []
    procs = []
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags)
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[build] {src} FAILED\n{out}\n")
        elif verbose and out.strip():
            sys.stderr.write(f"[build] {src}:\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    if verbose:
        print(f"[build] {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, extra_flags=[a for a in sys.argv[1:] if a.startswith("-X") or a == "-v"])
