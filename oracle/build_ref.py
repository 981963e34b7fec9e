#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle) -- build recipe for the reference's own device kernels.

The reference (AIS-Bonn/lattice_net) is CUDA-only and compiles its kernels at run time
through jitify -> NVRTC with `-std=c++11 --use_fast_math`
(/root/reference/include/lattice_net/jitify_helper/jitify_helper.cuh:29) for the
compute capability of the current device (/root/reference/deps/jitify/jitify.hpp:1786-1809).
This script does exactly that compilation ahead of time, reading the two kernel
headers *in place* from /root/reference (never copied into the repo) and writing only
build products to oracle/_ref/ (git-ignored, shipped to the GPU box):

    oracle/_ref/lattice_ref.ptx        PTX, `.target sm_100`; the driver JITs it on the B200
                                       exactly like jitify's cuModuleLoadData path does
    oracle/_ref/lattice_ref.names.json name expression -> lowered (mangled) kernel name

`-default-device` is needed because NVRTC 12.9 rejects the un-annotated HashTableGPU
constructors (HashTableGPU.cuh:15,18); it does not change the device code.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may use the output.
"""
import ctypes
import json
import os
import sys

REF_ROOT = os.environ.get("LATTICE_REF_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")

POS_DIMS = (3, 5)
# (kernel, template-arg tuples).  V / nc lists cover the test + bench configurations.
VALS_SMALL = (1, 3, 4, 8)
VALS_ALL = (1, 3, 4, 8, 16, 32, 64, 128)
VALS_CONV = (1, 3, 4, 8, 16, 32, 64, 96, 128, 192, 256, 384, 512)   # channel widths met by the LatticeNet architectures (ShapeNet, ScanNet, SemanticKITTI)
CLASSIFY = ((32, 7), (64, 16), (128, 7), (128, 20), (8, 4), (256, 20), (128, 21))
CLASSIFY_D5 = ((32, 7), (64, 16), (8, 4))          # pos_dim 5 (xyz + rgb lattices): enough to pin the d = 5 code path


def name_expressions():
    names = []
    for d in POS_DIMS:
        names.append(f"kernel_splat<{d},1>")
        names.append(f"coarsen<{d}>")
        for v in (VALS_CONV if d == 3 else VALS_ALL):
            names.append(f"im2row<{d},{v}>")
            names.append(f"row2im<{d},{v}>")
        for v in VALS_ALL:
            names.append(f"splatCacheNaive<{d},{v}>")
            names.append(f"slice_with_precomputation<{d},{v}>")
            names.append(f"slice_no_precomputation<{d},{v}>")
            names.append(f"slice_backwards_with_precomputation_no_homogeneous<{d},{v}>")
        for v in VALS_SMALL:
            names.append(f"distribute<{d},{v}>")
            names.append(f"im2rowindices<{d},{v}>")
            names.append(f"gather_with_precomputation<{d},{v}>")
            names.append(f"gather_backwards_with_precomputation<{d},{v}>")
        for v, nc in (CLASSIFY if d == 3 else CLASSIFY_D5):
            names.append(f"slice_classify_with_precomputation<{d},{v},{nc}>")
            names.append(f"slice_classify_backwards_with_precomputation<{d},{v},{nc}>")
    return names


def _load_nvrtc():
    for cand in ("/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.12"):
        try:
            return ctypes.CDLL(cand)
        except OSError:
            continue
    raise RuntimeError("libnvrtc not found")


def build(out_dir=OUT_DIR, arch="compute_100", verbose=True):
    kdir = os.path.join(REF_ROOT, "include", "lattice_net", "kernels")
    main_path = os.path.join(kdir, "LatticeGPU.cuh")
    hdr_path = os.path.join(kdir, "HashTableGPU.cuh")
    if not (os.path.isfile(main_path) and os.path.isfile(hdr_path)):
        raise FileNotFoundError(f"reference kernels not found under {kdir}")
    with open(main_path, "rb") as f:
        main_src = f.read()
    with open(hdr_path, "rb") as f:
        hdr_src = f.read()

    nv = _load_nvrtc()
    nv.nvrtcGetErrorString.restype = ctypes.c_char_p

    def chk(res, what):
        if res != 0:
            raise RuntimeError(f"NVRTC {what} failed: {nv.nvrtcGetErrorString(res).decode()}")

    major, minor = ctypes.c_int(), ctypes.c_int()
    chk(nv.nvrtcVersion(ctypes.byref(major), ctypes.byref(minor)), "version")
    prog = ctypes.c_void_p()
    hdr_srcs = (ctypes.c_char_p * 1)(hdr_src)
    hdr_names = (ctypes.c_char_p * 1)(b"lattice_net/kernels/HashTableGPU.cuh")
    chk(nv.nvrtcCreateProgram(ctypes.byref(prog), main_src, b"LatticeGPU.cuh", 1, hdr_srcs, hdr_names), "create")
    names = name_expressions()
    for n in names:
        chk(nv.nvrtcAddNameExpression(prog, n.encode()), f"add name {n}")
    opts = [b"-std=c++11", b"--use_fast_math", f"-arch={arch}".encode(), b"-default-device"]
    c_opts = (ctypes.c_char_p * len(opts))(*opts)
    res = nv.nvrtcCompileProgram(prog, len(opts), c_opts)
    log_size = ctypes.c_size_t()
    nv.nvrtcGetProgramLogSize(prog, ctypes.byref(log_size))
    if log_size.value > 1:
        log = ctypes.create_string_buffer(log_size.value)
        nv.nvrtcGetProgramLog(prog, log)
        if res != 0 or verbose:
            txt = log.value.decode(errors="replace")
            sys.stderr.write(txt[-4000:] if res != 0 else f"[build_ref] nvrtc log: {len(txt)} bytes (warnings)\n")
    chk(res, "compile")
    size = ctypes.c_size_t()
    chk(nv.nvrtcGetPTXSize(prog, ctypes.byref(size)), "ptx size")
    ptx = ctypes.create_string_buffer(size.value)
    chk(nv.nvrtcGetPTX(prog, ptx), "get ptx")
    lowered = {}
    for n in names:
        p = ctypes.c_char_p()
        chk(nv.nvrtcGetLoweredName(prog, n.encode(), ctypes.byref(p)), f"lowered {n}")
        lowered[n] = p.value.decode()
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "lattice_ref.ptx"), "wb") as f:
        f.write(ptx.raw[: size.value].rstrip(b"\0"))
    meta = {"nvrtc": f"{major.value}.{minor.value}", "options": [o.decode() for o in opts], "names": lowered,
            "source": "include/lattice_net/kernels/{LatticeGPU,HashTableGPU}.cuh (read in place from the reference tree)"}
    with open(os.path.join(out_dir, "lattice_ref.names.json"), "w") as f:
        json.dump(meta, f, indent=1)
    nv.nvrtcDestroyProgram(ctypes.byref(prog))
    _build_probe(nv, chk, out_dir, arch)
    if verbose:
        print(f"[build_ref] NVRTC {major.value}.{minor.value}: {len(names)} kernels -> {out_dir}/lattice_ref.ptx ({size.value} bytes)")
    return out_dir


PROBE_SRC = b"""
// hardware results of the approximate instructions the reference's fast-math build relies on
extern "C" __global__ void probe_rsqrt(const float* in, float* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(in[i])); out[i] = y; }
}
"""


def _build_probe(nv, chk, out_dir, arch):
    prog = ctypes.c_void_p()
    chk(nv.nvrtcCreateProgram(ctypes.byref(prog), PROBE_SRC, b"probe.cu", 0, None, None), "create probe")
    opts = [f"-arch={arch}".encode()]
    c_opts = (ctypes.c_char_p * len(opts))(*opts)
    chk(nv.nvrtcCompileProgram(prog, len(opts), c_opts), "compile probe")
    size = ctypes.c_size_t()
    chk(nv.nvrtcGetPTXSize(prog, ctypes.byref(size)), "probe ptx size")
    ptx = ctypes.create_string_buffer(size.value)
    chk(nv.nvrtcGetPTX(prog, ptx), "probe ptx")
    with open(os.path.join(out_dir, "probe.ptx"), "wb") as f:
        f.write(ptx.raw[: size.value].rstrip(b"\0"))
    nv.nvrtcDestroyProgram(ctypes.byref(prog))


if __name__ == "__main__":
    build()
