"""Install the reference's own Python layer, UNMODIFIED, into baseline/_ref (git-ignored; it travels to the GPU box with
gpurun like a built artefact) so that `bench.py --impl reference` drives the reference's lattice_funcs / lattice_modules /
models code rather than a restatement of it.

`pip install /root/reference` cannot work here: setup.py builds the C++ host library through CMake against EasyPBR,
Boost, Eigen, loguru -- none present, no network.  What is installed is therefore the pure-Python part:
    latticenet_py/lattice/{lattice_funcs,lattice_modules,lattice_wrapper,models,lovasz_loss}.py, latticenet_py/utils/utils.py
byte for byte (a sha256 manifest is written next to them); the compiled `latticenet` module they import is provided by
oracle/ref_arm.py on top of the reference's own CUDA kernels (oracle/_ref, NVRTC build of the unmodified headers).

    python -m oracle.install_ref_py          (also run by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["lattice/lattice_funcs.py", "lattice/lattice_modules.py", "lattice/lattice_wrapper.py", "lattice/models.py",
         "lattice/lovasz_loss.py", "utils/utils.py"]


def installed():
    return os.path.isfile(os.path.join(DST, "MANIFEST.json"))


def install(ref_root=None, verbose=True):
    ref_root = ref_root or os.environ.get("LATTICE_REF_ROOT", "/root/reference")
    src = os.path.join(ref_root, "latticenet_py")
    if not os.path.isdir(src):
        raise RuntimeError(f"{src} not found")
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, "latticenet_py", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "AIS-Bonn/lattice_net latticenet_py (unmodified copies)", "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"[install_ref_py] {len(FILES)} files -> {DST}")
    return DST


if __name__ == "__main__":
    install()
