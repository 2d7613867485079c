"""Build the UNMODIFIED reference CPU extensions into oracle/_ref/ (test infrastructure only).

This is checker code: nothing under d3d_b200/ may import it.  It compiles the reference's own
C++ sources *where they lie* under /root/reference (never copied into this repo) with
torch.utils.cpp_extension, following SURVEY.md section 8(c):

    box_impl   <- d3d/box/{impl,utils,iou,nms,dist}.cpp      (reference d3d/box/CMakeLists.txt:1-13)
    voxel_impl <- d3d/voxel/{impl,voxelize}.cpp              (reference d3d/voxel/CMakeLists.txt:1-11)
    point_impl <- d3d/point/{impl,scatter}.cpp               (reference d3d/point/CMakeLists.txt:1-13)

Outputs (.so + ninja files) go to oracle/_ref/<name>/ only.  oracle/_ref/ is git-ignored but is
NOT gpurun-ignored, so the built .so files travel to the GPU box where /root/reference does not
exist; there `load_ref()` only dlopens the prebuilt modules.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("D3D_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

MODULES = {
    "box_impl": ["d3d/box/impl.cpp", "d3d/box/utils.cpp", "d3d/box/iou.cpp", "d3d/box/nms.cpp", "d3d/box/dist.cpp"],
    "voxel_impl": ["d3d/voxel/impl.cpp", "d3d/voxel/voxelize.cpp"],
    "point_impl": ["d3d/point/impl.cpp", "d3d/point/scatter.cpp"],
}


WRAP_SO = os.path.join(OUT, "libdgal_wrap.so")   # d3d/dgal_wrap.h behind a C ABI (oracle/dgal_wrap_shim.cpp), plain g++


def build_wrap():
    """g++ on the shim, which includes the reference's d3d/dgal_wrap.h where it lies. Idempotent."""
    if os.path.exists(WRAP_SO) or not os.path.isdir(REF_ROOT):
        return os.path.exists(WRAP_SO)
    import subprocess
    os.makedirs(OUT, exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-w", "-shared", "-fPIC", "-I", REF_ROOT, "-I", os.path.join(REF_ROOT, "thirdparty"),
                           os.path.join(HERE, "dgal_wrap_shim.cpp"), "-o", WRAP_SO])
    return True


CUDA_MODULE = "box_impl_cuda"   # the reference's OWN CUDA kernels (iou_cuda.cu, nms_cuda.cu, dist_cuda.cu), unmodified, built for sm_100a:
CUDA_SOURCES = ["d3d/box/impl.cpp", "d3d/box/utils.cpp", "d3d/box/iou.cpp", "d3d/box/nms.cpp", "d3d/box/dist.cpp",   # the secondary GPU baseline of bench.py
                "d3d/box/iou_cuda.cu", "d3d/box/nms_cuda.cu", "d3d/box/dist_cuda.cu"]


def build_cuda(verbose=False):
    """Compile the reference's box extension WITH its CUDA kernels (a few minutes).  Idempotent; optional."""
    if os.path.exists(so_path(CUDA_MODULE)) or not os.path.isdir(REF_ROOT):
        return os.path.exists(so_path(CUDA_MODULE))
    from torch.utils.cpp_extension import load
    bdir = os.path.join(OUT, CUDA_MODULE)
    os.makedirs(bdir, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    load(name=CUDA_MODULE, sources=[os.path.join(REF_ROOT, s) for s in CUDA_SOURCES],
         extra_include_paths=[REF_ROOT, os.path.join(REF_ROOT, "thirdparty")],
         extra_cflags=["-O2", "-DNDEBUG", "-w", "-DBUILD_WITH_CUDA"],
         extra_cuda_cflags=["-O2", "-DNDEBUG", "-w", "-DBUILD_WITH_CUDA", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr"],
         build_directory=bdir, verbose=verbose)
    return os.path.exists(so_path(CUDA_MODULE))


def so_path(name):
    return os.path.join(OUT, name, name + ".so")


def have_ref():
    return all(os.path.exists(so_path(n)) for n in MODULES)


def build(verbose=False):
    """Compile the three reference extensions if the reference tree is present. Idempotent."""
    if not os.path.isdir(REF_ROOT):
        return have_ref()
    build_wrap()
    from torch.utils.cpp_extension import load
    for name, srcs in MODULES.items():
        if os.path.exists(so_path(name)):
            continue
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name,
             sources=[os.path.join(REF_ROOT, s) for s in srcs],
             extra_include_paths=[REF_ROOT, os.path.join(REF_ROOT, "thirdparty")],
             extra_cflags=["-O2", "-DNDEBUG", "-w"],
             build_directory=bdir, verbose=verbose)
    return have_ref()


_cache = {}


def load_ref(name):
    """dlopen a prebuilt reference extension (pybind11 module) from oracle/_ref/."""
    if name in _cache:
        return _cache[name]
    import torch  # noqa: F401  (the .so links against libtorch)
    p = so_path(name)
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("reference extensions available:", ok)
    if "--cuda" in sys.argv:
        print("reference CUDA extension available:", build_cuda(verbose="-v" in sys.argv))
