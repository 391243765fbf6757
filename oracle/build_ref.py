"""Build the UNMODIFIED reference pointnet2 CPU extension into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
path (openvino-sam-6d_b200/); only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may use it.

Recipe: the reference's own C++ sources are compiled *where they lie* under
/root/reference (never copied into this repo) with the flags of the CPU branch
of the reference's setup.py (PEM/model/pointnet2/setup.py:49-58:
``-O3 -DCUDA_AVAILABLE=0``).  Output: oracle/_ref/pointnet2_ref_ext.so, a
pybind11 torch extension exposing the nine functions of
PEM/model/pointnet2/_ext_src/src/bindings.cpp:11-24.  oracle/_ref/ is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import glob
import os
import sys

REF_ROOT = "/root/reference/SAM-6D/Pose_Estimation_Model"
EXT_SRC = os.path.join(REF_ROOT, "model", "pointnet2", "_ext_src")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "pointnet2_ref_ext"


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    """Compile if the reference tree is present; otherwise use the prebuilt .so."""
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(EXT_SRC):
        raise FileNotFoundError(
            "reference sources not found at %s and no prebuilt %s" % (EXT_SRC, so_path()))
    from torch.utils.cpp_extension import load
    os.makedirs(OUT_DIR, exist_ok=True)
    load(
        name=NAME,
        sources=sorted(glob.glob(os.path.join(EXT_SRC, "src", "*.cpp"))),
        extra_include_paths=[os.path.join(EXT_SRC, "include")],
        extra_cflags=["-O3", "-DCUDA_AVAILABLE=0"],
        build_directory=OUT_DIR,
        is_python_module=True,
        verbose=verbose,
    )
    return so_path()


CUDA_DIR = os.path.join(HERE, "_ref_cuda")
CUDA_NAME = "pointnet2_ref_cuda_ext"


def cuda_so_path():
    return os.path.join(CUDA_DIR, CUDA_NAME + ".so")


def build_cuda(verbose=False):
    """The reference's CUDA kernels (the four *_gpu.cu files of _ext_src/src, compiled UNMODIFIED from where
    they lie, with the nvcc defines of setup.py:36-42) for sm_100 -- the same-box GPU baseline the pointnet2
    kernels are timed beside and, where the tie / skip rules allow, checked against.  The host
    side is oracle/ref_cuda_shim.cpp, not the reference's .cpp files: in the reference's own CUDA build those
    re-define CUDA_AVAILABLE to 0 under g++ (utils.h:9-14) and raise "CUDA not available" for every CUDA
    tensor, so its kernels cannot be reached through its bindings.  Output: oracle/_ref_cuda/ (git-ignored,
    travels to the GPU box)."""
    if os.path.exists(cuda_so_path()):
        return cuda_so_path()
    if not os.path.isdir(EXT_SRC):
        raise FileNotFoundError("reference sources not found at %s and no prebuilt %s" % (EXT_SRC, cuda_so_path()))
    from torch.utils.cpp_extension import load
    os.makedirs(CUDA_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    load(
        name=CUDA_NAME,
        sources=[os.path.join(HERE, "ref_cuda_shim.cpp")] + sorted(glob.glob(os.path.join(EXT_SRC, "src", "*.cu"))),
        extra_include_paths=[os.path.join(EXT_SRC, "include")],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                           "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                           "-DCUDA_AVAILABLE=1"],
        with_cuda=True,
        build_directory=CUDA_DIR,
        is_python_module=True,
        verbose=verbose,
    )
    return cuda_so_path()


def load_ref_cuda_ext():
    import importlib.util
    import torch  # noqa: F401
    path = build_cuda()
    spec = importlib.util.spec_from_file_location(CUDA_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# The reference's own Python call sites, as a snapshot for the GPU box (where /root/reference does not exist):
# the `-m gpu` boundary test runs them UNMODIFIED on top of the drop-in.  Copied at build time into the
# git-ignored oracle/_ref/pem_py/ (test infrastructure; never part of the repository's history).
PY_SNAPSHOT = os.path.join(OUT_DIR, "pem_py")
PY_FILES = ("model/coarse_point_matching.py", "model/fine_point_matching.py", "model/transformer.py",
            "model/pointnet2/pointnet2_utils.py", "model/pointnet2/pytorch_utils.py", "utils/model_utils.py",
            "utils/loss_utils.py", "config/base.yaml")


def snapshot_reference_py():
    import shutil
    if not os.path.isdir(REF_ROOT):
        return PY_SNAPSHOT if os.path.isdir(PY_SNAPSHOT) else None
    for rel in PY_FILES:
        dst = os.path.join(PY_SNAPSHOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
    return PY_SNAPSHOT


def load_ref_ext():
    """Import the prebuilt reference extension (no compilation if already built)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    path = build()
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    m = load_ref_ext()
    print("built", p, [n for n in dir(m) if not n.startswith("_")])
