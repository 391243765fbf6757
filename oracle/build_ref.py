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
