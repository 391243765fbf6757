"""oracle/ref_live.py -- import the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
this module is used only (a) by oracle/make_golden.py to write the committed
fixtures under tests/golden/ and (b) by the ``not gpu`` tests that pin the
oracle against the live reference (skipped when the tree is absent).

The reference's ``utils/model_utils.py`` imports ``pointnet2_utils`` which
imports ``pointnet2._ext``; we satisfy that with the reference's own CPU
extension built by oracle/build_ref.py (oracle/_ref/), registered under the
module names the reference expects.
"""
import os
import sys
import types

REF_PEM = "/root/reference/SAM-6D/Pose_Estimation_Model"


def available():
    return os.path.isdir(REF_PEM)


def pem_root():
    """The reference PEM tree: the live one in this container, else the snapshot of its call-site files that
    oracle/build_ref.snapshot_reference_py() left under oracle/_ref/pem_py (what the GPU box has)."""
    if available():
        return REF_PEM
    snap = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pem_py")
    return snap if os.path.isdir(snap) else None


def load_call_sites(ext_module, root=None):
    """Import the reference's UNMODIFIED pointnet2_utils / model_utils / matching modules with `ext_module`
    registered as `pointnet2._ext` (the drop-in under test, or the reference's own build).  Returns a dict of
    the imported modules.  Any previously imported copies are dropped first."""
    root = root or pem_root()
    if root is None:
        raise FileNotFoundError("no reference tree and no snapshot under oracle/_ref/pem_py")
    for name in ("pointnet2_utils", "pytorch_utils", "model_utils", "loss_utils", "transformer",
                 "coarse_point_matching", "fine_point_matching", "utils.model_utils", "utils",
                 "model.pointnet2.pointnet2_utils", "model.pointnet2", "model"):
        sys.modules.pop(name, None)
    pkg = types.ModuleType("pointnet2")
    pkg.__path__ = []
    pkg._ext = ext_module
    sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = ext_module
    for sub in ("", "utils", "model", os.path.join("model", "pointnet2")):
        p = os.path.join(root, sub)
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import importlib
    out = {}
    for name in ("pointnet2_utils", "model_utils", "transformer", "coarse_point_matching", "fine_point_matching"):
        out[name] = importlib.import_module(name)
    out["root"] = root
    return out


def load():
    """Returns (model_utils, pointnet2_utils, ref_ext) of the live reference."""
    if not available():
        raise FileNotFoundError(REF_PEM)
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import build_ref
    ext = build_ref.load_ref_ext()
    pkg = types.ModuleType("pointnet2")
    pkg.__path__ = []
    pkg._ext = ext
    sys.modules.setdefault("pointnet2", pkg)
    sys.modules["pointnet2._ext"] = ext
    for sub in ("", "utils", "model", os.path.join("model", "pointnet2")):
        p = os.path.join(REF_PEM, sub)
        if p not in sys.path:
            sys.path.append(p)
    import model_utils
    import pointnet2_utils
    return model_utils, pointnet2_utils, ext
