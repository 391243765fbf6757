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
