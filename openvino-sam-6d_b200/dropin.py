"""Make the reference's own Python call sites run on libpem_b200.

``install_pointnet2_ext()`` registers this package's ``pointnet2_ext`` module as
``pointnet2._ext`` in ``sys.modules``: the reference's ``pointnet2_utils.py`` does
``import pointnet2._ext as _ext`` (pointnet2_utils.py:25-33) and then runs unmodified.
``patch_model_utils(mu)`` swaps the geometric functions and the operator-seam shells of an already imported reference
``utils/model_utils.py`` module for the accelerated ones (same signatures).
"""
import sys
import types


def install_pointnet2_ext():
    from . import pointnet2_ext
    pkg = sys.modules.get("pointnet2")
    if pkg is None:
        pkg = types.ModuleType("pointnet2")
        pkg.__path__ = []
        sys.modules["pointnet2"] = pkg
    pkg._ext = pointnet2_ext
    sys.modules["pointnet2._ext"] = pointnet2_ext
    return pointnet2_ext


def patch_model_utils(mu):
    from . import model_utils as fast
    for name in ("sample_pts_feats", "compute_feature_similarity", "compute_coarse_Rt",
                 "compute_fine_Rt", "weighted_procrustes", "WeightedProcrustes", "CustomSearchSorted",
                 "CustomSVD", "CustomSVDu", "CustomSVDv", "CustomDet"):
        setattr(mu, name, getattr(fast, name))
    return mu


def patch_call_sites(*modules):
    """The reference's matching modules bind the functions by name (`from model_utils import compute_coarse_Rt`,
    coarse_point_matching.py:5-9, fine_point_matching.py:9), so patching `model_utils` alone does not reach them:
    swap every accelerated name that a module holds."""
    from . import model_utils as fast
    names = ("sample_pts_feats", "compute_feature_similarity", "compute_coarse_Rt", "compute_fine_Rt",
             "weighted_procrustes", "WeightedProcrustes")
    for mod in modules:
        for name in names:
            if hasattr(mod, name):
                setattr(mod, name, getattr(fast, name))
    return modules
