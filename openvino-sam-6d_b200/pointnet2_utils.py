"""Host-side mirror of the reference's op wrappers (model/pointnet2/pointnet2_utils.py):
``furthest_point_sample``, ``gather_operation``, ``grouping_operation``, ``ball_query`` and the
``QueryAndGroup`` module, same names and argument meaning, inference only (no autograd
graph is recorded: the reference runs this path under ``torch.no_grad()``,
run_inference_custom_pytorch.py:444).

``QueryAndGroup`` uses the fused kernel (ball query + both groupings + centre subtraction +
concatenation written once); ``fused=False`` composes the three separate ops exactly like
pointnet2_utils.py:343-384 does.
"""
import torch
import torch.nn as nn

from . import _lib
from . import pointnet2_ext as _ext


def furthest_point_sample(xyz, npoint):
    """FurthestPointSampling.forward (pointnet2_utils.py:58-80): ``npoint`` is a tensor whose
    shape[0] is the sample count (an int is accepted too)."""
    n = int(npoint.shape[0]) if isinstance(npoint, torch.Tensor) else int(npoint)
    return _ext.furthest_point_sampling(xyz, n)


def gather_operation(features, idx):
    """GatherOperation.forward (pointnet2_utils.py:96-118): (B,C,N),(B,npoint) -> (B,C,npoint)."""
    return _ext.gather_points(features, idx)


def grouping_operation(features, idx):
    """GroupingOperation.forward (pointnet2_utils.py:234-255)."""
    return _ext.group_points(features, idx)


def ball_query(new_xyz, xyz, radius, nsample):
    """BallQuery.forward (pointnet2_utils.py:289-293)."""
    return _ext.ball_query(new_xyz, xyz, radius, nsample)


def gather_rows(feats, idx):
    """Point-major gather: feats (B,N,C) (any batch/row strides with unit channel stride),
    idx (B,M) i32 -> (B,M,C).  Equals
    ``gather_operation(feats.transpose(1,2).contiguous(), idx).transpose(1,2)``
    (model_utils.py:77-80, transformer.py:667-671) without materialising the transposes."""
    if not isinstance(feats, torch.Tensor) or feats.dtype != torch.float32:
        raise RuntimeError("feats must be a float tensor")
    if not feats.is_cuda:
        raise RuntimeError("feats must be a CUDA tensor (libpem_b200 has no CPU path)")
    if feats.ndim != 3:
        raise RuntimeError("feats must be (B,N,C)")
    if feats.stride(2) != 1:
        feats = feats.contiguous()
    _ext._check(idx, "idx", torch.int32, "an int")
    if idx.device != feats.device or idx.shape[0] != feats.shape[0]:
        raise RuntimeError("idx must be (B,M) on the device of feats")
    B, N, C = feats.shape
    M = idx.shape[1]
    out = torch.empty((B, M, C), dtype=torch.float32, device=feats.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().pem_gather_rows(feats.data_ptr(), idx.data_ptr(), B, N, C, M,
                                              feats.stride(0), feats.stride(1), out.data_ptr(), 0,
                                              _ext._stream()), "pem_gather_rows")
    return out


def query_and_group(xyz, new_xyz, features, radius, nsample, return_idx=False):
    """Fused QueryAndGroup, use_xyz=True: xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) or
    None -> (B,3+C,M,nsample) [, idx (B,M,nsample)]."""
    _ext._check(xyz, "xyz", torch.float32, "a float")
    _ext._check(new_xyz, "new_xyz", torch.float32, "a float")
    C = 0
    if features is not None:
        _ext._check(features, "features", torch.float32, "a float")
        C = features.shape[1]
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    out = torch.empty((B, 3 + C, M, nsample), dtype=torch.float32, device=xyz.device)
    idx = torch.empty((B, M, nsample), dtype=torch.int32, device=xyz.device) if return_idx else None
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().pem_query_and_group(
            xyz.data_ptr(), new_xyz.data_ptr(), None if features is None else features.data_ptr(),
            B, N, M, C, float(radius), int(nsample), out.data_ptr(),
            None if idx is None else idx.data_ptr(), _ext._stream()), "pem_query_and_group")
    return (out, idx) if return_idx else out


def pe_group2(xyz, radius1=0.1, nsample1=32, radius2=0.2, nsample2=64, centre_eps=1e-8):
    """Both grouping scales of PositionalEncoding.forward (fine_point_matching.py:113-137) in one
    kernel: returns (QueryAndGroup(r1,ns1)(xyz, xyz+eps, xyz^T), QueryAndGroup(r2,ns2)(...)),
    each (B,6,N,ns), bit-identical to the two separate calls."""
    _ext._check(xyz, "xyz", torch.float32, "a float")
    B, N, _ = xyz.shape
    out1 = torch.empty((B, 6, N, nsample1), dtype=torch.float32, device=xyz.device)
    out2 = torch.empty((B, 6, N, nsample2), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().pem_pe_group2(xyz.data_ptr(), B, N, float(centre_eps), float(radius1),
                                            int(nsample1), out1.data_ptr(), float(radius2),
                                            int(nsample2), out2.data_ptr(), _ext._stream()),
                   "pem_pe_group2")
    return out1, out2


def pe_masks(xyz, radius=0.2, centre_eps=1e-8, rel_margin=1e-3, max_ctas=0):
    """Tensor-core candidate filter for pe_group2_masked: one bit per (centre, point) pair, set when
    the point may lie within `radius` of the centre.  May be computed on the cloud BEFORE a rigid
    motion (ball membership is invariant; rel_margin covers the rounding of the motion)."""
    _ext._check(xyz, "xyz", torch.float32, "a float")
    B, N, _ = xyz.shape
    L = _lib.lib()
    nbytes = L.pem_pe_masks_bytes(B, N)
    masks = torch.empty(nbytes, dtype=torch.uint8, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(L.pem_pe_masks(xyz.data_ptr(), B, N, float(centre_eps), float(radius), float(rel_margin),
                                  int(max_ctas), masks.data_ptr(), nbytes, _ext._stream()), "pem_pe_masks")
    return masks


def pe_group2_masked(xyz, masks, radius1=0.1, nsample1=32, radius2=0.2, nsample2=64, centre_eps=1e-8):
    """pe_group2(xyz) visiting only the candidates of `masks` (made with radius >= both radii, on this
    cloud or on it before a rigid motion): bit-identical outputs."""
    _ext._check(xyz, "xyz", torch.float32, "a float")
    B, N, _ = xyz.shape
    out1 = torch.empty((B, 6, N, nsample1), dtype=torch.float32, device=xyz.device)
    out2 = torch.empty((B, 6, N, nsample2), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().pem_pe_group2_masked(xyz.data_ptr(), B, N, float(centre_eps), float(radius1),
                                                   int(nsample1), out1.data_ptr(), float(radius2),
                                                   int(nsample2), out2.data_ptr(), masks.data_ptr(),
                                                   masks.numel(), _ext._stream()), "pem_pe_group2_masked")
    return out1, out2


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:303-403.  Only the configuration PEM uses is accelerated
    (use_xyz=True, no uniform re-sampling); other flags raise."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False, fused=True):
        super().__init__()
        if sample_uniformly or ret_unique_cnt or normalize_xyz or ret_grouped_xyz:
            raise NotImplementedError("only the PEM configuration of QueryAndGroup is provided")
        self.radius, self.nsample, self.use_xyz, self.fused = radius, nsample, use_xyz, fused

    def forward(self, xyz, new_xyz, features=None):
        if self.fused and self.use_xyz:
            return query_and_group(xyz, new_xyz, features, self.radius, self.nsample)
        idx = ball_query(new_xyz, xyz, self.radius, self.nsample)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
