"""Mirror of the geometry in the reference's model/transformer.py that sits immediately upstream of
the coarse similarity: GeometricStructureEmbedding.get_embedding_indices
(transformer.py:304-341).  The embedding layers themselves (sinusoidal embedding, proj_d / proj_a,
the max over k) stay in PyTorch."""
import torch

from . import _lib
from . import pointnet2_ext as _ext


def get_embedding_indices(points, sigma_d=0.2, sigma_a=15.0, angle_k=3):
    """points (B,N,3) float32 CUDA -> (d_indices (B,N,N), a_indices (B,N,N,k)), as the reference's
    method returns them (config/base.yaml:27-29 for the defaults).  One kernel instead of ~20 torch ops
    and a (B,N,N,k,3) expansion."""
    _ext._check(points, "points", torch.float32, "a float")
    B, N, _ = points.shape
    d = torch.empty((B, N, N), dtype=torch.float32, device=points.device)
    a = torch.empty((B, N, N, int(angle_k)), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().pem_geo_embedding_indices(points.data_ptr(), B, N, int(angle_k), float(sigma_d),
                                                        float(sigma_a), d.data_ptr(), a.data_ptr(),
                                                        _ext._stream()), "pem_geo_embedding_indices")
    return d, a


def patch_geometric_structure_embedding(transformer_module):
    """Replace GeometricStructureEmbedding.get_embedding_indices of the reference's (imported)
    transformer module; forward() and the learned projections are untouched."""
    cls = transformer_module.GeometricStructureEmbedding

    @torch.no_grad()
    def _indices(self, points):
        return get_embedding_indices(points.contiguous(), self.sigma_d, self.sigma_a, self.angle_k)

    cls.get_embedding_indices = _indices
    return cls
