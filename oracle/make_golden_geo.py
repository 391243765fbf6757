"""oracle/make_golden_geo.py -- golden vectors for GeometricStructureEmbedding.get_embedding_indices
 from the LIVE reference in this container.

TEST INFRASTRUCTURE ONLY.  Imports the unmodified reference module model/transformer.py (through
oracle/ref_live.py), builds GeometricStructureEmbedding from config/base.yaml's values and records its
output on seeded clouds into tests/golden/geo_embedding.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_live  # noqa: E402


class _Cfg:
    sigma_d, sigma_a, angle_k, reduction_a, hidden_dim = 0.2, 15, 3, "max", 256  # config/base.yaml:27-31


def main():
    ref_live.load()
    sys.path.insert(0, ref_live.REF_PEM)           # transformer.py uses package-style imports (utils.*, model.*)
    from model import transformer                  # the reference module, unmodified
    gse = transformer.GeometricStructureEmbedding(_Cfg())
    out = {}
    for name, (B, N, seed) in {"sparse": (2, 197, 1), "small": (3, 33, 2)}.items():
        g = torch.Generator().manual_seed(seed)
        pts = (torch.randn(B, N, 3, generator=g) * 0.4).contiguous()
        d, a = gse.get_embedding_indices(pts)
        out[name + "_points"], out[name + "_d"], out[name + "_a"] = pts.numpy(), d.numpy(), a.numpy()
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "geo_embedding.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
