"""GeometricStructureEmbedding.get_embedding_indices (reference model/transformer.py:304-341; SURVEY
section 8f, rank 3): the oracle against the golden vectors recorded from the live reference and
against the live module itself, and the CUDA kernel against both."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

G = os.path.join(GOLDEN, "geo_embedding.npz")


@pytest.mark.parametrize("name", ["sparse", "small"])
def test_oracle_matches_golden(oracle, name):
    d = np.load(G)
    di, ai, _ = oracle.geo_embedding_indices(d[name + "_points"])
    np.testing.assert_array_equal(di.numpy(), d[name + "_d"])      # same torch ops: bit identical
    np.testing.assert_array_equal(ai.numpy(), d[name + "_a"])


def test_oracle_matches_live_reference(oracle):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_live
    if not ref_live.available():
        pytest.skip("/root/reference is not present on this machine")
    ref_live.load()
    sys.path.insert(0, ref_live.REF_PEM)
    from model import transformer

    class Cfg:
        sigma_d, sigma_a, angle_k, reduction_a, hidden_dim = 0.35, 10, 4, "max", 64

    gse = transformer.GeometricStructureEmbedding(Cfg())
    pts = torch.randn(2, 57, 3, generator=torch.Generator().manual_seed(9)) * 0.3
    d, a = gse.get_embedding_indices(pts)
    od, oa, _ = oracle.geo_embedding_indices(pts, 0.35, 10, 4)
    assert torch.equal(d, od) and torch.equal(a, oa)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sparse", "small"])
def test_kernel_matches_golden(pkg, dev, name):
    d = np.load(G)
    pts = torch.from_numpy(d[name + "_points"]).to(dev)
    di, ai = pkg.transformer.get_embedding_indices(pts)
    # distances: |p_i|^2 - 2 p_i.p_j + |p_j|^2 cancels (the reference's own diagonal is sqrt of rounding
    # noise, up to 3e-3 in these units), so the comparison is on the squared distance it is computed from
    sq = lambda x: (np.asarray(x, dtype=np.float64) * 0.2) ** 2
    np.testing.assert_allclose(sq(di.cpu().numpy()), sq(d[name + "_d"]), rtol=0, atol=3e-6)
    # angles in units of sigma_a degrees (range 0..12); atan2 near a clamp amplifies distance rounding
    assert np.abs(ai.cpu().numpy() - d[name + "_a"]).max() < 2e-3
    assert np.median(np.abs(ai.cpu().numpy() - d[name + "_a"])) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,k,sd,sa", [(16, 197, 3, 0.2, 15.0), (1, 5, 3, 0.2, 15.0), (3, 1000, 8, 0.5, 7.5),
                                          (2, 300, 1, 0.1, 30.0)])
def test_kernel_matches_oracle(pkg, oracle, dev, B, N, k, sd, sa):
    g = torch.Generator().manual_seed(N + k)
    pts = (torch.randn(B, N, 3, generator=g) * 0.4).contiguous()
    od, oa, oknn = oracle.geo_embedding_indices(pts, sd, sa, k)
    d, a = pkg.transformer.get_embedding_indices(pts.to(dev), sd, sa, k)
    assert tuple(d.shape) == (B, N, N) and tuple(a.shape) == (B, N, N, k)
    assert ((d.cpu().double() * sd) ** 2 - (od.double() * sd) ** 2).abs().max().item() < 3e-6  # see above
    # rows whose k+1 nearest are separated by more than rounding must agree everywhere
    dist = od * sd
    srt = dist.sort(dim=2)[0][:, :, : k + 2]
    clear = (srt[:, :, 1:] - srt[:, :, :-1]).min(dim=2)[0] > 1e-5          # (B,N)
    assert clear.float().mean().item() > 0.95
    err = (a.cpu() - oa).abs().amax(dim=(2, 3))
    assert err[clear].max().item() < 5e-3, err[clear].max().item()
    assert (a.cpu() - oa).abs()[clear].median().item() < 5e-6


@pytest.mark.gpu
def test_argument_checks(pkg, dev):
    p = torch.randn(1, 16, 3, device=dev)
    with pytest.raises(RuntimeError):
        pkg.transformer.get_embedding_indices(p, angle_k=9)       # more neighbours than the kernel keeps
    with pytest.raises(RuntimeError, match="contiguous"):
        pkg.transformer.get_embedding_indices(torch.randn(1, 3, 16, device=dev).transpose(1, 2))
