"""PositionalEncoding (reference model/fine_point_matching.py:102-144): the
oracle against the golden vectors recorded from the live reference module and against the live module
itself; the fused tensor-core path (neighbour lists + three chained split-bf16 GEMMs + amax) against both."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

G = os.path.join(GOLDEN, "pe_mlp.npz")


def _branches(d):
    br = lambda n: [(d[f"{n}{l}_w"], d[f"{n}{l}_g"], d[f"{n}{l}_b"], d[f"{n}{l}_m"], d[f"{n}{l}_v"], float(d[f"{n}{l}_eps"]))
                    for l in range(3)]
    return br("a"), br("b")


def test_oracle_matches_golden(oracle):
    d = np.load(G)
    b1, b2 = _branches(d)
    out, f1, f2 = oracle.positional_encoding(d["pts"], b1, b2, d["w3"], d["b3"])
    np.testing.assert_array_equal(f1.numpy(), d["f1"])     # same torch ops as the module: bit identical
    np.testing.assert_array_equal(f2.numpy(), d["f2"])
    np.testing.assert_array_equal(out.numpy(), d["out"])


def test_oracle_matches_live_reference(oracle):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_live
    if not ref_live.available():
        pytest.skip("/root/reference is not present on this machine")
    import contextlib
    import io
    import make_golden_pe
    ref_live.load()
    sys.path.insert(0, ref_live.REF_PEM)
    from model import fine_point_matching
    torch.manual_seed(11)
    pe = fine_point_matching.PositionalEncoding(64).eval()
    pts = torch.randn(1, 300, 3, generator=torch.Generator().manual_seed(12)) * 0.2
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        want = pe(pts)
    got, _, _ = oracle.positional_encoding(pts, make_golden_pe.branch_state(pe.mlp1), make_golden_pe.branch_state(pe.mlp2),
                                           pe.mlp3[0].weight.detach().reshape(64, 256).numpy(),
                                           pe.mlp3[0].bias.detach().numpy())
    assert torch.equal(got, want)


@pytest.mark.gpu
def test_neighbour_lists_bit_exact(pkg, oracle, dev):
    d = np.load(G)
    pts = torch.from_numpy(d["pts"]).to(dev)
    i1, i2 = pkg.positional_encoding.pe_neighbours(pts)
    ctr = d["pts"] + np.float32(1e-8)
    np.testing.assert_array_equal(i1.cpu().numpy(), oracle.ball_query(ctr, d["pts"], 0.1, 32))
    np.testing.assert_array_equal(i2.cpu().numpy(), oracle.ball_query(ctr, d["pts"], 0.2, 64))


@pytest.mark.gpu
def test_fused_path_matches_golden(pkg, dev):
    d = np.load(G)
    b1, b2 = _branches(d)
    P = pkg.positional_encoding
    out, pooled = P.positional_encoding(torch.from_numpy(d["pts"]).to(dev), P.fold_branch(b1), P.fold_branch(b2),
                                        d["w3"], d["b3"], return_pooled=True)
    f = np.concatenate([d["f1"], d["f2"]], 1).transpose(0, 2, 1)          # (B,N,256) as the fused path lays it out
    scale = np.abs(f).max()
    # three chained split-bf16 layers: ~2^-16 relative per layer
    assert np.abs(pooled.cpu().numpy() - f).max() < 2e-4 * scale, (np.abs(pooled.cpu().numpy() - f).max(), scale)
    assert np.abs(out.cpu().numpy() - d["out"]).max() < 2e-4 * np.abs(d["out"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,scale_pts", [(3, 2048, 0.3), (1, 100, 0.1), (2, 777, 0.05)])
def test_fused_branch_matches_oracle(pkg, oracle, dev, B, N, scale_pts):
    """Random weights and BatchNorm state, clouds of several sizes and densities (empty balls, full
    balls, sizes that are not multiples of the 128-row tile)."""
    g = torch.Generator().manual_seed(N)
    pts = (torch.randn(B, N, 3, generator=g) * scale_pts).contiguous()
    def rand_branch():
        out, ci = [], 6
        for co in (32, 64, 128):
            out.append(((torch.randn(co, ci, generator=g) * (2.0 / ci) ** 0.5).numpy(), (0.5 + torch.rand(co, generator=g)).numpy(),
                        (0.3 * torch.randn(co, generator=g)).numpy(), (0.2 * torch.randn(co, generator=g)).numpy(),
                        (0.5 + torch.rand(co, generator=g)).numpy(), 1e-5))
            ci = co
        return out
    b1, b2 = rand_branch(), rand_branch()
    w3 = (torch.randn(96, 256, generator=g) * 0.05).numpy()
    b3 = torch.randn(96, generator=g).numpy()
    want, f1, f2 = oracle.positional_encoding(pts, b1, b2, w3, b3)
    P = pkg.positional_encoding
    got, pooled = P.positional_encoding(pts.to(dev), P.fold_branch(b1), P.fold_branch(b2), w3, b3, return_pooled=True)
    f = torch.cat([f1, f2], 1).transpose(1, 2)
    assert (pooled.cpu() - f).abs().max().item() < 2e-4 * f.abs().max().item()
    assert (got.cpu() - want).abs().max().item() < 2e-4 * want.abs().max().item()
    # masks taken before a rigid motion serve the moved cloud (the stage's use)
    q, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
    moved = ((pts - 0.1) @ q).contiguous()
    masks = pkg.pointnet2_utils.pe_masks(pts.to(dev))
    got2 = P.positional_encoding(moved.to(dev), P.fold_branch(b1), P.fold_branch(b2), w3, b3, masks=masks)
    want2, _, _ = oracle.positional_encoding(moved, b1, b2, w3, b3)
    assert (got2.cpu() - want2).abs().max().item() < 2e-4 * want2.abs().max().item()
