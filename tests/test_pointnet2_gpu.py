"""GPU parity: pointnet2 ops through the C ABI vs the CPU oracle -- BIT EXACT
(indices and gathered values; SURVEY.md section 8c, stage S0)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(B, N, seed, scale=0.4, dup=0.1, near_origin=3):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(B, N, 3, generator=g) * scale
    nd = int(dup * N)
    if nd:
        dst = torch.randperm(N, generator=g)[:nd]
        src = torch.randint(0, N, (nd,), generator=g)
        p[:, dst] = p[:, src]
    for k in range(min(near_origin, N)):
        p[:, (7 * k + 3) % N] = 0.01 * torch.randn(3, generator=g)  # |p|^2 <= 1e-3: skipped
    return p.contiguous()


@pytest.mark.parametrize("B,N,M", [
    (1, 1, 1), (2, 7, 7), (3, 100, 33), (2, 1000, 128), (4, 1024, 200), (3, 1025, 64),
    (2, 2048, 196), (2, 3000, 300), (2, 4096, 64), (1, 5000, 100), (1, 8192, 50),
    (2, 9000, 40), (1, 16384, 64), (3, 16000, 33), (1, 20000, 32), (2, 40000, 20), (1, 70000, 24),
    (1, 210000, 48), (1, 300000, 8), (2, 5, 9), (80, 9000, 6),
])
def test_fps_bit_exact(pkg, oracle, dev, B, N, M):
    p = _cloud(B, N, seed=N + M)
    want = oracle.fps(p.numpy(), M)
    got = pkg.pointnet2_ext.furthest_point_sampling(p.to(dev), M)
    assert got.dtype == torch.int32 and tuple(got.shape) == (B, M)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("B,N,M", [(1, 1, 1), (3, 100, 33), (4, 1024, 200), (32, 2048, 196), (2, 3000, 50), (1, 20000, 16),
                                   (2, 5, 9)])
def test_fps_gather_bit_exact(pkg, oracle, dev, B, N, M):
    """pem_fps_gather: the indices of pem_fps and the sampled points themselves (the point half of
    sample_pts_feats) -- from the sampling kernel where a cloud lives in one CTA, else FPS + row gather."""
    p = _cloud(B, N, seed=3 * N + M)
    g = p.to(dev)
    idx = torch.full((B, M), -1, dtype=torch.int32, device=dev)
    out = torch.full((B, M, 3), float("nan"), device=dev)
    L = pkg._lib.lib()
    nb = L.pem_fps_workspace_bytes(B, N)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
    pkg._lib.check(L.pem_fps_gather(g.data_ptr(), B, N, M, 1, idx.data_ptr(), out.data_ptr(), ws.data_ptr(), nb,
                                    torch.cuda.current_stream().cuda_stream), "pem_fps_gather")
    want = oracle.fps(p.numpy(), M)
    np.testing.assert_array_equal(idx.cpu().numpy(), want)
    np.testing.assert_array_equal(out.cpu().numpy(), np.take_along_axis(p.numpy(), want[..., None].astype(np.int64), 1))


def test_fps_cuda_variant_rule(pkg, oracle, dev):
    """flags=0: the reference CUDA kernel's rule (no near-origin skip), lowest-index ties."""
    p = _cloud(2, 1500, seed=5)
    want = oracle.fps(p.numpy(), 77, skip_near_origin=False)
    got = pkg.pointnet2_ext.furthest_point_sampling(p.to(dev), 77, skip_near_origin=False)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_fps_all_points_identical_or_skipped(pkg, oracle, dev):
    p = torch.full((1, 64, 3), 0.5)
    np.testing.assert_array_equal(
        pkg.pointnet2_ext.furthest_point_sampling(p.to(dev), 8).cpu().numpy(), oracle.fps(p.numpy(), 8))
    z = torch.zeros((1, 64, 3))  # every point skipped: the CPU loop keeps returning 0
    np.testing.assert_array_equal(
        pkg.pointnet2_ext.furthest_point_sampling(z.to(dev), 8).cpu().numpy(), oracle.fps(z.numpy(), 8))


def test_fps_default_instances(pkg, oracle, dev):
    b = pkg.synth.make_batch([3, 4])
    for key in ("dense_pm", "dense_po"):
        want = oracle.fps(b[key].numpy(), 196)
        got = pkg.pointnet2_ext.furthest_point_sampling(b[key].to(dev), 196)
        np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("B,C,N,M", [(2, 3, 2048, 196), (2, 256, 2049, 196), (16, 128, 256, 64), (1, 5, 17, 40)])
def test_gather_bit_exact(pkg, oracle, dev, B, C, N, M):
    g = torch.Generator().manual_seed(C)
    f = torch.randn(B, C, N, generator=g)
    idx = torch.randint(0, N, (B, M), generator=g, dtype=torch.int32)
    idx[0, 0] = -1  # invalid indices produce 0.0 (CPU rule)
    idx[-1, -1] = N
    want = oracle.gather(f.numpy(), idx.numpy())
    got = pkg.pointnet2_ext.gather_points(f.to(dev), idx.to(dev))
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # point-major row gather == transpose . gather . transpose
    rows = pkg.pointnet2_utils.gather_rows(f.transpose(1, 2).contiguous().to(dev), idx.to(dev))
    np.testing.assert_array_equal(rows.cpu().numpy(), want.transpose(0, 2, 1))


@pytest.mark.parametrize("B,M,N,r,ns", [
    (2, 2048, 2048, 0.1, 32), (2, 2048, 2048, 0.2, 64), (1, 1024, 256, 0.1, 64),
    (2, 100, 3000, 0.15, 16), (1, 130, 1025, 0.3, 1), (1, 5, 9, 0.5, 128),
])
def test_ball_query_bit_exact(pkg, oracle, dev, B, M, N, r, ns):
    xyz = _cloud(B, N, seed=N + ns, scale=0.35, near_origin=0)
    if M == N:
        centres = (xyz + 1e-8).contiguous()  # fine_point_matching.py:117
    else:
        centres = _cloud(B, M, seed=M, scale=0.35, near_origin=0)
    centres[:, -1] = 50.0  # a centre with no neighbour: row of zeros
    want = oracle.ball_query(centres.numpy(), xyz.numpy(), r, ns)
    got = pkg.pointnet2_ext.ball_query(centres.to(dev), xyz.to(dev), r, ns)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert (want[:, -1] == 0).all()


@pytest.mark.parametrize("B,C,N,M,ns", [(2, 3, 2048, 2048, 32), (1, 256, 512, 100, 64), (7, 3, 2048, 64, 32)])
def test_group_bit_exact(pkg, oracle, dev, B, C, N, M, ns):
    g = torch.Generator().manual_seed(ns + C)
    f = torch.randn(B, C, N, generator=g)
    idx = torch.randint(0, N, (B, M, ns), generator=g, dtype=torch.int32)
    idx[0, 0, 0] = -3
    want = oracle.group(f.numpy(), idx.numpy())
    got = pkg.pointnet2_ext.group_points(f.to(dev), idx.to(dev))
    np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("r,ns", [(0.1, 32), (0.2, 64)])
def test_query_and_group_fused_bit_exact(pkg, oracle, dev, r, ns):
    b = pkg.synth.make_batch([5])
    xyz = b["dense_po"]
    new_xyz = (xyz + 1e-8).contiguous()
    feat = xyz.transpose(1, 2).contiguous()
    want, widx = oracle.query_and_group(xyz.numpy(), new_xyz.numpy(), feat.numpy(), r, ns)
    got, gidx = pkg.pointnet2_utils.query_and_group(xyz.to(dev), new_xyz.to(dev), feat.to(dev), r, ns,
                                                    return_idx=True)
    np.testing.assert_array_equal(gidx.cpu().numpy(), widx)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # the unfused composition (module with fused=False) gives the same tensor
    mod = pkg.pointnet2_utils.QueryAndGroup(r, ns, fused=False)
    np.testing.assert_array_equal(mod(xyz.to(dev), new_xyz.to(dev), feat.to(dev)).cpu().numpy(), want)


def test_sample_pts_feats(pkg, oracle, dev):
    b = pkg.synth.make_batch([6])
    p, f, idx = pkg.model_utils.sample_pts_feats(b["dense_pm"].to(dev), b["dense_fm"].to(dev), 196, True)
    op, of, oidx = oracle.sample_pts_feats(b["dense_pm"].numpy(), b["dense_fm"].numpy(), 196)
    np.testing.assert_array_equal(idx.cpu().numpy(), oidx)
    np.testing.assert_array_equal(p.cpu().numpy(), op)
    np.testing.assert_array_equal(f.cpu().numpy(), of)


def test_error_behaviour_matches_reference_checks(pkg, dev):
    ext = pkg.pointnet2_ext
    p = torch.randn(2, 64, 3, device=dev)
    with pytest.raises(RuntimeError, match="contiguous"):
        ext.furthest_point_sampling(p.transpose(1, 2), 4)
    with pytest.raises(RuntimeError, match="float"):
        ext.furthest_point_sampling(p.double(), 4)
    with pytest.raises(RuntimeError, match="int"):
        ext.gather_points(p.transpose(1, 2).contiguous(), torch.zeros(2, 4, dtype=torch.int64, device=dev))
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.furthest_point_sampling(p.cpu(), 4)
    with pytest.raises(NotImplementedError):
        ext.three_nn(p, p)


def test_pe_group2_fused_bit_exact(pkg, oracle, dev):
    """Both PositionalEncoding grouping scales in one kernel == two QueryAndGroup calls."""
    b = pkg.synth.make_batch([7, 8])
    for key in ("dense_po", "dense_pm"):
        xyz = b[key] - (b["gt_t"][:, None, :] if key == "dense_pm" else 0.0)
        xyz = xyz.contiguous()
        new_xyz = (xyz + 1e-8).contiguous()
        feat = xyz.transpose(1, 2).contiguous()
        o1, o2 = pkg.pointnet2_utils.pe_group2(xyz.to(dev))
        w1, _ = oracle.query_and_group(xyz.numpy(), new_xyz.numpy(), feat.numpy(), 0.1, 32)
        w2, _ = oracle.query_and_group(xyz.numpy(), new_xyz.numpy(), feat.numpy(), 0.2, 64)
        np.testing.assert_array_equal(o1.cpu().numpy(), w1)
        np.testing.assert_array_equal(o2.cpu().numpy(), w2)
    # ragged size, tiny nsample, radius large enough to fill every row early
    g = torch.Generator().manual_seed(2)
    xyz = (torch.randn(1, 777, 3, generator=g) * 0.2).contiguous()
    o1, o2 = pkg.pointnet2_utils.pe_group2(xyz.to(dev), 0.05, 3, 0.6, 128)
    feat = xyz.transpose(1, 2).contiguous()
    w1, _ = oracle.query_and_group(xyz.numpy(), (xyz + 1e-8).numpy(), feat.numpy(), 0.05, 3)
    w2, _ = oracle.query_and_group(xyz.numpy(), (xyz + 1e-8).numpy(), feat.numpy(), 0.6, 128)
    np.testing.assert_array_equal(o1.cpu().numpy(), w1)
    np.testing.assert_array_equal(o2.cpu().numpy(), w2)
    # cloud larger than the shared-memory staging limit (tiled path with global gathers)
    xyz = (torch.randn(1, 5000, 3, generator=g) * 0.3).contiguous()
    o1, o2 = pkg.pointnet2_utils.pe_group2(xyz.to(dev), 0.1, 32, 0.2, 64)
    feat = xyz.transpose(1, 2).contiguous()
    w1, _ = oracle.query_and_group(xyz.numpy(), (xyz + 1e-8).numpy(), feat.numpy(), 0.1, 32)
    w2, _ = oracle.query_and_group(xyz.numpy(), (xyz + 1e-8).numpy(), feat.numpy(), 0.2, 64)
    np.testing.assert_array_equal(o1.cpu().numpy(), w1)
    np.testing.assert_array_equal(o2.cpu().numpy(), w2)


@pytest.mark.parametrize("N,scale,moved", [(2048, 0.3, True), (777, 0.2, True), (2048, 0.05, True), (4096, 0.3, True),
                                           (2048, 0.3, False), (1500, 3.0, False), (33, 0.1, True)])
def test_pe_group2_masked_bit_exact(pkg, oracle, dev, N, scale, moved):
    """Tensor-core candidate masks (from the UNMOVED cloud when `moved`) + exact re-test of the
    candidates == QueryAndGroup on the final cloud, bit for bit: random rigid motions, a dense cloud
    where every neighbour list fills up at once, a sparse one with mostly empty balls, points
    planted exactly on both ball surfaces, sizes that are not multiples of the tiles."""
    g = torch.Generator().manual_seed(N)
    B = 3
    xyz = (torch.randn(B, N, 3, generator=g) * scale + torch.tensor([0.7, -0.2, 1.3])).contiguous()
    xyz[:, 1] = xyz[:, 0] + torch.tensor([0.1, 0.0, 0.0])   # |d| == r1 up to rounding
    xyz[:, 2] = xyz[:, 0] + torch.tensor([0.0, 0.2, 0.0])   # |d| == r2 up to rounding
    final = xyz
    if moved:
        q, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
        t = 0.3 * torch.randn(B, 1, 3, generator=g)
        final = ((xyz - t) @ q).contiguous()                  # MU:267 form of the rigid motion
    masks = pkg.pointnet2_utils.pe_masks(xyz.to(dev))
    o1, o2 = pkg.pointnet2_utils.pe_group2_masked(final.to(dev), masks)
    feat = final.transpose(1, 2).contiguous()
    w1, _ = oracle.query_and_group(final.numpy(), (final + 1e-8).numpy(), feat.numpy(), 0.1, 32)
    w2, _ = oracle.query_and_group(final.numpy(), (final + 1e-8).numpy(), feat.numpy(), 0.2, 64)
    np.testing.assert_array_equal(o1.cpu().numpy(), w1)
    np.testing.assert_array_equal(o2.cpu().numpy(), w2)
    # the filter is a filter: a few per cent of the pairs, never fewer than the true neighbours
    bits = masks.view(torch.int32).view(B, N, -1)
    frac = sum(bin(int(x) & 0xFFFFFFFF).count("1") for x in bits[0].flatten().tolist()) / float(N * N)
    true_frac = float((torch.cdist(final[0], final[0]) < 0.2).float().mean())
    assert true_frac <= frac <= true_frac * 1.05 + 1e-3, (frac, true_frac)
