"""Synthetic PEM instances of the reference's default shapes.

One *instance* is what the reference's ``Net.forward`` hands to its geometric
stage for one detected object (model/pose_estimation_model.py:29-55):

  dense_pm (N,3)   observed cloud, camera frame, divided by the object radius
  dense_fm (N,C)   per-point features of the observed cloud
  dense_po (N,3)   template cloud, object frame, radius-normalised (max norm 1)
  dense_fo (N,C)   per-point features of the template cloud
  model    (Nm,3)  CAD surface sample / (radius + 1e-6)
  bg       (C,)    background token
  rand     (3*P1,) the uniform draw ``compute_coarse_Rt`` consumes (MU:293)
  gt_R, gt_t       planted pose: observed = R @ template + t

The ViT and the geometric transformers are outside the hot path; their place is
taken by planted features: every observed point carries the feature of its
nearest template point plus unit Gaussian noise, so the pose is recoverable and
the result can be checked against the plant as well as against the oracle.
Observed clouds contain 10 % exact duplicates, like the reference's
sample-with-replacement of small masks (run_inference_custom_pytorch.py:337-338).

Generated with torch's CPU generator, one seed per instance (= instance id), so
every rank of a multi-GPU run can build its own slice independently.
"""
import math

import torch


def _random_rotation(g):
    q = torch.randn(4, generator=g, dtype=torch.float64)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    return torch.tensor([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=torch.float64)


def _ellipsoid_surface(n, axes, g):
    """Roughly area-uniform samples on an ellipsoid plus outward normals (fp64)."""
    u = torch.randn(n, 3, generator=g, dtype=torch.float64)
    u = u / u.norm(dim=1, keepdim=True)
    p = u * axes
    nrm = p / (axes * axes)
    return p, nrm / nrm.norm(dim=1, keepdim=True)


def make_instance(seed, n_dense=2048, n_model=1024, feat_dim=256, n_proposal1=6000,
                  noise=0.003, dup_frac=0.10, feat_noise=1.0, clutter_frac=0.12,
                  tmpl_bg_frac=0.05):
    g = torch.Generator().manual_seed(int(seed))
    axes = torch.tensor([1.0, 0.55 + 0.4 * torch.rand(1, generator=g).item(),
                         0.35 + 0.4 * torch.rand(1, generator=g).item()], dtype=torch.float64)
    # a bump breaks the ellipsoid's symmetries so the pose is unique
    tmpl, _ = _ellipsoid_surface(n_dense, axes, g)
    bump = torch.tensor([0.6, 0.5, 0.3], dtype=torch.float64)

    def deform(p):
        return p * (1.0 + 0.35 * torch.exp(-((p - bump * axes) ** 2).sum(1, keepdim=True) / 0.08))

    tmpl = deform(tmpl)
    model, _ = _ellipsoid_surface(n_model, axes, g)
    model = deform(model)
    radius = tmpl.norm(dim=1).max()
    tmpl = tmpl / radius
    model = model / (radius + 1e-6)

    R = _random_rotation(g)
    t = torch.tensor([2.0 * torch.rand(1, generator=g).item() - 1.0,
                      2.0 * torch.rand(1, generator=g).item() - 1.0,
                      4.0 + 8.0 * torch.rand(1, generator=g).item()], dtype=torch.float64)

    # observed: surface points whose normal faces the camera (camera at origin, looking +z)
    cand, cn = _ellipsoid_surface(8 * n_dense, axes, g)
    cand = deform(cand) / radius
    cam = cand @ R.T + t
    facing = ((cn @ R.T) * (-cam / cam.norm(dim=1, keepdim=True))).sum(1) > 0.15
    vis = cand[facing]
    sel = torch.randperm(vis.shape[0], generator=g)[:n_dense]
    if sel.numel() < n_dense:  # pad by resampling with replacement
        extra = torch.randint(0, vis.shape[0], (n_dense - sel.numel(),), generator=g)
        sel = torch.cat([sel, extra])
    obs_obj = vis[sel]
    obs = obs_obj @ R.T + t + noise * torch.randn(n_dense, 3, generator=g, dtype=torch.float64)

    f_tmpl = torch.randn(n_dense, feat_dim, generator=g)
    nn = torch.cdist(obs_obj.float(), tmpl.float()).argmin(1)
    f_obs = f_tmpl[nn] + feat_noise * torch.randn(n_dense, feat_dim, generator=g)
    f_obs = f_obs / f_obs.norm(dim=1, keepdim=True)
    f_tmpl = f_tmpl / f_tmpl.norm(dim=1, keepdim=True)

    bg = torch.randn(feat_dim, generator=g)
    bg = bg / bg.norm()
    # clutter: observed points just off the surface whose features look like background, and a
    # few template points with background-like features, so both foreground masks
    # (MU:231-234) have zeros in them
    ncl = int(clutter_frac * n_dense)
    if ncl > 0:
        cl = torch.randperm(n_dense, generator=g)[:ncl]
        obs[cl] = obs[cl] + 0.03 * torch.randn(ncl, 3, generator=g, dtype=torch.float64)
        fcl = bg * math.sqrt(feat_dim) + feat_noise * torch.randn(ncl, feat_dim, generator=g)
        f_obs[cl] = fcl / fcl.norm(dim=1, keepdim=True)
    nbg = int(tmpl_bg_frac * n_dense)
    if nbg > 0:
        tb = torch.randperm(n_dense, generator=g)[:nbg]
        fb = bg * math.sqrt(feat_dim) + feat_noise * torch.randn(nbg, feat_dim, generator=g)
        f_tmpl[tb] = fb / fb.norm(dim=1, keepdim=True)

    ndup = int(dup_frac * n_dense)
    if ndup > 0:
        dst = torch.randperm(n_dense, generator=g)[:ndup]
        src = torch.randint(0, n_dense, (ndup,), generator=g)
        obs[dst] = obs[src]
        f_obs[dst] = f_obs[src]

    rand = torch.rand(3 * n_proposal1, generator=g)
    return {
        "dense_pm": obs.float(), "dense_fm": f_obs.float(),
        "dense_po": tmpl.float(), "dense_fo": f_tmpl.float(),
        "model": model.float(), "bg": bg.float(), "rand": rand,
        "gt_R": R.float(), "gt_t": t.float(),
    }


def make_batch(instance_ids, **kw):
    """Stack instances; every value gets a leading batch dimension."""
    items = [make_instance(i, **kw) for i in instance_ids]
    return {k: torch.stack([it[k] for it in items]).contiguous() for k in items[0]}


def replicate_batch(batch, reps):
    """Tile a small generated batch to a larger one (bench only: generation is CPU-bound)."""
    return {k: v.repeat(reps, *([1] * (v.dim() - 1))).contiguous() for k, v in batch.items()}


def nominal_flops_per_instance(n_dense=2048, n_sparse=196, feat_dim=256):
    """Algorithmic similarity flops: 2*(M+1)^2*C coarse + fine."""
    return 2 * (n_sparse + 1) ** 2 * feat_dim + 2 * (n_dense + 1) ** 2 * feat_dim

