"""oracle/cpu_stage.py -- the PEM geometric stage on the HOST, in the reference's own
arithmetic.  TEST INFRASTRUCTURE / CPU BASELINE ONLY (bench.py `cpu_baseline` and
`--impl reference`, tests/, smoke()).

Same step sequence as openvino-sam-6d_b200/stage.py, executed with
  * the UNMODIFIED reference pointnet2 CPU extension (oracle/_ref, built from
    /root/reference by oracle/build_ref.py) called the way the reference's wrappers call it
    (pointnet2_utils.py:78,118,255,290; QueryAndGroup :343-384), falling back to the C
    restatement only if that build is absent; and
  * the torch-CPU restatement of utils/model_utils.py in pem_oracle.py (bit-identical to the
    live reference, see tests/test_oracle_cpu.py), including -- with faithful=True -- the
    reference's (B, 3*P1, N1*N2) compare/argmax sampler (MU:277-305), which is what its CPU
    path really executes.
"""
import os

import numpy as np
import torch

import build_ref
import pem_oracle as O

PE_SCALES = ((0.1, 32), (0.2, 64))


class _RefOps:
    """pointnet2 ops through the reference's compiled CPU module."""

    def __init__(self):
        self.ext = build_ref.load_ref_ext()
        self.kind = "reference"

    def fps(self, xyz, m):
        return self.ext.furthest_point_sampling(xyz.contiguous(), m)

    def gather(self, feat, idx):
        return self.ext.gather_points(feat.contiguous(), idx)

    def query_and_group(self, xyz, new_xyz, feat, r, ns):
        idx = self.ext.ball_query(new_xyz.contiguous(), xyz.contiguous(), r, ns)
        g = self.ext.group_points(xyz.transpose(1, 2).contiguous(), idx)
        g -= new_xyz.transpose(1, 2).unsqueeze(-1)
        return torch.cat([g, self.ext.group_points(feat.contiguous(), idx)], dim=1)


class _PortOps:
    """Same, through the C restatement (used only when oracle/_ref is unavailable)."""
    kind = "port"

    def fps(self, xyz, m):
        return torch.from_numpy(O.fps(xyz.numpy(), m))

    def gather(self, feat, idx):
        return torch.from_numpy(O.gather(feat.contiguous().numpy(), idx.numpy()))

    def query_and_group(self, xyz, new_xyz, feat, r, ns):
        return torch.from_numpy(O.query_and_group(xyz.numpy(), new_xyz.numpy(),
                                                  feat.contiguous().numpy(), r, ns)[0])


def pointnet2_ops():
    if os.path.exists(build_ref.so_path()) or os.path.isdir(build_ref.EXT_SRC):
        try:
            return _RefOps()
        except Exception:  # pragma: no cover - missing libtorch symbols etc.
            pass
    return _PortOps()


def run_stage_cpu(batch, n_sparse=196, n_proposal1=6000, n_proposal2=300, temp=0.1,
                  faithful=True, ops=None, chunk=2):
    """batch: dict from synth.make_batch (CPU tensors).  Returns R, t, score, stages."""
    ops = ops or pointnet2_ops()
    pm, fm, po, fo = batch["dense_pm"], batch["dense_fm"], batch["dense_po"], batch["dense_fo"]
    B = pm.shape[0]
    bg = batch["bg"][:, None, :]

    def sample(p, f):  # sample_pts_feats, MU:70-84
        idx = ops.fps(p, n_sparse)
        sp = ops.gather(p.transpose(1, 2).contiguous(), idx).transpose(1, 2)
        sf = ops.gather(f.transpose(1, 2).contiguous(), idx).transpose(1, 2)
        return sp.contiguous(), sf.contiguous(), idx

    sp, sf, idx_m = sample(pm, fm)
    tp, tf, idx_o = sample(po, fo)
    atten_c = O.feature_similarity(torch.cat([bg, sf], 1), torch.cat([bg, tf], 1), temp)
    R0s, t0s, pids, tops, scs, posed, w1s = [], [], [], [], [], [], []
    for s in range(0, B, chunk):  # ~3.5 GB of temporaries per instance in faithful mode
        e = slice(s, s + chunk)
        R0, t0, st = O.coarse_Rt(atten_c[e], sp[e], tp[e], batch["model"][e], n_proposal1,
                                 n_proposal2, batch["rand"][e], faithful=faithful)
        R0s.append(R0); t0s.append(t0); pids.append(st["proposal_id"])
        # for the parity tests: the scored hypotheses, and which of them are well posed (sigma2/sigma1 of
        # their 3x3 H above 1e-2; the rest are the degenerate triplets whose rotation LAPACK takes from noise)
        Htop = torch.gather(st["H"], 1, st["top"].reshape(st["top"].shape[0], -1, 1, 1).expand(-1, -1, 3, 3))
        sv = torch.linalg.svdvals(Htop.double())
        tops.append(st["top"]); scs.append(st["scores"]); w1s.append(st["w1"])
        posed.append(sv[..., 1] / sv[..., 0].clamp_min(1e-30) > 1e-2)
    R0, t0, pid = torch.cat(R0s), torch.cat(t0s), torch.cat(pids)
    p1_ = (pm - t0.unsqueeze(1)) @ R0
    pe = []
    for cloud in (p1_, po):
        for r, ns in PE_SCALES:
            pe.append(ops.query_and_group(cloud.contiguous(), (cloud + 1e-8).contiguous(),
                                          cloud.transpose(1, 2).contiguous(), r, ns))
    ff1, ff2 = torch.cat([bg, fm], 1), torch.cat([bg, fo], 1)
    fine_sparse = []
    for _ in range(3):
        fine_sparse.append(ops.gather(ff1.transpose(1, 2).contiguous(), idx_m).transpose(1, 2).contiguous())
        fine_sparse.append(ops.gather(ff2.transpose(1, 2).contiguous(), idx_o).transpose(1, 2).contiguous())
    atten_f = O.feature_similarity(ff1, ff2, temp)
    R, t, score, _ = O.fine_Rt(atten_f, pm, po, batch["model"])
    return R, t, score, {"init_R": R0, "init_t": t0, "proposal_id": pid, "coarse_top": torch.cat(tops),
                         "coarse_scores": torch.cat(scs), "coarse_posed": torch.cat(posed),
                         "coarse_w1": torch.cat(w1s), "sparse_pm": sp, "idx_m": idx_m,
                         "idx_o": idx_o, "pe": pe, "fine_sparse": fine_sparse, "ops": ops.kind}
