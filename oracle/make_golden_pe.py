"""oracle/make_golden_pe.py -- golden vectors for PositionalEncoding (model/fine_point_matching.py:102-144;
SURVEY.md section 8f, rank 1) from the LIVE reference module in this container.

TEST INFRASTRUCTURE ONLY.  Builds the reference's own PositionalEncoding with seeded random weights and
non-trivial BatchNorm running statistics, runs it in eval mode on a seeded cloud and records the weights,
the two pooled branches and the module's output into tests/golden/pe_mlp.npz.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_live  # noqa: E402


def branch_state(mlp):
    out = []
    for layer in mlp:                       # Conv2d(_ConvBase) = Sequential(conv, normlayer(bn), activation)
        conv, bn = layer[0], layer[1][0]    # BatchNorm2d(_BNBase) = Sequential(bn)
        out.append((conv.weight.detach().reshape(conv.weight.shape[0], -1).numpy(), bn.weight.detach().numpy(),
                    bn.bias.detach().numpy(), bn.running_mean.numpy(), bn.running_var.numpy(), bn.eps))
    return out


def main():
    ref_live.load()
    sys.path.insert(0, ref_live.REF_PEM)
    from model import fine_point_matching   # the reference module, unmodified
    torch.manual_seed(3)
    pe = fine_point_matching.PositionalEncoding(256)
    g = torch.Generator().manual_seed(4)
    for m in pe.modules():                  # trained-looking BatchNorm state
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
            m.running_mean = 0.1 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var = 0.5 + torch.rand(m.running_var.shape, generator=g)
    pe.eval()
    B, N = 2, 512
    pts = (torch.randn(B, N, 3, generator=g) * 0.25).contiguous()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):   # the reference prints shapes
        out = pe(pts)
        ctr = (pts + 0.00000001).contiguous()
        f1 = torch.amax(pe.mlp1(pe.group1(pts, ctr, pts.transpose(1, 2).contiguous())), dim=3)
        f2 = torch.amax(pe.mlp2(pe.group2(pts, ctr, pts.transpose(1, 2).contiguous())), dim=3)
    rec = {"pts": pts.numpy(), "out": out.numpy(), "f1": f1.numpy(), "f2": f2.numpy(),
           "w3": pe.mlp3[0].weight.detach().reshape(256, 256).numpy(), "b3": pe.mlp3[0].bias.detach().numpy()}
    for name, mlp in (("a", pe.mlp1), ("b", pe.mlp2)):
        for l, (w, gw, gb, rm, rv, eps) in enumerate(branch_state(mlp)):
            rec.update({f"{name}{l}_w": w, f"{name}{l}_g": gw, f"{name}{l}_b": gb, f"{name}{l}_m": rm,
                        f"{name}{l}_v": rv, f"{name}{l}_eps": np.float32(eps)})
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "pe_mlp.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, "out", out.shape, "f1", f1.shape)


if __name__ == "__main__":
    main()
