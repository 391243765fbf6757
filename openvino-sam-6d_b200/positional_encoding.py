"""Mirror of the reference's PositionalEncoding (model/fine_point_matching.py:102-144; SURVEY.md
section 8f, rank 1): both QueryAndGroup scales, the two SharedMLP([6,32,64,128]) + amax branches and
the final Conv1d(256, out_dim), with nothing between the cloud and the (B,N,256) pooled features
written to memory (the reference materialises (B,6,N,ns) groups and (B,128,N,ns) activations)."""
import torch

from . import _lib
from . import pointnet2_ext as _ext
from . import pointnet2_utils as _pu


def fold_branch(layers):
    """layers: three (conv_weight (Co,Ci[,1,1]), bn_weight, bn_bias, running_mean, running_var, eps).
    Returns (w1, w2, w3, shift): BatchNorm in eval mode, y = (x - mean) / sqrt(var + eps) * g + b, folded
    into the layer -- its scale g / sqrt(var + eps) into the rows of the weight matrix, the rest into one
    shift per channel (pytorch_utils.py:80-134: the convolution has no bias when bn)."""
    ws, sh = [], []
    for (w, g, b, m, v, eps) in layers:
        w = torch.as_tensor(w).float()
        g, b, m, v = (torch.as_tensor(z).float().to(w.device) for z in (g, b, m, v))
        s = g / torch.sqrt(v + float(eps))
        ws.append((w.reshape(w.shape[0], -1) * s[:, None]).contiguous())
        sh.append(b - m * s)
    return ws[0], ws[1], ws[2], torch.cat(sh).contiguous()


def branch_from_module(shared_mlp):
    """The same from a reference SharedMLP module (Sequential of Conv2d(_ConvBase) blocks)."""
    layers = []
    for block in shared_mlp:
        conv, bn = block[0], block[1][0]
        layers.append((conv.weight.detach(), bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                       bn.running_var, bn.eps))
    return fold_branch(layers)


def pe_neighbours(xyz, masks=None, radius1=0.1, nsample1=32, radius2=0.2, nsample2=64, centre_eps=1e-8):
    """Neighbour lists of both PE scales, (B,N,ns) int32 each == ball_query(r, ns, xyz, xyz + eps)."""
    _ext._check(xyz, "xyz", torch.float32, "a float")
    B, N, _ = xyz.shape
    if masks is None:
        masks = _pu.pe_masks(xyz, max(radius1, radius2), centre_eps)
    i1 = torch.empty((B, N, nsample1), dtype=torch.int32, device=xyz.device)
    i2 = torch.empty((B, N, nsample2), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().pem_pe_neighbours_masked(xyz.data_ptr(), B, N, float(centre_eps), float(radius1),
                                                       int(nsample1), i1.data_ptr(), float(radius2),
                                                       int(nsample2), i2.data_ptr(), masks.data_ptr(),
                                                       masks.numel(), _ext._stream()), "pem_pe_neighbours_masked")
    return i1, i2


def pe_branch(xyz, idx, branch, out, ch_off, centre_eps=1e-8):
    """One branch into channels [ch_off, ch_off+128) of out (B,N,ld): group -> SharedMLP -> amax."""
    w1, w2, w3, shift = (t.to(xyz.device).contiguous() for t in branch)
    B, N, _ = xyz.shape
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().pem_pe_mlp(xyz.data_ptr(), idx.data_ptr(), B, N, idx.shape[2], float(centre_eps),
                                         w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), shift.data_ptr(),
                                         out.data_ptr(), out.shape[2], int(ch_off),
                                         _ext._stream()), "pem_pe_mlp")
    return out


def positional_encoding(pts, branch1, branch2, w3, b3, masks=None, r1=0.1, ns1=32, r2=0.2, ns2=64,
                        return_pooled=False):
    """PositionalEncoding.forward(pts) (fine_point_matching.py:113-144): (B,N,3) -> (B,N,out_dim).
    `masks` may come from pe_masks on the cloud BEFORE a rigid motion (see pe_group2_masked)."""
    xyz = pts.contiguous()
    i1, i2 = pe_neighbours(xyz, masks, r1, ns1, r2, ns2)
    pooled = torch.empty((xyz.shape[0], xyz.shape[1], 256), dtype=torch.float32, device=xyz.device)
    pe_branch(xyz, i1, branch1, pooled, 0)
    pe_branch(xyz, i2, branch2, pooled, 128)
    w3 = torch.as_tensor(w3).float().to(xyz.device)
    out = torch.nn.functional.linear(pooled, w3.reshape(w3.shape[0], -1), torch.as_tensor(b3).float().to(xyz.device))
    return (out, pooled) if return_pooled else out


def patch_positional_encoding(fine_point_matching_module):
    """Replace PositionalEncoding.forward of the reference's (imported) module; weights are read from the
    module's own mlp1 / mlp2 / mlp3 at call time (eval mode)."""
    cls = fine_point_matching_module.PositionalEncoding

    @torch.no_grad()
    def forward(self, pts1, pts2=None):
        conv3 = self.mlp3[0]
        return positional_encoding(pts1, branch_from_module(self.mlp1), branch_from_module(self.mlp2),
                                   conv3.weight.detach(), conv3.bias.detach())

    cls.forward = forward
    return cls
