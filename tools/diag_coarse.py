import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package(); O = g.load_oracle()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_dense_gpu import _coarse_inputs
dev = torch.device("cuda:0")
ids = [20, 21, 22, 23]
b, sp, tp, atten = _coarse_inputs(pkg, O, ids)
oR, ot, ost = O.coarse_Rt(atten, sp, tp, b["model"], 6000, 300, b["rand"])
R, t, st = pkg.model_utils.compute_coarse_Rt(atten.to(dev), sp.to(dev), tp.to(dev), b["model"].to(dev), 6000, 300, rand=b["rand"].to(dev), return_stages=True)
st = {k: v.cpu() for k, v in st.items()}
cd = (st["cdf"] - ost["cdf"])
print("cdf max abs diff", cd.abs().max().item(), "mean abs", cd.abs().mean().item())
print("cdf equal frac", (st["cdf"] == ost["cdf"]).float().mean().item())
mis = st["corr"].long() != ost["idx"]
print("corr mismatch frac per inst", mis.float().mean(1))
for i in range(4):
    m = torch.nonzero(mis[i]).flatten()[:5]
    for d in m.tolist():
        r = b["rand"][i, d].item(); a, c = st["corr"][i, d].item(), ost["idx"][i, d].item()
        print(" inst", i, "draw", d, "r", r, "ours", a, "oracle", c, "our cdf[a]", st["cdf"][i, a].item(), "orc cdf[c]", ost["cdf"][i, c].item(), "orc cdf[c-1]", ost["cdf"][i, c-1].item())
    pid_o = ost["proposal_id"][i].item(); pid = st["proposal_id"][i].item()
    print("inst", i, "oracle winner", pid_o, "ours", pid, "corr equal for oracle winner:", torch.equal(st["corr"][i, 3*pid_o:3*pid_o+3].long(), ost["idx"][i, 3*pid_o:3*pid_o+3]))
    print("   in our sel:", (st["sel"][i] == pid_o).any().item(), "resid ours/orc", st["resid"][i, pid_o].item(), ost["resid"][i, pid_o].item())
    if (st["sel"][i] == pid_o).any():
        pos = torch.nonzero(st["sel"][i] == pid_o).flatten()[0]
        print("   our score for it", st["scores"][i, pos].item(), "oracle score", ost["scores"][i].max().item(), "our best", st["scores"][i].max().item())
    dR = (st["R"][i, pid_o] .reshape(3,3)- ost["R"][i, pid_o]).abs().max().item()
    print("   dR for that hyp", dR, "w1 equal", torch.equal(st["w1"][i], ost["w1"][i]))
    # score comparison across common selection
    common = sorted(set(st["sel"][i].tolist()) & set(ost["top"][i].tolist()))
    print("   common sel", len(common))
    so = {int(k): float(v) for k, v in zip(ost["top"][i], ost["scores"][i])}
    sg = {int(k): float(v) for k, v in zip(st["sel"][i], st["scores"][i])}
    rel = max(abs(so[k]-sg[k])/so[k] for k in common)
    worst = max(common, key=lambda k: abs(so[k]-sg[k])/so[k])
    print("   max rel score diff over common", rel, "at", worst, so[worst], sg[worst], "dR", (st["R"][i, worst].reshape(3,3)-ost["R"][i, worst]).abs().max().item())
