"""Times the tensor-core scoring kernel at the stage's shape (NM=... for other model sizes)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
L = ctypes.CDLL(pkg._lib.LIB_PATH)
B, H, n1, Nm = 16, 300, 196, int(os.environ.get("NM", 1024))
gen = torch.Generator().manual_seed(0)
q, _ = torch.linalg.qr(torch.randn(B, H, 3, 3, generator=gen))
R, t = q.contiguous().to(dev), (0.1 * torch.randn(B, H, 3, generator=gen)).to(dev)
pts, model = (0.4 * torch.randn(B, n1, 3, generator=gen)).to(dev), (0.4 * torch.randn(B, Nm, 3, generator=gen)).to(dev)
w1 = torch.zeros(B, n1)
w1[:, :152] = 1.0
w1 = w1.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for parts in (2, 2):
    for _ in range(3):
        pkg.model_utils.score_hypotheses(R, t, pts, w1, model)
    ms = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.model_utils.score_hypotheses(R, t, pts, w1, model); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print("parts", parts, "Nm", Nm, "us %.1f" % (sorted(ms)[len(ms) // 2] * 1e3))
