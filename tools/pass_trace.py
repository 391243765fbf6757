# NOTE: needs a -DPEM_DEV_KNOBS build (PEM_B200_LIB=scratch_ab/libpem_dev.so, see csrc/Makefile).
"""Device-side timeline of CTA 0 of one fine GEMM pass (PEM_TRACE in fine_tc.cu): what each role waits for, step by step.
roles: 0 producer, 1 MMA issuer, 2 first epilogue warp, 3 last epilogue warp.  Times in clocks since the first event."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
L = ctypes.CDLL(pkg._lib.LIB_PATH)
L.pem_dev_fine_pass.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
L.pem_dev_set_fine_trace.argtypes = [ctypes.c_void_p]
lib = pkg._lib.lib()
B = int(os.environ.get("B", 16))
base = pkg.synth.make_batch(list(range(4)))
b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, B // 4).items()}
bg = b["bg"][:, None, :]
f1, f2 = torch.cat([bg, b["dense_fm"]], 1).contiguous(), torch.cat([bg, b["dense_fo"]], 1).contiguous()
N, C = f1.shape[1] - 1, f1.shape[2]
nb = lib.pem_fine_fused_workspace_bytes(B, N, N, C)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
R = torch.empty(B, 9, device=dev); t = torch.empty(B, 3, device=dev); sc = torch.empty(B, device=dev)
st = torch.cuda.current_stream().cuda_stream
rc = lib.pem_fine_fused_ex(f1.data_ptr(), f2.data_ptr(), b["dense_pm"].data_ptr(), b["dense_po"].data_ptr(), b["model"].data_ptr(),
                           B, N, N, C, b["model"].shape[1], 0.1, 0.15, 1, R.data_ptr(), t.data_ptr(), sc.data_ptr(),
                           None, None, None, None, ws.data_ptr(), nb, st)
assert rc == 0, rc
torch.cuda.synchronize()
TL = 512
NAMES = {1: "task", 2: "step", 10: "wait acc", 11: "got acc", 12: "got full0", 13: "got full1", 14: "got full2", 15: "got full3",
         30: "wait full", 31: "got full", 32: "read done", 40: "task", 41: "steps done", 42: "task end"}
for mode in [int(x) for x in os.environ.get("MODES", "1,4").split(",")]:
    for flags in [int(x) for x in os.environ.get("FLAGS", "0").split(",")]:
        def run():
            rc = L.pem_dev_fine_pass(f1.data_ptr(), f2.data_ptr(), B, N, N, C, 0.1, flags | 8, mode, 0, ws.data_ptr(), nb, st)
            assert rc == 0, rc
        for _ in range(5): run()
        torch.cuda.synchronize()
        tr_all = torch.zeros(4 * TL * 2 + 6 * 160, dtype=torch.int64, device=dev)
        L.pem_dev_set_fine_trace(tr_all.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        L.pem_dev_set_fine_trace(None)
        tr = tr_all[:4 * TL * 2].view(4, TL, 2).cpu()
        span = tr_all[4 * TL * 2:].view(160, 6).cpu()
        span = span[span[:, 0] > 0]
        g0 = int(span[:, 1].min())
        print(f"==== mode {mode} flags {flags}: event time {e0.elapsed_time(e1) * 1e3:.1f} us; per CTA (globaltimer ns since the first CTA start): "
              f"start {int(span[:, 1].max()) - g0} max, dependency wait over {int(span[:, 3].min()) - g0}..{int(span[:, 3].max()) - g0}, "
              f"end {int(span[:, 5].min()) - g0}..{int(span[:, 5].max()) - g0}; body clocks {int((span[:, 4] - span[:, 2]).min())}..{int((span[:, 4] - span[:, 2]).max())}")
        body = (span[:, 4] - span[:, 2]).tolist(); ends = (span[:, 5] - g0).tolist()
        print("   body clocks per CTA:", " ".join(str(int(x)) for x in body))
        print("   end ns per CTA:", " ".join(str(int(x)) for x in ends))
        t0 = int(tr[:, :, 1][tr[:, :, 0] > 0].min())
        print(f"==== mode {mode} flags {flags} B {B}")
        for role in range(4):
            ev = [(int(a), int(c) - t0) for a, c in tr[role].tolist() if a > 0]
            print(f"-- role {role} ({len(ev)} events)")
            line, prev = [], None
            for tag, c in ev:
                line.append(f"{NAMES.get(tag, tag)}@{c}" + (f"(+{c - prev})" if prev is not None else ""))
                prev = c
                if tag in (2, 15, 32, 42, 1):
                    print("   " + "  ".join(line)); line = []
            if line: print("   " + "  ".join(line))
