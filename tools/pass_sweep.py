# NOTE: uses pem_dev_* tuning knobs: needs a -DPEM_DEV_KNOBS build (PEM_B200_LIB=scratch_ab/libpem_dev.so, see csrc/Makefile).
"""Fixed versus per-step cost of one fine GEMM pass: the pass timed for several instance counts (= rounds of tasks per CTA pair)
and steps-per-task settings, with roles switched off (flags: 1 no MMAs, 2 no streamed loads, 4 no epilogue math)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
L = ctypes.CDLL(pkg._lib.LIB_PATH)
L.pem_dev_fine_pass.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
lib = pkg._lib.lib()
base = pkg.synth.make_batch(list(range(4)))
SINGLE = int(os.environ.get("SINGLE", 1))
st = torch.cuda.current_stream().cuda_stream
for B in [int(x) for x in os.environ.get("BS", "4,8,16,24,36").split(",")]:
    b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, B // 4).items()}
    bg = b["bg"][:, None, :]
    f1, f2 = torch.cat([bg, b["dense_fm"]], 1).contiguous(), torch.cat([bg, b["dense_fo"]], 1).contiguous()
    N, C = f1.shape[1] - 1, f1.shape[2]
    nb = lib.pem_fine_fused_workspace_bytes(B, N, N, C)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    R = torch.empty(B, 9, device=dev); t = torch.empty(B, 3, device=dev); sc = torch.empty(B, device=dev)
    rc = lib.pem_fine_fused_ex(f1.data_ptr(), f2.data_ptr(), b["dense_pm"].data_ptr(), b["dense_po"].data_ptr(), b["model"].data_ptr(),
                               B, N, N, C, b["model"].shape[1], 0.1, 0.15, SINGLE, R.data_ptr(), t.data_ptr(), sc.data_ptr(),
                               None, None, None, None, ws.data_ptr(), nb, st)
    assert rc == 0, rc
    torch.cuda.synchronize()
    for spt in [int(x) for x in os.environ.get("SPT", "8,4").split(",")]:
        L.pem_dev_set_fine_steps_per_task(spt)
        for mode in [int(x) for x in os.environ.get("MODES", "1,4").split(",")]:
            row = []
            for flags in [int(x) for x in os.environ.get("FLAGS", "0,7").split(",")]:
                def run():
                    rc = L.pem_dev_fine_pass(f1.data_ptr(), f2.data_ptr(), B, N, N, C, 0.1, flags | (8 * SINGLE), mode, 0, ws.data_ptr(), nb, st)
                    assert rc == 0, rc
                for _ in range(10): run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): run()
                e1.record(); torch.cuda.synchronize()
                row.append("flags %d: %.1f us" % (flags, e0.elapsed_time(e1) / 10 * 1e3))
            print(f"B {B} tasks {8 * B * (8 // spt)} steps/task {spt} mode {mode}  " + "  ".join(row), flush=True)
    del ws
