# NOTE: uses pem_dev_* tuning knobs: build the library with `make -C openvino-sam-6d_b200/csrc clean all EXTRA=-DPEM_DEV_KNOBS`
# (the shipped library keeps these as compile-time constants).
"""Timing experiments on one fine GEMM pass: which role bounds it?
flags: 1 no MMAs, 2 no streamed TMA loads, 4 no epilogue math.  mode: 1 stats, 4 label + wsum (the merged pass), 2 label, 3 wsum."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
L = ctypes.CDLL(pkg._lib.LIB_PATH)
L.pem_dev_fine_pass.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
B = int(os.environ.get("B", 16))
base = pkg.synth.make_batch(list(range(4)))
b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, B // 4).items()}
bg = b["bg"][:, None, :]
f1, f2 = torch.cat([bg, b["dense_fm"]], 1).contiguous(), torch.cat([bg, b["dense_fo"]], 1).contiguous()
N, C = f1.shape[1] - 1, f1.shape[2]
nb = pkg._lib.lib().pem_fine_fused_workspace_bytes(B, N, N, C)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
R = torch.empty(B, 9, device=dev); t = torch.empty(B, 3, device=dev); sc = torch.empty(B, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib = pkg._lib.lib()
SINGLE = int(os.environ.get("SINGLE", 0))  # 1: the single-product fp16 mode (flag bit 3)
L.pem_dev_set_fine_epi_warps(int(os.environ.get("EPW", 16)))
L.pem_dev_set_fine_ring(int(os.environ.get("RING", 9)))
rc = lib.pem_fine_fused_ex(f1.data_ptr(), f2.data_ptr(), b["dense_pm"].data_ptr(), b["dense_po"].data_ptr(), b["model"].data_ptr(),
                           B, N, N, C, b["model"].shape[1], 0.1, 0.15, SINGLE, R.data_ptr(), t.data_ptr(), sc.data_ptr(),
                           None, None, None, None, ws.data_ptr(), nb, st)
assert rc == 0, rc
torch.cuda.synchronize()
for spt in [int(x) for x in os.environ.get("SPT", "0").split(",")]:
    L.pem_dev_set_fine_steps_per_task(spt)
    for mode in [int(x) for x in os.environ.get("MODES", "1,4,2,3").split(",")]:  # statistics, labels + weighted sums (merged), labels, weighted sums
        for flags in [int(x) for x in os.environ.get("FLAGS", "0,4,2,6,7").split(",")]:
            def run(prep=0):
                rc = L.pem_dev_fine_pass(f1.data_ptr(), f2.data_ptr(), B, N, N, C, 0.1, flags | (8 * SINGLE), mode, prep, ws.data_ptr(), nb, st)
                assert rc == 0, rc
            for _ in range(20): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): run()
            e1.record(); torch.cuda.synchronize()
            names = [n for bit, n in ((1, "noMMA"), (2, "noLoad"), (4, "noEpi")) if flags & bit]
            print(f"single {SINGLE} epw {os.environ.get('EPW', 16)} ring {os.environ.get('RING', 9)} steps/task {spt} mode {mode} {'+'.join(names) or 'full'}: pass {e0.elapsed_time(e1) / 10 * 1e3:.1f} us", flush=True)
