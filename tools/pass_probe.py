import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
L = ctypes.CDLL(pkg._lib.LIB_PATH)
L.pem_dev_fine_pass.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
L.pem_fine_fused_workspace_bytes.restype = ctypes.c_size_t
B, N, C = 16, 2048, 256
f1 = torch.randn(B, N + 1, C, device=dev); f2 = torch.randn(B, N + 1, C, device=dev)
nb = L.pem_fine_fused_workspace_bytes(B, N, N, C)
ws = torch.empty(nb, dtype=torch.uint8, device=dev); scratch = torch.zeros(1024, device=dev)
st = torch.cuda.current_stream().cuda_stream
for cg in (0, 2):
    for mode in (4, 5):
        def run():
            rc = L.pem_dev_fine_pass(f1.data_ptr(), f2.data_ptr(), B, N, N, C, 0.1, mode, cg, scratch.data_ptr(), ws.data_ptr(), nb, st)
            assert rc == 0, rc
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        print("cg", cg, "mode", {4: "no epilogue", 5: "TMEM loads only"}[mode], "prep+pass: %.1f us" % (e0.elapsed_time(e1) / 10 * 1e3))
