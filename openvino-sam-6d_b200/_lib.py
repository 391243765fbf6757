"""ctypes binding of libpem_b200.so (the C ABI declared in include/pem_b200.h).

The library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``).  There is
no fallback of any kind: if the shared object is missing or a CUDA device is absent the
ops raise -- the product path never routes through CPU code.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# PEM_B200_LIB: load a side build of the same sources instead (A/B timing of kernel variants, tools/ only)
LIB_PATH = os.environ.get("PEM_B200_LIB") or os.path.join(_HERE, "libpem_b200.so")
_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_f = ctypes.c_float
c_d = ctypes.c_double
c_ll = ctypes.c_longlong
c_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of include/pem_b200.h
SIGNATURES = {
    "pem_abi_version": (c_i, []),
    "pem_build_info": (ctypes.c_char_p, []),
    "pem_status_string": (ctypes.c_char_p, [c_i]),
    "pem_fps_workspace_bytes": (c_sz, [c_i, c_i]),
    "pem_fps": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_sz, c_p]),
    "pem_fps_gather": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_sz, c_p]),
    "pem_gather": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "pem_gather_rows": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_ll, c_ll, c_p, c_ll, c_p]),
    "pem_transform_points": (c_i, [c_p, c_p, c_p, c_i, c_i, c_f, c_p, c_p]),
    "pem_ball_query": (c_i, [c_p, c_p, c_i, c_i, c_i, c_f, c_i, c_p, c_p]),
    "pem_group": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "pem_query_and_group": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_p, c_p, c_p]),
    "pem_pe_group2": (c_i, [c_p, c_i, c_i, c_f, c_f, c_i, c_p, c_f, c_i, c_p, c_p]),
    "pem_similarity_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "pem_feature_similarity": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_p, c_p, c_sz, c_p]),
    "pem_coarse_assign": (c_i, [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "pem_coarse_hypotheses": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "pem_select_smallest": (c_i, [c_p, c_i, c_i, c_i, c_p, c_p]),
    "pem_score_hypotheses": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_i, c_p, c_i, c_p, c_p]),
    "pem_score_hypotheses_ex": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_p]),
    "pem_select_best": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "pem_fine_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "pem_fine_Rt": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p, c_p, c_p,
                          c_p, c_p, c_p, c_sz, c_p]),
    "pem_fine_fused_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i]),
    "pem_feature_similarity_tc": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_p, c_p, c_sz, c_p]),
    "pem_fine_fused": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p,
                             c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "pem_fine_fused_ex": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_i, c_p, c_p,
                                c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "pem_fine_fused_pass_ex": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_p, c_sz, c_p]),
    "pem_pe_masks_bytes": (c_sz, [c_i, c_i]),
    "pem_pe_masks": (c_i, [c_p, c_i, c_i, c_f, c_f, c_f, c_i, c_p, c_sz, c_p]),
    "pem_pe_group2_masked": (c_i, [c_p, c_i, c_i, c_f, c_f, c_i, c_p, c_f, c_i, c_p, c_p, c_sz, c_p]),
    "pem_fine_fused_pass": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_f, c_i, c_p, c_sz, c_p]),
    "pem_pe_neighbours_masked": (c_i, [c_p, c_i, c_i, c_f, c_f, c_i, c_p, c_f, c_i, c_p, c_p, c_sz, c_p]),
    "pem_pe_mlp": (c_i, [c_p, c_p, c_i, c_i, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p]),
    "pem_geo_embedding_indices": (c_i, [c_p, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p]),
    "pem_depth_to_metres": (c_i, [c_p, c_i, c_i, c_f, c_p, c_p]),
    "pem_backproject_depth": (c_i, [c_p, c_i, c_i, c_d, c_d, c_d, c_d, c_i, c_p, c_p]),
    "pem_instance_scratch_bytes": (c_sz, [c_i, c_i, c_i]),
    "pem_instance_candidates": (c_i, [c_p, c_p, c_i, c_i, c_i, c_d, c_d, c_d, c_d, c_i, c_d, c_p, c_p, c_sz, c_p]),
    "pem_instance_sample": (c_i, [c_p, c_i, c_i, c_d, c_d, c_d, c_d, c_i, c_p, c_p, c_sz, c_i, c_p, c_i, c_p,
                                  c_i, c_i, c_p, c_p, c_p, c_p]),
    "pem_pack_records": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p]),
    "pem_weighted_procrustes": (c_i, [c_p, c_p, c_p, c_i, c_i, c_f, c_f, c_p, c_p, c_p]),
    "pem_svd3_batched": (c_i, [c_p, c_ll, c_p, c_p, c_p, c_p]),
    "pem_det3_batched": (c_i, [c_p, c_ll, c_p, c_p]),
    "pem_searchsorted": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "pem_bench_fp32_peak": (c_i, [c_i, c_i, c_i, c_p, c_p]),
    "pem_peer_table_bytes": (c_sz, [c_i, c_i]),
    "pem_peer_alloc": (c_i, [c_sz, ctypes.POINTER(c_p), c_p]),
    "pem_peer_open": (c_i, [c_p, ctypes.POINTER(c_p)]),
    "pem_peer_release": (c_i, [c_p, c_i]),
    "pem_gather_poses": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, ctypes.POINTER(c_p), c_p, c_p]),
    "pem_gather_poses_wait": (c_i, [c_p, c_i, c_i, c_p, c_i, c_p]),
    "pem_half_to_float": (c_i, [c_p, c_p, c_sz, c_p]),
}


def build(force=False):
    """Compile libpem_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if force:
        subprocess.check_call(cmd + ["clean"])
    subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libpem_b200.so not found at %s: run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().pem_status_string(int(status)).decode()
        raise RuntimeError("%s failed: %s (status %d)" % (what, msg, status))
