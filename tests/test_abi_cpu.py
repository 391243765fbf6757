"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol
include/pem_b200.h declares, and validates arguments before touching the GPU."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pem_[a-z0-9_A-Z]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(pkg):
    lib = pkg._lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    # and the ctypes table covers exactly the header
    assert sorted(pkg._lib.SIGNATURES) == names


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "pem_b200.h")).read()
    assert "torch" not in src.lower().replace("pytorch", "") or "at::" not in src
    assert 'extern "C"' in src and "at::Tensor" not in src


def test_precision_constants_match_the_header(pkg):
    """The Python mirrors pass PEM_PREC_* by value: they must be the header's numbers."""
    src = open(os.path.join(ROOT, "include", "pem_b200.h")).read()
    defs = dict(re.findall(r"#define\s+(PEM_PREC_[A-Z0-9_]+)\s+(\d+)", src))
    want = {"split_bf16": int(defs["PEM_PREC_SPLIT_BF16"]), "fp16": int(defs["PEM_PREC_FP16"])}
    assert pkg.model_utils.FINE_PRECISIONS == want
    assert pkg.stage.FINE_PRECISIONS == want


def test_identification(pkg):
    lib = pkg._lib.lib()
    assert lib.pem_abi_version() == 1
    assert b"sm_100a" in lib.pem_build_info()
    assert lib.pem_status_string(0) == b"ok"
    assert b"NULL" in lib.pem_status_string(-1)


def test_argument_validation_without_gpu(pkg):
    """Invalid arguments are rejected with negative codes before any CUDA call."""
    lib = pkg._lib.lib()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.pem_fps(None, 1, 8, 2, 1, p, None, 0, None) == -1
    assert lib.pem_fps(p, 0, 8, 2, 1, p, None, 0, None) == -2
    assert lib.pem_fps(p, 1, 300000, 2, 1, p, None, 0, None) == -4      # needs a workspace
    assert lib.pem_gather(p, p, 1, 0, 8, 2, p, None) == -2
    assert lib.pem_ball_query(p, p, 1, 4, 4, 0.1, 500, p, None) == -3   # nsample > 128
    assert lib.pem_coarse_assign(p, 1, 500, 500, p, p, p, None) == -3   # matrix exceeds the cluster capacity
    assert lib.pem_select_smallest(p, 1, 10, 11, p, None) == -2         # P2 > P1
    assert lib.pem_fine_Rt(p, p, p, p, 1, 8, 8, 8, 0.15, p, p, p, None, None, None, None, None, 0, None) == -4
    # the tensor-core entry points
    assert lib.pem_pe_masks_bytes(2, 2048) == 2 * 2048 * 64 * 4
    assert lib.pem_pe_masks(None, 1, 64, 1e-8, 0.2, 1e-3, 0, p, 64, None) == -1
    assert lib.pem_pe_masks(p, 1, 8192, 1e-8, 0.2, 1e-3, 0, p, 1 << 30, None) == -3     # cloud does not fit in shared memory
    assert lib.pem_pe_masks(p, 1, 64, 1e-8, 0.2, 1e-3, 0, p, 8, None) == -4             # mask buffer too small
    assert lib.pem_pe_group2_masked(p, 1, 64, 1e-8, 0.1, 32, p, 0.2, 64, p, p, 8, None) == -4
    assert lib.pem_fine_fused(p, p, p, p, p, 1, 8, 8, 100, 8, 0.1, 0.15, p, p, p, None, None, None, None, p, 1 << 30,
                              None) == -3                                                # C must be a multiple of 64
    assert lib.pem_fine_fused(p, p, p, p, p, 1, 8, 8, 64, 8, 0.1, 0.15, p, p, p, None, None, None, None, p, 16,
                              None) == -4
    assert lib.pem_fine_fused_pass(p, p, 1, 8, 8, 64, 0.1, 7, p, 1 << 30, None) == -3
    # ... and their precision-taking forms: PEM_PREC_SPLIT_BF16 = 0, PEM_PREC_FP16 = 1, nothing else
    assert lib.pem_fine_fused_ex(p, p, p, p, p, 1, 8, 8, 64, 8, 0.1, 0.15, 2, p, p, p, None, None, None, None, p,
                                 1 << 30, None) == -3
    assert lib.pem_fine_fused_ex(p, p, p, p, p, 1, 8, 8, 64, 8, 0.1, 0.15, 1, p, p, p, None, None, None, None, p,
                                 16, None) == -4
    assert lib.pem_fine_fused_ex(p, p, p, p, p, 1, 8, 8, 100, 8, 0.1, 0.15, 1, p, p, p, None, None, None, None, p,
                                 1 << 30, None) == -3
    assert lib.pem_fine_fused_pass_ex(p, p, 1, 8, 8, 64, 0.1, 1, 5, p, 1 << 30, None) == -3
    # option bits on the precision word: PEM_FINE_THREE_PASS = 0x100, PEM_FINE_STEPS_PER_TASK(n) = n << 16; others rejected
    assert lib.pem_fine_fused_ex(p, p, p, p, p, 1, 8, 8, 64, 8, 0.1, 0.15, 1 | 0x200, p, p, p, None, None, None, None,
                                 p, 1 << 30, None) == -3
    assert lib.pem_fine_fused_ex(p, p, p, p, p, 1, 8, 8, 64, 8, 0.1, 0.15, 1 | 0x100 | (2 << 16), p, p, p, None, None,
                                 None, None, p, 16, None) == -4      # options accepted: stops at the workspace check
    # the scoring kernel named by the caller: 0 auto, 1 SIMT, 2 tensor cores
    assert lib.pem_score_hypotheses_ex(p, p, None, 1, 8, 8, p, p, 4, p, 16, p, 3, None) == -2
    assert lib.pem_score_hypotheses_ex(p, p, None, 1, 8, 8, p, p, 300, p, 16, p, 0, None) == -3   # > 256 query points
    # instance preparation
    assert lib.pem_instance_scratch_bytes(3, 480, 640) == 3 * 480 * 640 * 4
    assert lib.pem_backproject_depth(p, 4, 4, 1.0, 1.0, 0.0, 0.0, 2, p, None) == -3     # precision is 0 or 1
    assert lib.pem_depth_to_metres(None, 4, 4, 1.0, p, None) == -1
    assert lib.pem_instance_candidates(p, p, 1, 4, 4, 1.0, 1.0, 0.0, 0.0, 0, 0.2, p, p, 8, None) == -4
    assert lib.pem_instance_sample(p, 4, 4, 1.0, 1.0, 0.0, 0.0, 0, p, p, 64, 1, p, 0, p, 8, 224, p, p, None,
                                   None) == -2                                           # no kept detection
    # operator seam
    assert lib.pem_svd3_batched(None, 4, p, p, p, None) == -1
    assert lib.pem_svd3_batched(p, 4, None, None, None, None) == -1      # at least one output
    assert lib.pem_svd3_batched(p, 0, p, p, p, None) == -2
    assert lib.pem_det3_batched(p, 4, None, None) == -1
    assert lib.pem_searchsorted(p, p, 1, 0, 4, p, None) == -2
    assert lib.pem_searchsorted(p, None, 1, 8, 4, p, None) == -1
    assert lib.pem_fps_workspace_bytes(4, 2048) == 0
    assert lib.pem_fps_workspace_bytes(2, 210000) == 0 and lib.pem_fps_workspace_bytes(2, 300000) == 2 * 300000 * 4
    assert lib.pem_fine_workspace_bytes(16, 2048, 2048) > 16 * 2049 * 4 * 6


def test_product_package_never_imports_the_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    pk = os.path.join(ROOT, "openvino-sam-6d_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pem_oracle" not in text and "cpu_stage" not in text and "oracle/" not in text, f


def test_missing_library_fails_loudly(pkg, monkeypatch):
    import pytest
    monkeypatch.setattr(pkg._lib, "_lib", None)
    monkeypatch.setattr(pkg._lib, "LIB_PATH", "/nonexistent/libpem_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg._lib.lib()
