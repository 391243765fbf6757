"""Drop-in for the reference's compiled module ``pointnet2._ext``
(model/pointnet2/_ext_src/src/bindings.cpp:11-24).

Same function names, argument order, dtypes, output shapes and error behaviour
(``RuntimeError`` for a non-contiguous or wrongly typed tensor, like the reference's
TORCH_CHECKs in _ext_src/include/utils.h:20-35) for the four ops on the inference path.
Outputs are freshly allocated on the input's device; kernels are enqueued on the current
CUDA stream without synchronising.  CUDA tensors only: there is no CPU path.

The five symbols the inference path never calls (gradients, three_nn / three_interpolate)
exist so that ``import pointnet2._ext`` consumers keep importing, and raise when called.
"""
import torch

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t, name, dtype, what):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s tensor" % (name, what))
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (libpem_b200 has no CPU path)" % name)


def furthest_point_sampling(points, nsamples, skip_near_origin=True):
    """(B,N,3) f32, int -> (B,nsamples) i32.  sampling.h:11 / sampling.cpp:184-212.

    ``skip_near_origin`` selects the reference CPU rule (default, the oracle) or the
    reference CUDA rule (sampling_gpu.cu:106); ties always go to the lowest index."""
    _check(points, "points", torch.float32, "a float")
    B, N, _ = points.shape
    M = int(nsamples)
    out = torch.empty((B, M), dtype=torch.int32, device=points.device)
    if B == 0 or M == 0:
        return out
    L = _lib.lib()
    ws_bytes = L.pem_fps_workspace_bytes(B, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=points.device) if ws_bytes else None
    with torch.cuda.device(points.device):
        _lib.check(L.pem_fps(_ptr(points), B, N, M, 1 if skip_near_origin else 0, _ptr(out),
                             _ptr(ws), ws_bytes, _stream()), "pem_fps")
    return out


def gather_points(points, idx):
    """(B,C,N) f32, (B,M) i32 -> (B,C,M).  sampling.h:9 / sampling.cpp:120-150."""
    _check(points, "points", torch.float32, "a float")
    _check(idx, "idx", torch.int32, "an int")
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().pem_gather(_ptr(points), _ptr(idx), B, C, N, M, _ptr(out),
                                         _stream()), "pem_gather")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """(B,M,3),(B,N,3) f32 -> (B,M,nsample) i32.  ball_query.h:8-9 / ball_query.cpp:64-94."""
    _check(new_xyz, "new_xyz", torch.float32, "a float")
    _check(xyz, "xyz", torch.float32, "a float")
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    out = torch.empty((B, M, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(new_xyz.device):
        _lib.check(_lib.lib().pem_ball_query(_ptr(new_xyz), _ptr(xyz), B, M, N, float(radius),
                                             int(nsample), _ptr(out), _stream()),
                   "pem_ball_query")
    return out


def group_points(points, idx):
    """(B,C,N) f32, (B,M,ns) i32 -> (B,C,M,ns).  group_points.h:8 / group_points.cpp:79-108."""
    _check(points, "points", torch.float32, "a float")
    _check(idx, "idx", torch.int32, "an int")
    B, C, N = points.shape
    _, M, ns = idx.shape
    out = torch.empty((B, C, M, ns), dtype=torch.float32, device=points.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(points.device):
        _lib.check(_lib.lib().pem_group(_ptr(points), _ptr(idx), B, C, N, M, ns, _ptr(out),
                                        _stream()), "pem_group")
    return out


def _inference_only(name):
    def fn(*_a, **_k):
        raise NotImplementedError(
            "%s is outside the PEM inference hot path (SURVEY.md section 2, component 1) and is "
            "not provided by libpem_b200" % name)
    fn.__name__ = name
    return fn


gather_points_grad = _inference_only("gather_points_grad")
group_points_grad = _inference_only("group_points_grad")
three_nn = _inference_only("three_nn")
three_interpolate = _inference_only("three_interpolate")
three_interpolate_grad = _inference_only("three_interpolate_grad")
