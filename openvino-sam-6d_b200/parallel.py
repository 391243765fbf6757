"""Multi-GPU plumbing for the geometric stage: detected instances are independent, so each rank
owns a contiguous slice of the batch (no data-path collective) and only the final 64-byte
pose record per instance is gathered.  The reference has no
counterpart (its inference path is single-device)."""
import torch

RECORD_FLOATS = 16  # R (9) | t (3) | pose_score | proposal id | instance id | pad  = 64 bytes


def shard_instances(n_instances, world):
    """Contiguous [lo, hi) slices; the first n % world ranks get one extra instance."""
    base, extra = divmod(n_instances, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def pack_records(R, t, score, proposal_id, instance_id):
    n = R.shape[0]
    rec = torch.zeros((n, RECORD_FLOATS), dtype=torch.float32, device=R.device)
    rec[:, 0:9] = R.reshape(n, 9)
    rec[:, 9:12] = t
    rec[:, 12] = score
    rec[:, 13] = proposal_id.to(torch.float32)
    rec[:, 14] = instance_id.to(torch.float32)
    return rec


def unpack_records(rec):
    n = rec.shape[0]
    return (rec[:, 0:9].reshape(n, 3, 3), rec[:, 9:12], rec[:, 12], rec[:, 13].to(torch.int64),
            rec[:, 14].to(torch.int64))


def gather_records(rec, n_instances, world, out=None):
    """All ranks end up with the (n_instances,16) record table ordered by rank slice.  Even slices
    (the usual case) are ONE collective straight into `out` (preallocated by the caller, or fresh);
    uneven slices are padded to the largest slice for the collective and trimmed afterwards."""
    import torch.distributed as dist
    if world == 1:
        return rec
    shards = shard_instances(n_instances, world)
    width = max(hi - lo for lo, hi in shards)
    if all(hi - lo == width for lo, hi in shards) and rec.shape[0] == width and rec.is_contiguous():
        if out is None:
            out = torch.empty((world * width, RECORD_FLOATS), dtype=torch.float32, device=rec.device)
        dist.all_gather_into_tensor(out, rec)
        return out
    padded = torch.zeros((width, RECORD_FLOATS), dtype=torch.float32, device=rec.device)
    padded[: rec.shape[0]] = rec
    full = torch.empty((world * width, RECORD_FLOATS), dtype=torch.float32, device=rec.device)
    dist.all_gather_into_tensor(full, padded)
    parts = [full[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(shards)]
    return torch.cat(parts, 0)


# ---------------------------------------------------------------------------------------------
# The per-step record gather of a running stage
# ---------------------------------------------------------------------------------------------
class RecordGather:
    """Every rank ends each step with the (world * B, 16) record table of ALL ranks in `self.table`.

    mode "peer" (default at world > 1): no collective call.  The stage's last kernel (pem_gather_poses)
    stores its records straight into every rank's table over NVLink and a wait kernel (pem_gather_poses_wait)
    completes the step -- both INSIDE the stage's CUDA graph, so a step at N GPUs is one graph launch like at
    1 GPU.  Setup (once, collective over the default process group): every rank allocates its table in the
    library (cudaMalloc + IPC handle), the 64-byte handles are all-gathered, peers are opened.
    mode "nccl": one `all_gather_into_tensor` of the stage's records after the graph replay (run()).
    world == 1: the table is the stage's own record buffer.

    Attach BEFORE the stage's first run (the launch plan changes)."""

    def __init__(self, stage, rank, world, mode="peer", timeout_ms=2000):
        import ctypes
        self.stage, self.rank, self.world = stage, rank, world
        self.mode = mode if world > 1 else "local"
        B = stage.B
        self._opened, self._own = [], None
        if self.mode == "local":
            self.table = stage.records
            self.how = "single GPU: none"
            return
        self.table = torch.empty((world * B, RECORD_FLOATS), dtype=torch.float32, device=stage.dev)
        if self.mode == "nccl":
            self.how = "nccl all_gather_into_tensor of 64-byte records after the graph replay"
            return
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        nbytes = L.pem_peer_table_bytes(world, B)
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(stage.dev):
            _lib.check(L.pem_peer_alloc(nbytes, ctypes.byref(own), handle), "pem_peer_alloc")
        self._own = own.value
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=stage.dev)
        allh = torch.empty((world, 64), dtype=torch.uint8, device=stage.dev)
        dist.all_gather_into_tensor(allh, mine)
        allh = allh.cpu()
        ptrs = (ctypes.c_void_p * world)()
        for r in range(world):
            if r == rank:
                ptrs[r] = self._own
                continue
            h = (ctypes.c_ubyte * 64)(*allh[r].tolist())
            pp = ctypes.c_void_p()
            with torch.cuda.device(stage.dev):
                _lib.check(L.pem_peer_open(h, ctypes.byref(pp)), "pem_peer_open (rank %d)" % r)
            ptrs[r] = pp.value
            self._opened.append(pp.value)
        self._ptrs, self._L, self.timeout_ms = ptrs, L, int(timeout_ms)
        dist.barrier()  # every table exists and is zeroed before any rank's first step
        stage.gather = self
        stage.plan = None
        stage.graph = None
        self.how = "peer stores over NVLink from the record-packing kernel + flag wait, inside the step's CUDA " \
                   "graph (pem_gather_poses / pem_gather_poses_wait); no collective call"

    # the two launches the stage's plan uses in place of pem_pack_records (mode "peer")
    def launch_push(self, st):
        s, p = self.stage, (lambda t: t.data_ptr())
        return self._L.pem_gather_poses(p(s.R), p(s.t), p(s.score), p(s.proposal_id), p(s.instance_ids), s.B,
                                        self.rank, self.world, self._ptrs, p(s.records), st)

    def launch_wait(self, st):
        return self._L.pem_gather_poses_wait(self._own, self.world, self.stage.B, self.table.data_ptr(),
                                             self.timeout_ms, st)

    def run(self):
        """Per-step host call on the stage's stream: only the NCCL mode has anything left to launch."""
        if self.mode == "nccl":
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.table, self.stage.records)

    def timed_out(self):
        """True when a wait kernel gave up on a peer (header word 1 of the own table)."""
        if self.mode != "peer":
            return False
        torch.cuda.synchronize(self.stage.dev)
        hdr = _as_tensor(self._own, 2, torch.int32, self.stage.dev)  # a tensor view of the table's header
        return bool(hdr.cpu()[1].item())

    def close(self):
        if self.mode != "peer" or self._own is None:
            return
        torch.cuda.synchronize(self.stage.dev)
        for p in self._opened:
            self._L.pem_peer_release(p, 1)
        self._L.pem_peer_release(self._own, 0)
        self._opened, self._own = [], None


class _RawCuda:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def _as_tensor(ptr, n, dtype, device):
    typestr = {torch.int32: "<i4", torch.float32: "<f4", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_RawCuda(ptr, n, typestr), device=device)


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPU cores nearest its GPU (the `CPU Affinity` column of `nvidia-smi topo -m`)
    BEFORE it allocates pinned host buffers, so that first-touch places them on the GPU's NUMA node: at 8
    ranks the host-facing path reads ~0.5 GB/s x 8 of pinned memory and is limited by the host memory system.
    Best effort: returns the core list, or None when the topology cannot be read."""
    import os
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    lines = [l for l in out.splitlines() if l.strip()]
    if not lines:
        return None
    header = lines[0]
    import re
    cols = re.split(r"\t+|\s{2,}", header.strip())
    try:
        k = [c.strip() for c in cols].index("CPU Affinity")
    except ValueError:
        return None
    for l in lines[1:]:
        cells = re.split(r"\t+|\s{2,}", l.strip())
        if cells and cells[0].strip() == "GPU%d" % local_rank and len(cells) > k + 1:
            spec = cells[k + 1].strip()  # +1: the row label column
            cores = set()
            try:
                for part in spec.split(","):
                    lo, _, hi = part.partition("-")
                    cores.update(range(int(lo), int(hi or lo) + 1))
                allowed = os.sched_getaffinity(0)
                cores = (cores & allowed) or None
                if cores:
                    os.sched_setaffinity(0, cores)
                    return sorted(cores)
            except (ValueError, OSError):
                return None
    return None
