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
