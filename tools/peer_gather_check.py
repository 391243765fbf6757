"""Two (or more) ranks, one GPU each: the in-graph peer-store record gather (parallel.RecordGather mode "peer")
against the NCCL all-gather, over many steps with ranks deliberately skewed.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/peer_gather_check.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    pkg = entry.load_package()
    B = 4
    ids = [rank * B + i for i in range(B)]
    host = pkg.synth.make_batch(ids, n_dense=512, n_model=256, feat_dim=64, n_proposal1=512)
    kw = dict(n_dense=512, n_sparse=64, feat_dim=64, n_model=256, n_proposal1=512, n_proposal2=64)
    stage = pkg.stage.GeometricStage(B, dev, **kw)
    stage.instance_ids.copy_(torch.tensor(ids, dtype=torch.float32))
    stage.load(host)
    peer = pkg.parallel.RecordGather(stage, rank, world, mode="peer")
    ref_table = torch.empty((world * B, 16), device=dev)
    for step in range(40):
        if step % 7 == rank:          # skew the ranks: the wait kernel must really wait
            time.sleep(0.05)
        stage.instance_ids.copy_(torch.tensor(ids, dtype=torch.float32) + 1000.0 * step)
        torch.cuda.synchronize(dev)
        stage.run()                   # graph replay: pack + peer stores + wait are inside
        with torch.cuda.stream(stage.stream):
            dist.all_gather_into_tensor(ref_table, stage.records)
        stage.stream.synchronize()
        assert torch.equal(peer.table, ref_table), (rank, step, (peer.table != ref_table).nonzero()[:4])
        assert peer.table[:, 14].tolist() == [float(i + 1000 * step) for i in range(world * B)], (rank, step)
    assert not peer.timed_out()
    assert stage.kernels_per_step > 20 and stage.graph is not None
    dist.barrier()
    peer.close()
    if rank == 0:
        print("PEER_GATHER_OK world=%d steps=40" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
