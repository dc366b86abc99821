"""N-GPU check of shard.PeerGather (run under torchrun on one NVLink node):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 scripts/peer_gather_check.py
Compares 12 overlapped steps (even and ragged shards, producer overwriting its buffer right after start) with the NCCL
all-gather and prints which form ran + the time of 20 gathers of a 42 MB shard in both forms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from wavelet_monodepth_b200 import shard

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for n_global in (world * 4, world * 4 + 1):
    lo, hi = shard.shard_bounds(n_global, world, rank)
    pg = shard.PeerGather(n_global)
    buf = torch.empty(hi - lo, 1, 64, 256, device=dev)
    handles, wants = [], []
    for k in range(12):
        src = torch.arange(lo, hi, device=dev, dtype=torch.float32).view(-1, 1, 1, 1).expand_as(buf) * 10 + k
        buf.copy_(src)
        handles.append(pg.start(buf))
        buf.fill_(-1.0)
        wants.append(torch.arange(0, n_global, device=dev, dtype=torch.float32).view(-1, 1, 1, 1).expand(n_global, 1, 64, 256) * 10 + k)
        if k >= 1:
            got = handles[k - 1].wait()
            assert torch.equal(got, wants[k - 1]), (n_global, k - 1)
    assert torch.equal(handles[-1].wait(), wants[-1])
    torch.cuda.synchronize()
    if rank == 0:
        print("n_global %d: 12 overlapped steps equal the expected global tensor; form: %s" %
              (n_global, "peer copy engines" if pg._peer not in (None, False) else "NCCL fallback (%s)" % pg.why_not), flush=True)

n_global = world * 32
x = torch.rand(32, 1, 320, 1024, device=dev)
for name, g in (("peer", shard.PeerGather(n_global)), ("nccl", shard.OverlappedGather(n_global))):
    hs = [g.start(x)]
    hs[-1].wait(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        hs.append(g.start(x))
        hs.pop(0).wait()
    hs.pop(0).wait()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("%s: %.3f ms per all-gather of %d x 42 MB" % (name, e0.elapsed_time(e1) / 20, world), flush=True)
dist.barrier()
dist.destroy_process_group()
