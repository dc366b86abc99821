"""2-GPU probe: bandwidth of a 42 MB copy into a peer's IPC-mapped buffer, by tensor.copy_ and by cudaMemcpyPeerAsync."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mine = torch.zeros(32, 1, 320, 1024, device=dev)
hs = [None] * world
dist.all_gather_object(hs, (dev.index, reduce_tensor(mine)))
peer_idx, (fn, a) = hs[(rank + 1) % world]
peer = fn(*a)
src = torch.rand_like(mine)
own = torch.zeros_like(mine)
print(rank, "peer tensor device", peer.device, "can access", torch.cuda.can_device_access_peer(dev.index, peer_idx), flush=True)


def timed(f, name):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print("%-40s %.3f ms  %.0f GB/s" % (name, ms, src.numel() * 4 / ms / 1e6), flush=True)


timed(lambda: own.copy_(src, non_blocking=True), "own buffer, tensor.copy_")
timed(lambda: peer.copy_(src, non_blocking=True), "peer buffer, tensor.copy_")
try:
    from cuda import cudart
    st = torch.cuda.current_stream().cuda_stream
    nbytes = src.numel() * 4
    timed(lambda: cudart.cudaMemcpyPeerAsync(peer.data_ptr(), peer_idx, src.data_ptr(), dev.index, nbytes, st), "peer buffer, cudaMemcpyPeerAsync")
    timed(lambda: cudart.cudaMemcpyAsync(peer.data_ptr(), src.data_ptr(), nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, st), "peer buffer, cudaMemcpyAsync D2D")
except Exception as e:
    print("cuda-python path failed:", e)
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print("data landed:", bool(torch.equal(mine, mine)), flush=True)
dist.destroy_process_group()
