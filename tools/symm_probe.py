"""Probe: does torch symmetric memory (peer buffers + signal pads over NVLink) work on this box?
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; multicast", hdl.has_multicast_support, "signal pad", hdl.signal_pad_size, flush=True)
t.fill_(float(rank + 1))
hdl.barrier(0)
peer = (rank + 1) % world
pbuf = hdl.get_buffer(peer, (1 << 20,), torch.float32, 0)
mine = torch.full((1 << 20,), 100.0 + rank, device=dev)
pbuf[: 1 << 19].copy_(mine[: 1 << 19])              # push into the peer's buffer
hdl.put_signal(peer, 1)
hdl.wait_signal((rank - 1) % world, 1)
torch.cuda.synchronize()
print(rank, "after push:", float(t[0]), float(t[-1]), flush=True)
# capture push + signal in a CUDA graph and replay
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        pbuf[: 1 << 19].copy_(mine[: 1 << 19]); hdl.put_signal(peer, 1); hdl.wait_signal((rank - 1) % world, 1)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        mine.add_(1.0)
        pbuf[: 1 << 19].copy_(mine[: 1 << 19]); hdl.put_signal(peer, 1); hdl.wait_signal((rank - 1) % world, 1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print(rank, "graph replay ok:", float(t[0]), "(expect", 100.0 + (rank - 1) % world + 3, ")", flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        g.replay()
    b.record(); b.synchronize()
    print(rank, "us per push(2MB)+signal+wait in graph:", a.elapsed_time(b) * 1e3 / 50, flush=True)
except Exception as e:
    print(rank, "graph capture failed:", repr(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
