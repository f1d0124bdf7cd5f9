"""Probe: torch symmetric memory on this box (peer pointers, barrier, P2P store bandwidth)."""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 64 * 1024 * 1024            # int64 elements = 512 MB
t = sm.empty(n, dtype=torch.int64, device=f"cuda:{lr}")
h = sm.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs][:4], "signal", hasattr(h, "signal_pad_ptrs"), [m for m in dir(h) if not m.startswith("_")])
src = torch.full((n // world,), rank + 1, dtype=torch.int64, device=f"cuda:{lr}")
h.barrier()
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for d in range(world):
        peer = h.get_buffer(d, (n,), torch.int64)
        peer[rank * (n // world):(rank + 1) * (n // world)].copy_(src)        # P2P store into peer d
    e1.record(); torch.cuda.synchronize()
    h.barrier()
    ms = e0.elapsed_time(e1)
    print(rank, f"iter {it}: pushed {src.numel()*8*world/1e6:.0f} MB in {ms:.3f} ms = {src.numel()*8*world/ms/1e6:.0f} GB/s")
ok = all(int(t[r * (n // world)].item()) == r + 1 for r in range(world))
print(rank, "data ok", ok)
dist.destroy_process_group()
