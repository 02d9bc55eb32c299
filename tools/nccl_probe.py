"""all_gather bandwidth and peer-access probe between the ranks of one box (diagnostic)."""
import os, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = dist.get_world_size()
print("probe can_access_peer", [torch.cuda.can_device_access_peer(local, j) for j in range(torch.cuda.device_count()) if j != local], flush=True)
for mb in (1, 21, 84):
    x = torch.zeros(mb * 1024 * 1024 // 8, dtype=torch.float64, device=dev); out = torch.empty(w * x.numel(), dtype=torch.float64, device=dev)
    for _ in range(3): dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize()
    if dist.get_rank() == 0: print(f"probe all_gather {mb} MB per rank: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, dtype=torch.float64, device=dev); h = symm.rendezvous(t, dist.group.WORLD.group_name)
    print("probe symmetric memory ok: peers", h.world_size, "buffer ptrs", [hex(p) for p in h.buffer_ptrs][:4], flush=True)
except Exception as e:
    print("probe symmetric memory failed:", type(e).__name__, str(e)[:300], flush=True)
dist.barrier(); dist.destroy_process_group()
