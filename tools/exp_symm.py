import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = symm.empty(1 << 20, dtype=torch.int64, device=torch.device("cuda", lr))
t.zero_()
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, [n for n in dir(hdl) if not n.startswith("_")][:20], flush=True)
peers = [hdl.get_buffer(r, (1 << 20,), torch.int64) for r in range(world)]
torch.cuda.synchronize(); dist.barrier()
# every rank writes its id into slot [rank] of every peer's buffer
for r in range(world):
    peers[r][rank * 8:(rank + 1) * 8].fill_(100 + rank)
torch.cuda.synchronize(); dist.barrier()
print(rank, "local view", t[: world * 8].tolist(), "peer ptrs", [hex(p.data_ptr()) for p in peers], flush=True)
dist.destroy_process_group()
