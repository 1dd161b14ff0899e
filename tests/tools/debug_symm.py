"""Manual experiment (2+ GPUs, torchrun): does torch symmetric memory work on this box, and can a peer buffer be written?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1024, 512, dtype=torch.float32, device=dev)
t.zero_()
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs][:4], "multicast ptr", h.multicast_ptr, flush=True)
h.barrier()
peer = (rank + 1) % world
pb = h.get_buffer(peer, (1024, 512), torch.float32)
pb += float(rank + 1)          # remote read-modify-write over NVLink
torch.cuda.synchronize()
h.barrier()
torch.cuda.synchronize()
exp = float(((rank - 1) % world) + 1)
print(rank, "value", float(t[0, 0]), "expected", exp, "ok" if float(t.mean()) == exp else "MISMATCH", flush=True)
# barrier cost
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): h.barrier()
e1.record(); torch.cuda.synchronize()
print(rank, "symm barrier us", e0.elapsed_time(e1) / 20 * 1e3, flush=True)
dist.destroy_process_group()
