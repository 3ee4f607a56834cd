"""Minimal NCCL probe: init + one all_reduce per rank (torchrun)."""
import os, time, torch, torch.distributed as dist
t0 = time.time()
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(1 << 20, device="cuda") * (local + 1)
dist.all_reduce(x); torch.cuda.synchronize()
print("rank %d ok: sum %.1f in %.1f s" % (local, float(x[0]), time.time() - t0), flush=True)
dist.destroy_process_group()
