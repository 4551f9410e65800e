"""All-reduce latency of the X-update's T x k partial (1.6 MB fp32 at C2) under the NCCL settings in the environment.
torchrun --nproc-per-node N tools/nccl_probe.py ; NCCL_ALGO / NCCL_PROTO are read when the communicator is created."""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
for n in (400_000, 1_600_000 // 4 * 4, 6_400_000):
    x = torch.ones(n, dtype=torch.float32, device="cuda")
    for _ in range(20):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("N=%d algo=%s proto=%s  %8d floats: %.1f us per all-reduce" % (world, os.environ.get("NCCL_ALGO", "default"),
              os.environ.get("NCCL_PROTO", "default"), n, e0.elapsed_time(e1) * 1e3 / 200), flush=True)
dist.destroy_process_group()
