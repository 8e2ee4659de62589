"""Check (1 GPU, torchrun) that dropping NCCL_DEBUG=VERSION before NCCL initialises keeps
NCCL's version banner off stdout, as bench.py relies on."""
import os
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    del os.environ["NCCL_DEBUG"]
import torch
import torch.distributed as dist
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
t = torch.ones(4, device="cuda")
dist.all_reduce(t)
torch.cuda.synchronize()
print("STDOUT-LINE")
dist.destroy_process_group()
