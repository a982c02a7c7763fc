"""NCCL all-reduce timing at the training step's bucket sizes (one process per GPU, under torchrun).
Usage: python -m torch.distributed.run --nproc-per-node N tools/allreduce_probe.py"""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for name, n, dt in (("1 MB fp32", 1 << 18, torch.float32), ("33 MB fp32 (ET arena)", 8_192_939, torch.float32),
                    ("103 MB bf16 (trunk arena)", 51_602_144, torch.bfloat16), ("206 MB fp32 (trunk arena)", 51_602_144, torch.float32),
                    ("1 GiB fp32", 1 << 28, torch.float32)):
    x = torch.ones(n, dtype=dt, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 10
    e0.record()
    for _ in range(it):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    nbytes = n * x.element_size()
    if rank == 0:
        print(f"{name:28s} {ms:8.3f} ms  algbw {nbytes / ms / 1e6:7.1f} GB/s  busbw {nbytes / ms / 1e6 * 2 * (world - 1) / world:7.1f} GB/s", flush=True)
dist.destroy_process_group()
