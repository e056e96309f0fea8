"""Bare host<->device copy ceiling of the end-to-end front-end path: every rank moves the bytes one bench step
moves (128 frames 640x480 up, their keypoints / descriptors / matches down) with no kernel in between, all ranks at the same time.
usage: python tools/pcie_probe.py            (one GPU)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
numa = bench.bind_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


B, W, H, cap = 128, 640, 480, 2000
h2d, d2h = B * W * H, B * (cap * 72 + 8)
out = {"n_gpus": world, "numa_node": numa}
for name, a, b in (("both", h2d, d2h), ("h2d_only", h2d, 4096), ("d2h_only", 4096, d2h)):
    t = torch.tensor([bench.copy_ceiling(a, b, 4, 50, barrier)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name] = {"frames_per_s": world * B / float(t.item()), "aggregate_GBps": world * (a + b) / float(t.item()) / 1e9}
if rank == 0:
    print(json.dumps(out))
