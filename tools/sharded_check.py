"""Sharded global BA against the one-GPU path (and the compiled reference when oracle/_ref is there): same problem, same Huber widths.
usage: python tools/sharded_check.py [K P D]                      (one rank)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/sharded_check.py [K P D]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib
from mageslam_b200.sharded import ShardedGlobalBA
from tests.oracle_ba import rel_frobenius

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K, P, D = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (120, 6000, 8)
prob = synth.ba_problem(K=K, P=P, obs_per_point=D, seed=2, loop=True)
hub = [1.8]
one = BundlerLib().load(prob)                       # the whole problem on this rank's GPU: the answer to reproduce
sh = ShardedGlobalBA(prob, rank, world, dist if world > 1 else None)
for step in range(3):
    t0 = time.perf_counter(); m1 = one.StepBundleAdjustment(hub, 1e9); torch.cuda.synchronize(); t1 = time.perf_counter()
    ms = sh.StepBundleAdjustment(hub); torch.cuda.synchronize(); t2 = time.perf_counter()
    p1, r1 = one.poses(); p2, r2 = sh.poses()
    ids, pts = sh.points()
    e = (rel_frobenius(p2, p1), rel_frobenius(r2, r1), rel_frobenius(pts, one.points()[ids]))
    if rank == 0:
        print("step %d: one GPU %.2f ms (mean %.6f, lambda %.6g) | %d rank(s) %.2f ms (mean sq. error %.6f, lambda %.6g, %d trials) | relF positions %.2e rotations %.2e points %.2e" % (
            step, (t1 - t0) * 1e3, m1, one.GetCurrentLambda(), world, (t2 - t1) * 1e3, ms, sh.GetCurrentLambda(), sh.trials, *e))
    assert max(e) < 1e-4 and abs(sh.GetCurrentLambda() - one.GetCurrentLambda()) <= 1e-6 * abs(one.GetCurrentLambda()), (rank, e)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
