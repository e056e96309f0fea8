#!/usr/bin/env python3
"""BASELINE config 5: independent 1280x720 sequences, one per GPU: ORB extract + match per frame and a local-BA window (config-3
shaped, 10 KF / 2000 pts / 8000 obs, 10 LM iterations) every 20 frames. Replicas only -- no data-path collective.
    python tools/config5_sequences.py [--frames 256]                     (1 GPU)
    python -m torch.distributed.run --nproc-per-node N tools/config5_sequences.py   (N GPUs, NCCL barrier + max-reduce only)
Prints one JSON line with aggregate frames/s and BA LM iterations/s."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ba-every", type=int, default=20)
    ap.add_argument("--pipelined", type=int, default=1, help="1: mage_frontend_submit/_wait, 0: synchronous mage_frontend_process")
    a = ap.parse_args()
    import torch, torch.distributed as dist
    from mageslam_b200 import synth
    from mageslam_b200.frontend import FrontEnd
    from mageslam_b200.orb import FeatureExtractorSettings
    from mageslam_b200.bundler import BundlerLib, StepMany
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, B = 1280, 720, a.batch
    vid = synth.video_frames(32, W, H, seed=10 + rank)                         # sequence seeds 10..17
    frames = torch.from_numpy(np.concatenate([vid] * (a.frames // 32), 0)).pin_memory()
    fe = FrontEnd(FeatureExtractorSettings.tier(), W, H, B, chunk=B)
    outs = fe.alloc_outputs(pinned=True)
    n_windows = a.frames // a.ba_every
    windows = [BundlerLib().load(synth.ba_problem(seed=100 * rank + w % 4)) for w in range(n_windows)]
    fe.Process(frames[:B], outs); fe.Reset()
    StepMany(windows[:1], [1.8], 1e9)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    kp = m = 0
    outs2 = [outs, fe.alloc_outputs(pinned=True)] if a.pipelined else None
    if a.pipelined:                       # mage_frontend_submit / _wait: two calls in flight
        for k, i in enumerate(range(0, a.frames, B)):
            fe.Submit(frames[i:i + B], outs2[k & 1])
            if k:
                fe.Wait(); _, _, cnt, _, mc = fe.views(outs2[(k - 1) & 1]); kp += int(cnt.sum()); m += int(mc.sum())
        fe.Wait(); _, _, cnt, _, mc = fe.views(outs2[(a.frames // B - 1) & 1]); kp += int(cnt.sum()); m += int(mc.sum())
    else:
        for i in range(0, a.frames, B):
            kps, desc, cnt, mt, mc = fe.Process(frames[i:i + B], outs)
            kp += int(cnt.sum()); m += int(mc.sum())
    torch.cuda.synchronize()
    t_frames = time.perf_counter() - t0
    means = StepMany(windows, [1.8] * 10, 1e9)      # fresh windows: includes the host-side structure build of every window
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, t_frames], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, t_frames = float(t[0].item()), float(t[1].item())
    iters = sum(b.stats()["lm_iterations"] for b in windows) - 1
    if rank == 0:
        print(json.dumps({"config": "8 x 1280x720 sequences (one per GPU), ORB extract+match per frame + local BA every %d frames" % a.ba_every,
                          "n_gpus": world, "frames_per_gpu": a.frames, "frames_per_s": world * a.frames / dt, "ba_lm_iters_per_s": world * iters / dt,
                          "frontend_only_frames_per_s": world * a.frames / t_frames, "ba_only_lm_iters_per_s": world * iters / max(dt - t_frames, 1e-9),
                          "keypoints_per_frame": kp / a.frames, "matches_per_frame": m / a.frames, "ba_mean_sq_error": float(np.mean(means)),
                          "wall_s": dt, "note": "end to end through the host-buffer C ABI (pinned frames in, host results out)"}))
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
