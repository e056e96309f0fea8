"""Sharded global bundle adjustment: one problem spread over the ranks of a torch.distributed job (SURVEY.md 8(e) / 8(f) "next").

The landmarks -- with their observations -- are dealt out to the ranks (landmark i goes to rank i % world), every rank keeps ALL
cameras. Per Levenberg-Marquardt trial there is one real exchange step, the all-reduce of the reduced camera system S (n x n, n = 6 x
free cameras: 71 MB at the 500-key-frame tier size) and its right-hand side; three small all-reduces carry the robust chi2, the
diagonal for the initial damping and the gain-ratio denominator. Every rank then factorises the same S (the tcgen05 dense solver of
csrc/dense_ldlt.cuh) and back-substitutes its own landmarks. The accept / reject logic is g2o's
(ref Dependencies/g2o/g2o/core/optimization_algorithm_levenberg.cpp:57-174), run on the host from all-reduced -- hence bitwise equal --
numbers, so all ranks take the same decisions. Device work goes through the C ABI (mage_ba_shard_prepare / mage_ba_shard_stage);
torch is used for the collectives only (NCCL on the device buffers; gloo on host copies for CPU-side tests of this logic).

At the tier size one GPU solves the problem in 3.2 ms per step and two thirds of that is the dense factorisation, which every rank
repeats: sharding buys memory capacity (the observations and landmark blocks are divided by the number of ranks), not time."""
import ctypes as C
import math
import time

import numpy as np

from ._lib import check, lib
from .bundler import BundlerLib

LINEARIZE, SCHUR, SOLVE, RESTORE, FINISH = 1, 2, 3, 4, 5
STAGE_NAMES = {1: "linearize", 2: "schur", 3: "solve_update", 4: "restore", 5: "finish"}


def shard_problem(prob, rank, world):
    """The rank's part of a synth.ba_problem-shaped dict: all cameras, the points i with i % world == rank (renumbered in order) and
    their observations (in their original order). Returns (shard, global indices of the shard's points)."""
    P = len(prob["points"])
    mine = np.arange(rank, P, world)
    local = np.full(P, -1, np.int64); local[mine] = np.arange(len(mine))
    keep = local[np.asarray(prob["obs_pt"], np.int64)] >= 0
    out = {k: prob[k] for k in ("cam_pos", "cam_rot", "intrinsics", "fixed")}
    out["points"] = np.ascontiguousarray(np.asarray(prob["points"])[mine])
    out["obs_uv"] = np.ascontiguousarray(np.asarray(prob["obs_uv"])[keep])
    out["obs_cam"] = np.ascontiguousarray(np.asarray(prob["obs_cam"])[keep])
    out["obs_pt"] = np.ascontiguousarray(local[np.asarray(prob["obs_pt"], np.int64)[keep]].astype(np.int32))
    out["obs_info"] = np.ascontiguousarray(np.asarray(prob["obs_info"])[keep])
    return out, mine


def lm_decision(current_chi, temp_chi, scale, lam, ni):
    """One accept / reject step of g2o's Levenberg-Marquardt (ref optimization_algorithm_levenberg.cpp:116-147).
    Returns (accepted, rho, lambda, ni)."""
    rho = (current_chi - temp_chi) / (scale + 1e-3)
    if rho > 0 and math.isfinite(temp_chi):
        alpha = min(1.0 - (2.0 * rho - 1.0) ** 3, 2.0 / 3.0)
        return True, rho, lam * max(1.0 / 3.0, alpha), 2.0
    return False, rho, lam * ni, ni * 2.0


class _DeviceVector:
    """a device buffer of the library as something torch.as_tensor can alias (no copy)"""
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class _CudaShard:
    """The rank's shard on its GPU: the stages run in libmage_b200.so (k_ba_shard_stage), the exchanged buffers are the library's
    own device arrays aliased as torch tensors."""

    def __init__(self, shard):
        import torch
        self.ba = BundlerLib().load(shard)
        n = C.c_int(0); pS, pb, px = C.c_void_p(), C.c_void_p(), C.c_void_p()
        L = lib()
        L.mage_ba_shard_prepare.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.mage_ba_shard_stage.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]
        check(L.mage_ba_shard_prepare(self.ba._h, C.byref(n), C.byref(pS), C.byref(pb), C.byref(px)))
        self.n = n.value
        self.S = torch.as_tensor(_DeviceVector(pS.value, self.n * self.n), device="cuda")
        self.bs = torch.as_tensor(_DeviceVector(pb.value, self.n), device="cuda")
        self.xchg = torch.as_tensor(_DeviceVector(px.value, 16 + 2 * self.n), device="cuda")
        # S and its right-hand side are neighbours in the library's work arena (a few bytes of alignment padding at most between them):
        # one all-reduce over the span carries both
        gap = pb.value - pS.value - 8 * self.n * self.n
        self.system = torch.as_tensor(_DeviceVector(pS.value, self.n * self.n + gap // 8 + self.n), device="cuda") if 0 <= gap <= 4096 and gap % 8 == 0 else None
        self.sync = torch.cuda.synchronize

    def stage(self, stage, delta, lam, lead):
        check(lib().mage_ba_shard_stage(self.ba._h, stage, float(delta), float(lam), int(lead)))


class ShardedGlobalBA:
    """One rank's handle on a sharded problem. `dist` = the initialised torch.distributed module (None: a single rank).
    `backend` (tests of the host logic on CPU ranks over gloo) replaces the CUDA shard by an object with the same five stages and
    exchange buffers (n, S, bs, xchg, stage(), sync()); the product path never passes it."""

    def __init__(self, prob, rank=0, world=1, dist=None, user_lambda=-1.0, backend=None):
        import torch
        self.torch, self.dist, self.rank, self.world = torch, dist if world > 1 else None, rank, world
        if backend is None:
            shard, self.point_ids = shard_problem(prob, rank, world)
            backend = _CudaShard(shard)
            self.ba = backend.ba
        self.be = backend
        self.n, self.S, self.bs, self.xchg = backend.n, backend.S, backend.bs, backend.xchg
        self.system = getattr(backend, "system", None)
        # every rank must see the same free cameras (a camera without an observation in some shard would be left out there)
        nn = torch.tensor([self.n, -self.n], dtype=torch.int64, device=self.xchg.device)
        if self.dist is not None:
            self.dist.all_reduce(nn, op=self.dist.ReduceOp.MAX)
        if int(nn[0]) != self.n or int(-nn[1]) != self.n:
            raise ValueError("the shards disagree on the free cameras (reduced systems of %d .. %d unknowns): too few landmarks per camera for %d ranks" % (-int(nn[1]), int(nn[0]), world))
        self.lam, self.ni, self.iteration, self.user_lambda = -1.0, 2.0, 0, user_lambda
        self.trials = 0
        self.collectives = 0
        self.seconds = {}                 # host-clock seconds per stage / kind of exchange, accumulated (each stage call is synchronous)

    def _stage(self, stage, delta=0.0, lam=0.0):
        t0 = time.perf_counter()
        self.be.stage(stage, delta, lam, 1 if self.rank == 0 else 0)      # returns when the stage's kernel has finished
        self.seconds[STAGE_NAMES[stage]] = self.seconds.get(STAGE_NAMES[stage], 0.0) + time.perf_counter() - t0

    def _reduce(self, t, op):
        t0 = time.perf_counter()
        if self.dist is not None:
            self.dist.all_reduce(t, op=op)
            self.collectives += 1
        self.be.sync()
        name = "all_reduce_system" if t.numel() >= self.n * self.n else "all_reduce_small"
        self.seconds[name] = self.seconds.get(name, 0.0) + time.perf_counter() - t0

    def _sum(self, t):
        self._reduce(t, self.dist.ReduceOp.SUM if self.dist is not None else None)

    def _max(self, t):
        self._reduce(t, self.dist.ReduceOp.MAX if self.dist is not None else None)

    def StepBundleAdjustment(self, huber_widths):
        """LM iterations, one per Huber width (BundlerLib::StepBundleAdjustment without outlier removal). Returns the mean squared
        reprojection error over all ranks' observations."""
        n = self.n
        for delta in np.asarray(huber_widths, np.float32):      # the reference's widths are floats (ref BundlerLib.h:58)
            delta = float(delta)
            self._stage(LINEARIZE, delta)
            if self.iteration == 0:                       # the largest landmark diagonal entry is only needed for the initial damping
                self._max(self.xchg[1:2])
                md_land = float(self.xchg[1])
            self._sum(self.xchg)                          # chi2, diag(Hpp), bp in one vector (the slots in between are not read afterwards)
            current_chi = float(self.xchg[0])
            if self.iteration == 0:                       # ref :69-90 computeLambdaInit: tau * max diagonal entry of the Hessian
                md = max(md_land, float(self.xchg[16:16 + n].abs().max()))
                self.lam = self.user_lambda if self.user_lambda > 0 else 1e-5 * md
                self.ni = 2.0
            rho, qmax, finite = 0.0, 0, True
            while True:
                self._stage(SCHUR, delta, self.lam)
                if self.system is not None:
                    self._sum(self.system)
                else:
                    self._sum(self.S); self._sum(self.bs)
                self._stage(SOLVE, delta, self.lam)
                self._sum(self.xchg[0:3])
                tri = self.xchg[0:3].cpu().numpy()
                ok = int(round(float(tri[2]))) == self.world
                temp_chi = float(tri[0]) if ok else float(np.finfo(np.float64).max)
                accepted, rho, self.lam, self.ni = lm_decision(current_chi, temp_chi, float(tri[1]), self.lam, self.ni)
                self.trials += 1
                if accepted:
                    current_chi = temp_chi
                else:
                    self._stage(RESTORE)
                    if not math.isfinite(self.lam):
                        finite = False
                        break
                qmax += 1
                if not (rho < 0 and qmax < 10):
                    break
            self.iteration += 1
            if qmax == 10 or rho == 0 or not finite:
                break
        self._stage(FINISH, float(np.float32(huber_widths[-1])) if len(huber_widths) else 0.0)
        self._sum(self.xchg[0:2])
        tot = self.xchg[0:2].cpu().numpy()
        return float(tot[0] / max(tot[1], 1.0))

    def GetCurrentLambda(self):
        return self.lam

    def poses(self):
        return self.ba.poses()

    def points(self):
        """(global point indices of this shard, their positions)"""
        return self.point_ids, self.ba.points()
