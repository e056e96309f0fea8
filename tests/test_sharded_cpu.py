"""CPU tests of the sharded global-BA host logic (mageslam_b200/sharded.py): the partition of the problem, g2o's accept / reject step,
and the exchange choreography of the LM loop run by two gloo ranks -- the device stages replaced by a small numpy stand-in with the
same five stages and exchange buffers (a linear least-squares problem with camera-like and landmark-like unknowns), so that the
collectives, their order and the decisions taken from all-reduced numbers are what is exercised. The real stages (k_ba_shard_stage)
are checked on the GPU in tests/test_sharded_gpu.py."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mageslam_b200 import synth
from mageslam_b200.sharded import FINISH, LINEARIZE, RESTORE, SCHUR, SOLVE, ShardedGlobalBA, lm_decision, shard_problem


def test_shard_problem_is_a_partition():
    prob = synth.ba_problem(K=12, P=300, obs_per_point=4, seed=5)
    for world in (1, 2, 3, 8):
        seen_obs, seen_pts = 0, []
        for rank in range(world):
            sh, ids = shard_problem(prob, rank, world)
            assert np.array_equal(sh["cam_pos"], prob["cam_pos"]) and np.array_equal(sh["fixed"], prob["fixed"])      # all cameras everywhere
            assert np.array_equal(sh["points"], np.asarray(prob["points"])[ids])
            assert sh["obs_pt"].min() >= 0 and sh["obs_pt"].max() < len(ids)
            # the shard's observations are the original ones of its points, in their original order
            keep = np.isin(np.asarray(prob["obs_pt"]), ids)
            assert np.array_equal(sh["obs_uv"], np.asarray(prob["obs_uv"])[keep])
            assert np.array_equal(ids[sh["obs_pt"]], np.asarray(prob["obs_pt"])[keep])
            seen_obs += len(sh["obs_pt"]); seen_pts += list(ids)
        assert seen_obs == len(prob["obs_pt"]) and sorted(seen_pts) == list(range(len(prob["points"])))


def test_lm_decision_follows_g2o():
    # ref optimization_algorithm_levenberg.cpp:116-147: accepted -> lambda *= max(1/3, 1 - (2 rho - 1)^3) capped at 2/3, ni = 2
    ok, rho, lam, ni = lm_decision(10.0, 4.0, 6.0 - 1e-3, 3.0, 8.0)
    assert ok and abs(rho - 1.0) < 1e-12 and abs(lam - 1.0) < 1e-12 and ni == 2.0
    ok, rho, lam, ni = lm_decision(10.0, 9.0, 2.0 - 1e-3, 3.0, 2.0)       # rho = 0.5 -> alpha = 1 -> min(1, 2/3)
    assert ok and abs(lam - 2.0) < 1e-12
    ok, rho, lam, ni = lm_decision(10.0, 11.0, 1.0, 3.0, 4.0)             # worse: rejected, lambda *= ni, ni *= 2
    assert not ok and rho < 0 and lam == 12.0 and ni == 8.0
    ok, _, lam, _ = lm_decision(10.0, float("inf"), 1.0, 3.0, 2.0)        # a failed solve never passes
    assert not ok and lam == 6.0
    ok, _, _, _ = lm_decision(10.0, float("nan"), 1.0, 3.0, 2.0)
    assert not ok


class ToyShard:
    """numpy stand-in of a rank's shard: residuals r_k = a_k . c + g_k l_j(k) - y_k over n camera-like unknowns c (shared by all
    ranks) and this rank's scalar landmarks l; the same stages and exchange buffers as k_ba_shard_stage."""

    def __init__(self, n, A, g, lidx, y, n_land):
        self.n, self.A, self.g, self.lidx, self.y = n, A, g, lidx, y
        self.c, self.l = np.zeros(n), np.zeros(n_land)
        self.S, self.bs, self.xchg = torch.zeros(n * n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64), torch.zeros(16 + 2 * n, dtype=torch.float64)
        self.stages = []

    def sync(self):
        pass

    def _res(self):
        return self.A @ self.c + self.g * self.l[self.lidx] - self.y

    def stage(self, stage, delta, lam, lead):
        n, X = self.n, self.xchg.numpy()
        self.stages.append(stage)
        if stage == LINEARIZE:
            r = self._res()
            self.Hpp, self.bp = self.A.T @ self.A, -self.A.T @ r
            self.Hll = np.bincount(self.lidx, self.g * self.g, len(self.l)); self.bl = -np.bincount(self.lidx, self.g * r, len(self.l))
            self.W = np.zeros((len(self.l), n)); np.add.at(self.W, self.lidx, self.A * self.g[:, None])
            X[0], X[1], X[16:16 + n], X[16 + n:16 + 2 * n] = r @ r, self.Hll.max(), np.diag(self.Hpp), self.bp
        elif stage == SCHUR:
            self.bak = (self.c.copy(), self.l.copy())
            self.Dinv = 1.0 / (self.Hll + lam)
            self.S.numpy()[:] = (self.Hpp - self.W.T @ (self.Dinv[:, None] * self.W)).ravel()
            self.bs.numpy()[:] = self.bp - self.W.T @ (self.Dinv * self.bl)
        elif stage == SOLVE:
            S = self.S.numpy().reshape(n, n) + lam * np.eye(n)
            xc = np.linalg.solve(S, self.bs.numpy())
            xl = self.Dinv * (self.bl - self.W @ xc)
            self.c, self.l = self.c + xc, self.l + xl
            r = self._res()
            X[0], X[1], X[2] = r @ r, xl @ (lam * xl + self.bl) + (xc @ (lam * xc + X[16 + n:16 + 2 * n]) if lead else 0.0), 1.0
        elif stage == RESTORE:
            self.c, self.l = self.bak[0].copy(), self.bak[1].copy()
        elif stage == FINISH:
            r = self._res()
            X[0], X[1] = r @ r, len(r)


TOY_STEPS = 2          # past convergence the accept / reject decisions hang on rounding noise: compare the well-determined steps


def toy_problem(seed=3, n=6, L=40, per=5):
    rng = np.random.default_rng(seed)
    lidx = np.repeat(np.arange(L), per)
    A, g = rng.normal(size=(L * per, n)), rng.normal(size=L * per) + 2.0
    c_true, l_true = rng.normal(size=n), rng.normal(size=L)
    y = A @ c_true + g * l_true[lidx] + 1e-3 * rng.normal(size=L * per)
    return n, A, g, lidx, y, L


def toy_shard(rank, world):
    n, A, g, lidx, y, L = toy_problem()
    mine = np.arange(rank, L, world)
    local = np.full(L, -1); local[mine] = np.arange(len(mine))
    keep = local[lidx] >= 0
    return ToyShard(n, A[keep], g[keep], local[lidx[keep]], y[keep], len(mine)), mine


def run_toy(rank, world, d):
    be, mine = toy_shard(rank, world)
    sh = ShardedGlobalBA(None, rank, world, d, backend=be)
    means = [sh.StepBundleAdjustment([1.8]) for _ in range(TOY_STEPS)]
    return sh, be, mine, means


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh, be, mine, means = run_toy(rank, world, dist)
    q.put((rank, means, sh.lam, sh.trials, sh.collectives, be.c.tolist(), mine.tolist(), be.l.tolist(), be.stages))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_lm_loop_equals_one_rank_gloo():
    sh1, be1, _, means1 = run_toy(0, 1, None)
    # the single-rank loop solves the toy problem (linear: LM converges to the least-squares solution)
    n, A, g, lidx, y, L = toy_problem()
    J = np.zeros((len(y), n + L)); J[:, :n] = A; J[np.arange(len(y)), n + lidx] = g
    best = np.linalg.lstsq(J, y, rcond=None)[0]
    assert means1[-1] <= means1[0] and np.allclose(be1.c, best[:n], atol=1e-3)
    world, port = 2, 29617
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res
    # both ranks took the same decisions from the all-reduced numbers: same means, lambda, trial count, stage sequence, camera unknowns
    assert r0[1] == r1[1] and r0[2] == r1[2] and r0[3] == r1[3] and r0[8] == r1[8] and r0[5] == r1[5]
    # ... and they are the single-rank run's up to the summation order of the all-reduce
    assert np.allclose(r0[1], means1, rtol=1e-9) and math.isclose(r0[2], sh1.lam, rel_tol=1e-6) and r0[3] == sh1.trials
    assert np.allclose(r0[5], be1.c, rtol=1e-8, atol=1e-10)
    lands = np.zeros(L); lands[r0[6]] = r0[7]; lands[r1[6]] = r1[7]
    assert np.allclose(lands, be1.l, rtol=1e-8, atol=1e-10)
    # exchange steps: the MAX for the initial damping once, per LM iteration one small all-reduce after LINEARIZE, per lambda trial S + bs + the 3 trial scalars, one at the end of a
    # call (the free-camera check at construction is not counted)
    steps = TOY_STEPS
    assert r0[4] == 1 + steps + 3 * r0[3] + steps
