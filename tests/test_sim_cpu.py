"""CPU tests of the host-side simulation loop: Statistic semantics, stop rule, and the multi-rank
path (world_size 2, gloo): frame sharding by global frame index + counter all-reduce give exactly
the single-rank counters.  The decoder / channel stand-ins here are the CPU oracle (test code may
use it; the product path never does)."""
import os
import sys

import numpy as np
import pytest

from conftest import DATA, ROOT, OracleCode, ip, fp

from cuda_ldpc_b200 import sim

BL = os.path.join(DATA, "bldpc")


class OracleRunner:
    """Philox-free stand-in: noise from numpy keyed by the global frame index, decode + statistic
    through the oracle."""

    def __init__(self, lib, code, batch, maxit):
        self.lib, self.code, self.batch, self.maxit = lib, code, batch, maxit
        self.cnt = np.zeros(6, np.int64)

    def reset(self):
        self.cnt[:] = 0

    def run(self, sigma, first_frame):
        c, F = self.code, self.batch
        y = np.empty((c.N, F), np.float32)
        for f in range(F):  # one generator per GLOBAL frame: independent of the sharding
            y[:, f] = 1.0 + sigma * np.random.default_rng(1000 + first_frame + f).standard_normal(c.N)
        D = np.zeros((c.N + 1) * F, np.int32)
        it = np.zeros(F, np.int32)
        assert self.lib.orc_layered_i8(c.J, c.L, c.Z, ip(c.H), fp(np.ascontiguousarray(y)), F, self.maxit, 8.0, 31,
                                       0, 0, 2, ip(D), ip(it), None, None) == 0
        Dm = D.reshape(c.N + 1, F)
        err = Dm[: c.K].sum(0)
        flag = Dm[c.N]
        self.cnt += np.array([F, ((err != 0) | (flag == 0)).sum(), err.sum(), it.sum(),
                              ((err != 0) & (flag == 1)).sum(), ((err == 0) & (flag == 0)).sum()], np.int64)

    def counters(self):
        return self.cnt.copy()


def test_statistic_semantics_match_oracle_statistic(oracle):
    """the counter formulas used by the runners = B/Simulation.cu:245-285 as restated in the oracle"""
    import ctypes as C
    rng = np.random.default_rng(3)
    N, K, F = 40, 24, 64
    D = np.zeros((N + 1, F), np.int32)
    D[:N] = rng.random((N, F)) < 0.02
    D[N] = rng.random(F) < 0.7
    it = rng.integers(1, 11, F).astype(np.int32)

    class Cn(C.Structure):
        _fields_ = [(n, C.c_long) for n in sim.COUNTER_NAMES]
    c = Cn()
    c.num_Frames = F
    oracle.orc_statistic(C.byref(c), None, ip(np.ascontiguousarray(D)), ip(it), N, F, K, 50, 100)
    err = D[:K].sum(0)
    want = [F, ((err != 0) | (D[N] == 0)).sum(), err.sum(), it.sum(), ((err != 0) & (D[N] == 1)).sum(),
            ((err == 0) & (D[N] == 0)).sum()]
    assert [getattr(c, n) for n in sim.COUNTER_NAMES] == [int(x) for x in want]


def test_single_rank_loop_and_stop_rule(oracle):
    code = OracleCode(oracle, os.path.join(BL, "J4_L24_Z96_BlockH.txt"), 4, 24, 96)
    r = OracleRunner(oracle, code, 32, 10)
    res = sim.run_snr_point(r, 2.0, 0.63, least_errors=5, least_frames=64, length=code.K)
    assert res.num_Frames % 32 == 0 and res.num_Frames >= 64 and res.num_Error_Frames >= 5
    assert 0 < res.FER <= 1 and res.row().startswith(" 2.0")
    res2 = sim.run_snr_point(r, 9.0, 0.25, least_errors=5, least_frames=64, max_frames=96, length=code.K)
    assert res2.num_Frames == 96 and res2.num_Error_Frames == 0  # stopped by max_frames only
    assert sim.batch_first_frame(3, 1, 4, 100) == 1300


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch.distributed as dist
    from conftest import OracleCode
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    lib.orc_layered_i8.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int, C.c_int,
                                                    C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                                    C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    code = OracleCode(lib, os.path.join(BL, "J4_L24_Z96_BlockH.txt"), 4, 24, 96)
    r = OracleRunner(lib, code, 16, 10)
    res = sim.run_snr_point(r, 2.0, 0.63, least_errors=10 ** 9, least_frames=0, max_frames=128, rank=rank, world=world,
                            all_reduce=sim.torch_all_reduce(), length=code.K)
    if rank == 0:
        q.put(res.counters.tolist())
    dist.destroy_process_group()


def test_two_ranks_gloo_equal_single_rank(oracle):
    """128 frames decoded by 1 rank and by 2 ranks (gloo): identical counters."""
    import torch.multiprocessing as mp
    code = OracleCode(oracle, os.path.join(BL, "J4_L24_Z96_BlockH.txt"), 4, 24, 96)
    single = sim.run_snr_point(OracleRunner(oracle, code, 16, 10), 2.0, 0.63, least_errors=10 ** 9, least_frames=0,
                               max_frames=128, length=code.K)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == single.counters.tolist()
    assert got[0] == 128 and got[1] > 0
