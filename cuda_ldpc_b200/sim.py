"""SNR-point simulation loop over the C-ABI, one process per GPU.

Mirrors the reference's driver (bldpc_实习/main.cu:106-157, Simulation.cu:111-156, Statistic
Simulation.cu:245-285): per SNR point the counters are reset, batches of F frames are generated,
decoded and counted until `num_Error_Frames >= leastErrorFrames and num_Frames >= leastTestFrames`.
Differences: the channel runs on the device and is keyed by the GLOBAL frame index (so the counters
after a given number of frames do not depend on how many GPUs shared the work), and with
torch.distributed the frame range is sharded round-robin by batch over the ranks; the only exchange
is one all-reduce(sum) of the six int64 counters per round (48 bytes) to evaluate the stop rule.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import binary as B
from ._lib import LdpcError, lib

COUNTER_NAMES = ("num_Frames", "num_Error_Frames", "num_Error_Bits", "Total_Iteration", "num_False_Frames",
                 "num_Alarm_Frames")


@dataclass
class SimResult:
    snr_db: float
    counters: np.ndarray = field(default_factory=lambda: np.zeros(6, np.int64))
    length: int = 1
    seconds: float = 0.0

    def __getattr__(self, name):
        if name in COUNTER_NAMES:
            return int(self.counters[COUNTER_NAMES.index(name)])
        raise AttributeError(name)

    @property
    def FER(self):
        return self.num_Error_Frames / max(self.num_Frames, 1)

    @property
    def BER(self):
        return self.num_Error_Bits / max(self.num_Frames, 1) / self.length

    @property
    def AverageIT(self):
        return self.Total_Iteration / max(self.num_Frames, 1)

    def row(self):
        """the reference's result row (Simulation.cu:271)"""
        return (f" {self.snr_db:.1f} {self.num_Frames:8d}  {self.num_Error_Frames:4d}  {self.FER:6.4e}  {self.BER:6.4e}  "
                f"{self.AverageIT:.2f}  {self.num_False_Frames / max(self.num_Frames, 1):6.4e} "
                f"{self.num_Alarm_Frames / max(self.num_Frames, 1):6.4e}")


def batch_first_frame(round_idx, rank, world, batch):
    """global index of the first frame of the batch `rank` decodes in round `round_idx`"""
    return (round_idx * world + rank) * batch


class CudaBatchRunner:
    """channel + decode + statistic for one batch on the current CUDA device, all through the C-ABI"""

    def __init__(self, code, batch, *, maxit, schedule=B.SCHED_LAYERED, early_exit=B.EXIT_SYNDROME, msg_max=31,
                 llr_scale=8.0, beta_num=0, beta_shift=0, codeword=None, seed=173, msg_dtype=None):
        import torch
        self.torch, self.code, self.batch, self.maxit, self.seed = torch, code, batch, maxit, seed
        dev = torch.device("cuda", torch.cuda.current_device())
        # flooding fp32 reads a channel buffer; the layered int8 decoder generates the channel itself
        self.fused = schedule == B.SCHED_LAYERED
        self.y = None if self.fused else torch.empty(code.N * batch, dtype=torch.float32, device=dev)
        self.out = torch.empty(code.out_bytes(batch, B.OUT_BITPACK), dtype=torch.uint8, device=dev)
        self.iters = torch.empty(batch, dtype=torch.int32, device=dev)
        self.ok = torch.empty(batch, dtype=torch.int32, device=dev)
        self.cnt = torch.zeros(6, dtype=torch.int64, device=dev)
        self.cw = None if codeword is None else torch.as_tensor(np.asarray(codeword, np.uint8), device=dev)
        self.kw = dict(schedule=schedule, early_exit=early_exit, msg_max=msg_max, llr_scale=llr_scale,
                       beta_num=beta_num, beta_shift=beta_shift, out_format=B.OUT_BITPACK)
        if msg_dtype is not None:  # layered: B.DTYPE_INT8 (default) or B.DTYPE_FP16
            self.kw["msg_dtype"] = msg_dtype

    def reset(self):
        self.cnt.zero_()

    def run(self, sigma, first_frame):
        st = self.torch.cuda.current_stream().cuda_stream
        cw = self.cw.data_ptr() if self.cw is not None else None
        if self.fused:
            kw = {k: v for k, v in self.kw.items() if k != "schedule"}
            self.code.decode_channel(self.batch, self.maxit, sigma, seed=self.seed, first_frame=first_frame,
                                     codeword=self.cw, out=self.out, iters_out=self.iters, ok_out=self.ok, **kw)
        else:
            rc = lib.ldpc_awgn_bpsk(self.code.handle, self.y.data_ptr(), self.batch, B.LAYOUT_NF, float(sigma),
                                    int(self.seed), int(first_frame), cw, st)
            if rc < 0:
                raise LdpcError(rc, "ldpc_awgn_bpsk")
            self.code.decode(self.y.view(self.code.N, self.batch), self.maxit, out=self.out, iters_out=self.iters,
                             ok_out=self.ok, **self.kw)
        rc = lib.ldpc_statistic(self.code.handle, self.out.data_ptr(), B.OUT_BITPACK, self.ok.data_ptr(),
                                self.iters.data_ptr(), self.batch, self.code.K, cw, self.cnt.data_ptr(), st)
        if rc < 0:
            raise LdpcError(rc, "ldpc_statistic")

    def counters(self):
        return self.cnt.cpu().numpy().astype(np.int64)


def run_snr_point(runner, snr_db, sigma, *, least_errors=50, least_frames=10000, max_frames=0, rank=0, world=1,
                  all_reduce=None, length=1):
    """One SNR point.  `runner` has reset() / run(sigma, first_frame) / counters() -> int64[6] (local,
    cumulative).  all_reduce(np.int64[6]) -> summed over ranks (identity for world == 1)."""
    import time
    runner.reset()
    res = SimResult(snr_db=snr_db, length=length)
    t0 = time.perf_counter()
    r = 0
    while True:
        runner.run(sigma, batch_first_frame(r, rank, world, runner.batch))
        r += 1
        total = runner.counters()
        if all_reduce is not None:
            total = all_reduce(total)
        res.counters = total
        if (res.num_Error_Frames >= least_errors and res.num_Frames >= least_frames) or \
                (max_frames and res.num_Frames >= max_frames):
            break
    res.seconds = time.perf_counter() - t0
    return res


def torch_all_reduce(group=None):
    """all-reduce(sum) of the six counters over torch.distributed (NCCL on GPUs, gloo in CPU tests)"""
    import torch
    import torch.distributed as dist

    def fn(c):
        backend = dist.get_backend(group)
        t = torch.as_tensor(c, dtype=torch.int64)
        if backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy().astype(np.int64)
    return fn
