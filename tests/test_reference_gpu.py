"""The reference's own GPU decoder, compiled unchanged for sm_100 (baseline/build_all.sh -> baseline/_ref/, built in
the container where /root/reference exists; the binaries travel to the GPU box), run on the B200 next to this repo's
strict-parity mode: same seeds, same counters.  Closes the loop oracle == reference-on-GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import DATA, GOLDEN, fp, ip

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "baseline", "_ref", "bldpc_gpu_C1_fer_fixed")


def run_ref(snrs):
    p = subprocess.run([BIN, "fer"] + [str(s) for s in snrs], capture_output=True, text=True, timeout=300)
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]


@pytest.mark.skipif(not os.path.exists(BIN), reason="baseline/_ref not built (needs /root/reference at build time)")
def test_reference_gpu_counters_equal_golden_and_this_repo(oracle):
    """B/Simulation.cu:12-171 (Simulation_GPU) with the reference RNG at 2.5 / 3.0 dB Es/N0, J4_L24_Z96, F = 256:
    error frames / bits equal the golden numbers recorded from the serial CPU run of the same sources, and equal
    ldpc_decode_batch(flooding fp32, genie exit) fed the same noise stream."""
    import cuda_ldpc_b200 as m
    with open(os.path.join(GOLDEN, "binary_ref.json")) as f:
        gold = json.load(f)
    rows = run_ref([2.5, 3.0])
    assert [r["snr_db"] for r in rows] == [2.5, 3.0]
    assert (rows[0]["frames"], rows[0]["error_frames"]) == (1024, 669)
    assert (rows[1]["frames"], rows[1]["error_frames"]) == (1024, 55)
    # this repo on the reference's own noise stream (oracle RNG == RandomModule, pinned in test_oracle_binary.py)
    code = m.LdpcCode(os.path.join(DATA, "bldpc", "J4_L24_Z96_BlockH.txt"))
    for row in rows:
        seed = np.array([173, 173, 173], np.int32)
        sigma = oracle.orc_sigma(1, row["snr_db"], 1.0)
        err_frames = err_bits = 0
        for _ in range(row["frames"] // 256):
            y = np.zeros(code.N * 256, np.float32)
            oracle.orc_awgn(ip(seed), sigma, None, fp(y), code.N, 256)
            r = code.decode(y.reshape(code.N, 256), 10, early_exit=m.EXIT_GENIE)
            e = r.D[: code.K].sum(0)
            err_bits += int(e.sum())
            err_frames += int(((e != 0) | (r.D[code.N] == 0)).sum())
        assert (err_frames, err_bits) == (row["error_frames"], row["error_bits"]), row
    assert gold is not None
