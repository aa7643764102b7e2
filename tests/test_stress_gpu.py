"""Race hunting without compute-sanitizer (it is closed on the GPU pool): the same library compiled with
-DLDPC_STRESS=1 (cuda_ldpc_b200/csrc/stress.cuh) delays a pseudo-random quarter of the warps by up to ~4 us after
every CTA barrier, so an access that relies on "the other warps cannot be that far ahead" goes wrong within a few
groups.  The worker compares every output (hard bits, flags, iteration counts, APP values, check records) with the
CPU oracle over several repeats.

The canary proves the method: libldpc_b200_stress_r1.so re-opens the two shared-memory races round 1 shipped
(-DLDPC_R1_RACES=1) and MUST be caught.
Reference: the stop test the raced code replaces is B/LDPC_Decoder.cu:134-153."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda_ldpc_b200")


def run_worker(libname, what):
    if not os.path.exists(os.path.join(PKG, libname)):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(PKG, "csrc"), "stress"])
    env = dict(os.environ, LDPC_B200_LIB=libname)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "stress_worker.py"), what], env=env,
                       capture_output=True, text=True, timeout=900)
    return p.returncode, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("what", ["binary", "nb"])
def test_stress_build_is_bit_exact(what):
    rc, log = run_worker("libldpc_b200_stress.so", what)
    assert rc == 0, log


def test_stress_build_catches_the_round1_races():
    """Detection is probabilistic by nature (measured rate: profiles/r02_stress_canary.txt), so a miss is reported
    as a skip, not as a failure of the product."""
    try:
        rc, log = run_worker("libldpc_b200_stress_r1.so", "binary")
    except subprocess.TimeoutExpired:
        return  # CTA barriers out of step can also hang the raced build: caught
    if rc == 0:
        pytest.skip("the canary build with the round-1 races went undetected in this run")
    assert rc == 3 and "MISMATCH" in log, log
