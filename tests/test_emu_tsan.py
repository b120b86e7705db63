"""Race check of the kernel sources: the thread-per-CUDA-thread emulation (std::thread + std::barrier at every
__syncthreads / warp barrier) built with -fsanitize=thread runs one small problem through the forward solve, both backward
sweeps and the forward sweep; ThreadSanitizer must stay silent.  On the GPU k_riccati_bdf relies on warp barriers only
(one warp per problem), so a missing barrier would be a latent bug there."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model", ["pendulum", "quadrotor"])
def test_emulated_kernels_are_race_free(model):
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not tsan or not os.path.exists(tsan):
        pytest.skip("libtsan not available")
    env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="report_bugs=1 exitcode=0 halt_on_error=0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "tsan_run.py"), model], capture_output=True, text=True,
                       env=env, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    assert out.count("%s mode" % model) == 2, out[-2000:]
    assert "WARNING: ThreadSanitizer" not in out, out[:4000]
