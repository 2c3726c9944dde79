"""CPU test: the frontend's per-frame arithmetic (the same __host__ __device__ code the kernel runs) emulated
lane by lane on the host and checked against a double-precision DFT."""
import os
import subprocess

import pytest

from tests.util import ROOT


@pytest.mark.parametrize("n_mels", [64, 80])
def test_frame_pipeline_on_host(tmp_path, n_mels):
    exe = str(tmp_path / "fft_host_check")
    src = os.path.join(ROOT, "tests", "host", "fft_host_check.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
    out = subprocess.run([exe, str(n_mels)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "worst_power_rel" in out.stdout


def test_tensor_core_dft_arithmetic_on_host(tmp_path):
    """The TF32 (hi, lo) split two-stage DFT of frontend_tc.cu - operand images decoded through the swizzle map, the pieces as
    the tensor core reads them (13 low mantissa bits ignored), twiddles, mirrored bin map - against a double-precision DFT and
    an fp32 FFT."""
    exe = str(tmp_path / "tc_dft_host_check")
    src = os.path.join(ROOT, "tests", "host", "tc_dft_host_check.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_bench_warmup_stop_is_collective():
    """bench.py warms its end-to-end loops until two consecutive runs agree within 3 %.  The loop body holds barriers, so with
    several ranks the decision must come from the REDUCED time: every rank then runs the same number of times (a per-rank
    decision hung the 8-GPU config-3 run: one rank left the loop a barrier early)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    # two "ranks" whose own times would stop after different numbers of warm runs
    series = {0: [1.00, 0.99, 0.90, 0.895, 0.894, 0.894, 0.894, 0.894, 0.894, 0.894, 0.894, 0.894],
              1: [1.00, 0.90, 0.80, 0.700, 0.699, 0.699, 0.699, 0.699, 0.699, 0.699, 0.699, 0.699]}
    calls = {0: 0, 1: 0}

    def make_run(rank):
        def run():
            calls[rank] += 1
            return series[rank][calls[rank] - 1]
        return run

    own = [bench.stable_repeats(make_run(r), 2)[2] for r in (0, 1)]
    assert own[0] != own[1]                                  # the hazard: per-rank decisions differ
    calls.update({0: 0, 1: 0})

    def agree_for(rank):                                     # stands in for ctx.reduce_max: the max over the ranks' current run
        def agree(t):
            i = calls[rank] - 1
            return max(series[0][i], series[1][i])
        return agree

    used = [bench.stable_repeats(make_run(r), 2, agree=agree_for(r))[2] for r in (0, 1)]
    assert used[0] == used[1] and calls[0] == calls[1]
