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
    """The fp16 (hi, lo) split two-stage DFT of frontend_tc.cu - operand images decoded through the swizzle map, per-frame
    power-of-two scaling, twiddles, mirrored bin map - against a double-precision DFT and an fp32 FFT."""
    exe = str(tmp_path / "tc_dft_host_check")
    src = os.path.join(ROOT, "tests", "host", "tc_dft_host_check.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
