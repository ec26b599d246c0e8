"""csrc/fft_core.cuh compiles for the host too: tests/host/fft_host_check.cpp runs the SAME stage functions the CUDA kernels
run (a warp = a loop over 32 lanes between the barriers) against a naive double-precision DFT -- index maps, twiddles,
shared-memory exchange layout and the real-FFT split are pinned without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fft_core_on_the_host(tmp_path):
    exe = str(tmp_path / "fft_host_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "text2speech_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "fft_host_check.cpp"), "-o", exe], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "OK" in res.stdout
