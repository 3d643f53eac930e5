"""A short randomised differential run of the CUDA convert+scale path against the oracle (tools/gpu_fuzz.py):
random sizes, methods, input/output formats, colorimetry and chroma siting, a different slice of the space per seed."""
import subprocess
import sys
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_configurations_match_oracle(cuda_device, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_fuzz.py"), "6", str(seed)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout
