"""The reciprocal-table quotient of avg_distance_tile_kernel is the IEEE quotient (CPU, exhaustive to b = 4096)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(shutil.which("gcc") is None or not _cpu_has_fma(), reason="needs gcc and a host CPU with FMA")
def test_fma_corrected_reciprocal_quotient_is_correctly_rounded(tmp_path):
    exe = str(tmp_path / "check_recip_division")
    src = os.path.join(ROOT, "tools", "check_recip_division.c")
    subprocess.run(["gcc", "-O2", "-mfma", "-o", exe, src, "-lm"], check=True)
    out = subprocess.run([exe, "4096"], check=True, capture_output=True, text=True).stdout
    assert "0 mismatches" in out, out
