"""The Fortran shim, the patched drivers' main loops and the whole drivers on top of the REAL libwuming_b200.so: the cases live in
tests/gpu_shim_cases.py; each group runs in a child process (python -m pytest on that file), so that the RTLD_GLOBAL binding of the
translated shim to the CUDA library stays out of this process and a crash there is one failed test here, not the end of the run.
Last file of the suite by name."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUPS = ["test_driver_through_the_shim_sync_every_call", "test_driver_through_the_shim_resident", "test_moments_through_the_shim",
          "test_patched_main_loop_on_the_gpu", "test_the_whole_weibel_driver_on_the_gpu", "test_the_whole_reconnection_driver_on_the_gpu"]


@pytest.mark.parametrize("group", GROUPS)
def test_gpu_shim_cases(group):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "gpu_shim_cases.py"), "-q", "-x", "-k", group,
                        "-p", "no:cacheprovider"], capture_output=True, text=True, cwd=ROOT, timeout=800)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-1500:])
    assert r.returncode == 0, f"{group}: exit code {r.returncode}\n{tail}"
    assert " passed" in r.stdout and " failed" not in r.stdout and " skipped" not in r.stdout.splitlines()[-1], tail
