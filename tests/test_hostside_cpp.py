"""The compiled host-side mirror (hostside/wuming_b200.hpp: the reference's module procedures in C++, the language class of the
reference's host code) and its example driver (hostside/weibel3d_main.cpp = app__main of 3d/proj/weibel/app.f90:88-160)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "hostside", "weibel3d_main")


def _build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "hostside")], check=True)


def test_driver_builds_and_fails_loudly_without_a_gpu():
    import torch
    _build()
    assert os.path.exists(EXE)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, "8", "4", "4", "2", "1", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1                                  # the reference's "message; stop"
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def _energies(out):
    rows = [list(map(float, l.split())) for l in out.splitlines() if l.strip() and not l.startswith("#")]
    return np.array(rows)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [0, 1], ids=["five-calls", "wm_step"])
def test_driver_energy_history_matches_the_python_mirror(fused):
    """the C++ driver and the Python mirror run the same library on the same device-generated load: same energy history"""
    import wumingpic_b200 as wm
    _build()
    nx, ny, nz, n0, steps, cadence = 16, 8, 6, 6, 6, 2
    r = subprocess.run([EXE, str(nx), str(ny), str(nz), str(n0), str(steps), str(cadence), str(fused)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    got = _energies(r.stdout)
    q, rr, _ = wm.weibel_constants(n0)
    b = wm.Backend(3, 3 * n0 * nx, 2, nx + 1, 2, ny + 1, 2, nz + 1, q=q, r=rr)
    b.load_weibel(n0)
    b.set_fused(bool(fused))
    ref = [np.concatenate([[0.0], b.energy()])]
    for it in range(1, steps + 1):
        b.step(2, nx + 1, 1)
        if it % cadence == 0:
            ref.append(np.concatenate([[float(it)], b.energy()]))
    ref = np.array(ref)
    assert got.shape == (len(ref), 6)
    assert np.allclose(got[:, 0], ref[:, 0])
    assert np.allclose(got[:, 1:5], ref[:, 1:5], rtol=2e-5)          # printed with 6 significant digits (e12.5)
    assert np.allclose(got[:, 5], ref[:, 1:5].sum(axis=1), rtol=2e-5)
    b.close()
