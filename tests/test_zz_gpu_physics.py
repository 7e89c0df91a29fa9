"""The CUDA path against an analytic answer (no oracle in the comparison): a cold pair plasma with omega_pe = omega_pi = 0.1
oscillates at 0.1 sqrt(2) -- through wm_step (fused kernel, on-device CG, lazy sort) on the device-resident state."""
import numpy as np
import pytest

from oracle.pyoracle import World3, weibel_constants
from tests.util import active_mask, backend_for, upload_from_world

pytestmark = pytest.mark.gpu


def test_gpu_cold_plasma_oscillates_at_the_plasma_frequency():
    n0, nx, ny, nz = 8, 32, 4, 4
    q, r, _ = weibel_constants(n0)
    w = World3(nx, ny, nz, n0 * nx * 3, q=q, r=r)          # the oracle world only builds the initial state
    w.load_weibel(n0, v_thi=0.0, v_the=0.0, t_ani=1.0)
    up, m = w.arr("up"), active_mask(w.arr("np2"), w.np)
    amp = 1e-3
    x = up[..., 0]
    up[1, ..., 3] = np.where(m[1], amp * np.sin(2 * np.pi * (x[1] - 2) / nx), 0.0)
    up[0, ..., 3] = np.where(m[0], -amp * np.sin(2 * np.pi * (x[0] - 2) / nx), 0.0)
    b = backend_for(w)
    upload_from_world(b, w)
    e_field = []
    for _ in range(200):
        b.step(2, nx + 1, 1)
        e_field.append(b.energy()[2])
    assert b.stats()["error_flags"] == 0
    e = np.array(e_field)
    f = np.abs(np.fft.rfft(e - e.mean()))
    k = int(np.argmax(f[1:])) + 1
    kk = k + 0.5 * (f[k - 1] - f[k + 1]) / (f[k - 1] - 2 * f[k] + f[k + 1])
    omega = 2 * np.pi * kk / len(e) / 2
    assert abs(omega - 0.1 * np.sqrt(2)) < 0.01 * 0.1 * np.sqrt(2), omega
    b.close(); w.close()
