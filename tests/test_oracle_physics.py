"""Physics known-answer tests of the oracle: results that follow from the equations the reference solves, not from its code.
The reference ships no golden vectors for this path (the pin to its own translated source is
tests/test_ref_transpiled.py); these tie the normalisation of the whole
loop -- the 4 pi factors of the field update, q = sqrt(m / (4 pi n0)) omega_p of the loaders (3d/proj/weibel/app.f90:298-305),
the deposit, the push -- to analytic values."""
import numpy as np

from oracle.pyoracle import World3, weibel_constants
from tests.util import active_mask


def test_cold_plasma_oscillates_at_the_plasma_frequency():
    """A cold pair plasma (mass ratio 1, omega_pe = omega_pi = 0.1) with a small sinusoidal counter-streaming perturbation
    oscillates at sqrt(omega_pe^2 + omega_pi^2) = 0.1 sqrt(2); the field energy goes as sin^2, i.e. at twice that."""
    n0, nx, ny, nz = 8, 32, 2, 2
    q, r, _ = weibel_constants(n0)
    w = World3(nx, ny, nz, n0 * nx * 3, q=q, r=r)
    w.load_weibel(n0, v_thi=0.0, v_the=0.0, t_ani=1.0)
    up, m = w.arr("up"), active_mask(w.arr("np2"), w.np)
    amp = 1e-3
    x = up[..., 0]
    up[1, ..., 3] = np.where(m[1], amp * np.sin(2 * np.pi * (x[1] - 2) / nx), 0.0)
    up[0, ..., 3] = np.where(m[0], -amp * np.sin(2 * np.pi * (x[0] - 2) / nx), 0.0)
    w.arr("gp")[...] = up
    e_field = []
    for _ in range(200):
        w.step()
        e_field.append(w.energy()[2])
    assert w.error() == 0
    e = np.array(e_field)
    f = np.abs(np.fft.rfft(e - e.mean()))
    k = int(np.argmax(f[1:])) + 1
    kk = k + 0.5 * (f[k - 1] - f[k + 1]) / (f[k - 1] - 2 * f[k] + f[k + 1])      # parabolic peak interpolation
    omega = 2 * np.pi * kk / len(e) / 2
    assert abs(omega - 0.1 * np.sqrt(2)) < 0.01 * 0.1 * np.sqrt(2), omega          # measured: 0.14133 against 0.14142
    w.close()


def test_gyration_angle_per_step_is_the_boris_angle():
    """One particle in a uniform B, E = 0: every Buneman-Boris step turns the momentum by 2 atan(q B dt / (2 gamma m c)) about B
    (the Vay update has the same rotation when E = 0); |u| and u_parallel stay put."""
    q, r, _ = weibel_constants(1)
    bz = 3.7
    for vay in (False, True):
        w = World3(8, 8, 8, 8 * 3, q=q, r=r)
        uf = w.arr("uf")
        uf[...] = 0.0
        uf[..., 2] = bz
        up, np2, cc = w.arr("up"), w.arr("np2"), w.arr("cumcnt")
        np2[...] = 0
        cc[...] = 0
        np2[1, 3, 3] = 1
        cc[1, 3, 3, 4:] = 1
        u0 = np.array([0.3, -0.2, 0.15])
        up[1, 3, 3, 0, :6] = [5.5, 5.5, 5.5, *u0]
        gam = np.sqrt(1 + u0 @ u0 / w.c ** 2)
        (w.particle_solv_vay if vay else w.particle_solv)()
        u1 = w.arr("gp")[1, 3, 3, 0, 3:6]
        ang = np.arctan2(u1[1], u1[0]) - np.arctan2(u0[1], u0[0])
        expect = -2 * np.arctan(q[1] * bz * w.delt / (2 * gam * r[1] * w.c))       # dphi/dt = -q B / (gamma m c)
        assert abs(ang - expect) < 1e-14, (vay, ang, expect)
        assert abs(u1[2] - u0[2]) < 1e-16 and abs(np.hypot(u1[0], u1[1]) - np.hypot(u0[0], u0[1])) < 1e-15
        w.close()
