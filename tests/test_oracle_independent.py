"""A second, independent restatement of the two particle kernels of the reference -- written in numpy straight from the Fortran text,
sharing no code with oracle/*.cpp -- checked against the C++ oracle.  The reference ships no golden vectors for this path and cannot
be built by a Fortran compiler here; the pin to the reference's output is tests/test_ref_transpiled.py (the reference's
own source, translated).  This file is the independent second reading: it rules out transcription slips in the oracle, which would
have to be made identically twice, in two languages, to go unnoticed -- and it does not depend on the translator.

    push    3d/common/particle.f90:75-91 (tmpf staging), :108-184 (shape factors, 27-point gather), :186-222 (Buneman-Boris, move)
            3d/common/particle.f90:369-406 (Vay)
    deposit 3d/common/field.f90:252-396 (S0, shifted S1 through smo_1..3, DS, the three W sums, the transposed flush)
"""
import numpy as np
import pytest

from oracle.pyoracle import World3, weibel_constants
from tests.util import active_mask, make_world3


def _stage_fields(uf):
    """tmpf(1:6,i,j,k), particle.f90:79-87, on the box of uf (k, j, i, comp); the last plane/row/column of the result is unused"""
    t = np.zeros_like(uf)
    f = uf
    t[:-1, :-1, :, 0] = 2.5e-1 * (+f[:-1, :-1, :, 0] + f[:-1, 1:, :, 0] + f[1:, :-1, :, 0] + f[1:, 1:, :, 0])
    t[:-1, :, :-1, 1] = 2.5e-1 * (+f[:-1, :, :-1, 1] + f[:-1, :, 1:, 1] + f[1:, :, :-1, 1] + f[1:, :, 1:, 1])
    t[:, :-1, :-1, 2] = 2.5e-1 * (+f[:, :-1, :-1, 2] + f[:, :-1, 1:, 2] + f[:, 1:, :-1, 2] + f[:, 1:, 1:, 2])
    t[:, :, :-1, 3] = 5e-1 * (+f[:, :, :-1, 3] + f[:, :, 1:, 3])
    t[:, :-1, :, 4] = 5e-1 * (+f[:, :-1, :, 4] + f[:, 1:, :, 4])
    t[:-1, :, :, 5] = 5e-1 * (+f[:-1, :, :, 5] + f[1:, :, :, 5])
    return t


def _shape(dh):
    return 5e-1 * (5e-1 - dh) * (5e-1 - dh), 7.5e-1 - dh * dh, 5e-1 * (5e-1 + dh) * (5e-1 + dh)


def _cells_of(w):
    """(isp, k, j, ii) -> loop cell i of every active particle, from cumcnt (the push uses the loop cell, never int(x))"""
    np2, cc = w.arr("np2"), w.arr("cumcnt")
    m = active_mask(np2, w.np)
    ii = np.broadcast_to(np.arange(w.np), m.shape)
    cell = (ii[..., None] >= cc[..., None, :]).sum(axis=-1) - 1      # number of cumcnt entries <= ii, minus one
    return m, cell + 2                                                 # nxgs = 2


def numpy_push(w, vay=False):
    """gp of every active particle, from up / uf / cumcnt, in the statement order of particle.f90"""
    up, uf = w.arr("up"), w.arr("uf")
    tm = _stage_fields(uf)
    m, ci = _cells_of(w)
    isp, kk, jj, _ = np.nonzero(m)
    p = up[m]                                    # (n, 7)
    i = ci[m]; j = jj + 2; k = kk + 2            # nys = nzs = 2 on a single rank
    c, delt, d_delx = w.c, w.delt, 1.0 / w.delx
    q, r = w.q[isp], w.r[isp]
    fac1 = q / r * 5e-1 * delt
    txxx = fac1 * fac1
    fac2 = q * delt / r
    shx = _shape(p[:, 0] * d_delx - 5e-1 - i)
    shy = _shape(p[:, 1] * d_delx - 5e-1 - j)
    shz = _shape(p[:, 2] * d_delx - 5e-1 - k)
    f = []
    for comp in range(6):
        tot = None
        for dk in (-1, 0, 1):
            plane = None
            for dj in (-1, 0, 1):
                # box index = global index (two ghost layers, origin 2 -> index 2 is the first interior cell)
                row = (+tm[k + dk, j + dj, i - 1, comp] * shx[0] + tm[k + dk, j + dj, i, comp] * shx[1]
                       + tm[k + dk, j + dj, i + 1, comp] * shx[2]) * shy[dj + 1]
                plane = row if plane is None else plane + row
            term = plane * shz[dk + 1]
            tot = term if tot is None else tot + term
        f.append(tot)
    bpx, bpy, bpz, epx, epy, epz = f
    g = np.empty_like(p)
    if not vay:
        uvm1 = p[:, 3] + fac1 * epx; uvm2 = p[:, 4] + fac1 * epy; uvm3 = p[:, 5] + fac1 * epz
        gam = np.sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3)
        igam = 1e0 / gam
        fac1r = fac1 * igam
        fac2r = fac2 / (gam + txxx * (bpx * bpx + bpy * bpy + bpz * bpz) * igam)
        uvm4 = uvm1 + fac1r * (+uvm2 * bpz - uvm3 * bpy)
        uvm5 = uvm2 + fac1r * (+uvm3 * bpx - uvm1 * bpz)
        uvm6 = uvm3 + fac1r * (+uvm1 * bpy - uvm2 * bpx)
        uvm1 = uvm1 + fac2r * (+uvm5 * bpz - uvm6 * bpy)
        uvm2 = uvm2 + fac2r * (+uvm6 * bpx - uvm4 * bpz)
        uvm3 = uvm3 + fac2r * (+uvm4 * bpy - uvm5 * bpx)
        g[:, 3] = uvm1 + fac1 * epx; g[:, 4] = uvm2 + fac1 * epy; g[:, 5] = uvm3 + fac1 * epz
        gam = 1e0 / np.sqrt(1e0 + (+g[:, 3] * g[:, 3] + g[:, 4] * g[:, 4] + g[:, 5] * g[:, 5]) / (c * c))
    else:
        uvm1, uvm2, uvm3 = p[:, 3], p[:, 4], p[:, 5]
        gam = np.sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3)
        fac1r = fac1 / gam
        uvm4 = uvm1 + fac2 * epx + fac1r * (+uvm2 * bpz - uvm3 * bpy)
        uvm5 = uvm2 + fac2 * epy + fac1r * (+uvm3 * bpx - uvm1 * bpz)
        uvm6 = uvm3 + fac2 * epz + fac1r * (+uvm1 * bpy - uvm2 * bpx)
        taux, tauy, tauz = fac1 * bpx / c, fac1 * bpy / c, fac1 * bpz / c
        tau2 = taux * taux + tauy * tauy + tauz * tauz
        ua = (uvm4 * taux + uvm5 * tauy + uvm6 * tauz) / c
        sigma = 1.0 + (uvm4 * uvm4 + uvm5 * uvm5 + uvm6 * uvm6) / (c * c) - tau2
        gam2 = 0.5 * (sigma + np.sqrt(sigma * sigma + 4.0 * (tau2 + ua * ua)))
        gam = np.sqrt(gam2)
        txxx = 1.0 / (tau2 + gam2)
        g[:, 3] = txxx * (gam2 * uvm4 + c * ua * taux + gam * (uvm5 * tauz - uvm6 * tauy))
        g[:, 4] = txxx * (gam2 * uvm5 + c * ua * tauy + gam * (uvm6 * taux - uvm4 * tauz))
        g[:, 5] = txxx * (gam2 * uvm6 + c * ua * tauz + gam * (uvm4 * tauy - uvm5 * taux))
        gam = 1e0 / gam
    g[:, 0] = p[:, 0] + g[:, 3] * delt * gam
    g[:, 1] = p[:, 1] + g[:, 4] * delt * gam
    g[:, 2] = p[:, 2] + g[:, 5] * delt * gam
    g[:, 6] = p[:, 6]
    return m, g


@pytest.mark.parametrize("vay", [False, True], ids=["buneman-boris", "vay"])
def test_push_matches_an_independent_numpy_restatement(vay):
    w = make_world3(10, 6, 5, 4, steps=2)
    rng = np.random.default_rng(11)
    uf = w.arr("uf")
    uf[...] = 5.0 * rng.standard_normal(uf.shape)          # strong fields: every term of the update matters
    (w.particle_solv_vay if vay else w.particle_solv)()
    m, g = numpy_push(w, vay)
    ref = w.arr("gp")[m]
    assert np.array_equal(ref[:, 6].view(np.int64), g[:, 6].view(np.int64))
    scale = np.abs(ref[:, :6]).max(axis=0)
    assert (np.abs(ref[:, :6] - g[:, :6]).max(axis=0) / scale).max() < 5e-15
    w.close()


def numpy_deposit_one(x0, x1, cell, qdxdt):
    """pjx, pjy, pjz(-2:2,-2:2,-2:2) of ONE particle, field.f90:252-383, returned as uj increments d[(ip,jp,kp)] -> (jx,jy,jz)
    after the transposed flush of :390-396"""
    fac = 1e0 / 3e0
    s0, ds = [], []
    for a in range(3):
        dh = x0[a] - 5e-1 - cell[a]
        s = np.zeros(5)
        s[1], s[2], s[3] = _shape(dh)
        i2 = int(x1[a])
        dh = x1[a] - 5e-1 - i2
        inc = i2 - cell[a]
        s1_1, s1_2, s1_3 = _shape(dh)
        smo_1 = -(inc - abs(inc)) * 5e-1 + 0
        smo_2 = -abs(inc) + 1
        smo_3 = (inc + abs(inc)) * 5e-1 + 0
        d = np.array([s1_1 * smo_1, s1_1 * smo_2 + s1_2 * smo_1, s1_2 * smo_2 + s1_3 * smo_1 + s1_1 * smo_3,
                      s1_3 * smo_2 + s1_2 * smo_3, s1_3 * smo_3])
        s0.append(s)
        ds.append(d - s)
    (s0x, s0y, s0z), (dsx, dsy, dsz) = s0, ds
    pjx = np.zeros((5, 5, 5)); pjy = np.zeros((5, 5, 5)); pjz = np.zeros((5, 5, 5))      # index = offset + 2
    for kp in range(5):
        for jp in range(5):
            dstmp = ((s0y[jp] + 5e-1 * dsy[jp]) * s0z[kp] + (5e-1 * s0y[jp] + fac * dsy[jp]) * dsz[kp]) * qdxdt
            pjtmp = 0.0
            for r in range(4):
                pjtmp = pjtmp - dsx[r] * dstmp
                pjx[r + 1, jp, kp] += pjtmp
            dstmp = ((s0x[jp] + 5e-1 * dsx[jp]) * s0z[kp] + (5e-1 * s0x[jp] + fac * dsx[jp]) * dsz[kp]) * qdxdt
            pjtmp = 0.0
            for r in range(4):
                pjtmp = pjtmp - dsy[r] * dstmp
                pjy[r + 1, jp, kp] += pjtmp
            dstmp = ((s0x[jp] + 5e-1 * dsx[jp]) * s0y[kp] + (5e-1 * s0x[jp] + fac * dsx[jp]) * dsy[kp]) * qdxdt
            pjtmp = 0.0
            for r in range(4):
                pjtmp = pjtmp - dsz[r] * dstmp
                pjz[r + 1, jp, kp] += pjtmp
    out = np.zeros((5, 5, 5, 3))                  # [kp, jp, ip, comp]
    for kp in range(5):
        for jp in range(5):
            for ip in range(5):
                out[kp, jp, ip] = (pjx[ip, jp, kp], pjy[jp, ip, kp], pjz[kp, ip, jp])
    return out


def test_deposit_matches_an_independent_numpy_restatement():
    rng = np.random.default_rng(5)
    q, r, _ = weibel_constants(1)
    for trial in range(12):
        w = World3(8, 8, 8, 8 * 3, q=q, r=r)
        up, gp, np2, cc = w.arr("up"), w.arr("gp"), w.arr("np2"), w.arr("cumcnt")
        np2[...] = 0
        cc[...] = 0
        isp = trial % 2
        np2[isp, 3, 3] = 1
        cc[isp, 3, 3, 4:] = 1                        # the particle belongs to x-cell 5, pencil (j, k) = (5, 5)
        x0 = np.array([5.0, 5.0, 5.0]) + rng.random(3)
        x1 = x0 + rng.uniform(-0.95, 0.95, 3)
        up[isp, 3, 3, 0, :3] = x0
        gp[isp, 3, 3, 0, :3] = x1
        w.field_fdtd_i(1)                            # ele_cur only
        uj = w.arr("uj")                             # (k, j, i, 3) on the box, index = global index
        mine = numpy_deposit_one(x0, x1, (5, 5, 5), q[isp] * w.delx / w.delt)
        got = uj[3:8, 3:8, 3:8]
        assert np.abs(got - mine).max() <= 4e-16 * max(np.abs(mine).max(), 1e-300), trial
        outside = uj.copy()
        outside[3:8, 3:8, 3:8] = 0
        assert not outside.any()
        w.close()


def test_field_solve_matches_numpy_fft_restatement():
    """field__fdtd_i on a periodic box, 3d/common/field.f90:126-191 + cgm :409-560, restated with numpy rolls (periodic ghost cells
    are periodic images) and -- instead of conjugate gradients -- an exact FFT inversion of the same constant-coefficient operator
    (f4 - sum of the six neighbour shifts) phi = f5 gkl.  Checks the right-hand side, the operator and its coefficients f1..f5
    (field.f90:57-63), the CG result to its own tolerance (1e-6 of |b|), and the explicit dE update."""
    w = make_world3(12, 8, 6, 4, steps=3)
    c, delt, delx, gfac = w.c, w.delt, w.delx, w.gfac
    f1 = c * delt / delx
    f2 = gfac * f1 * f1
    f3 = 4.0 * np.pi * delx / c
    f5 = (delx / (c * delt * gfac)) ** 2
    f4 = 6.0 + f5
    w.particle_solv()
    w.field_fdtd_i(1)
    w.field_fdtd_i(2)                                  # ele_cur + bc__curre
    inner = (slice(2, -2),) * 3
    uf = w.arr("uf")[inner].copy()                     # (k, j, i, 6) interior, before the update
    uj = w.arr("uj")[inner].copy()
    sh = lambda a, dk, dj, di: np.roll(a, (-dk, -dj, -di), axis=(0, 1, 2))      # a(i+di, j+dj, k+dk)  # noqa: E731
    lap = lambda a: (sh(a, -1, 0, 0) + sh(a, 0, -1, 0) + sh(a, 0, 0, -1) - 6e0 * a + sh(a, 0, 0, 1) + sh(a, 0, 1, 0) + sh(a, 1, 0, 0))  # noqa: E731
    B = [uf[..., n] for n in range(3)]; E = [uf[..., 3 + n] for n in range(3)]; J = [uj[..., n] for n in range(3)]
    gkl = [
        f2 * (lap(B[0]) + f3 * (-sh(J[2], 0, -1, 0) + J[2] + sh(J[1], -1, 0, 0) - J[1])) - f1 * (-sh(E[2], 0, -1, 0) + E[2] + sh(E[1], -1, 0, 0) - E[1]),
        f2 * (lap(B[1]) + f3 * (-sh(J[0], -1, 0, 0) + J[0] + sh(J[2], 0, 0, -1) - J[2])) - f1 * (-sh(E[0], -1, 0, 0) + E[0] + sh(E[2], 0, 0, -1) - E[2]),
        f2 * (lap(B[2]) + f3 * (-sh(J[1], 0, 0, -1) + J[1] + sh(J[0], 0, -1, 0) - J[0])) - f1 * (-sh(E[1], 0, 0, -1) + E[1] + sh(E[0], 0, -1, 0) - E[0]),
    ]
    w.field_fdtd_i(3)
    got = w.arr("gkl")
    for n in range(3):
        assert np.abs(got[..., n] - gkl[n]).max() <= 1e-13 * np.abs(gkl[n]).max(), n
    # exact solution of (f4 - shifts) phi = f5 gkl by FFT
    nz, ny, nx = gkl[0].shape
    kz, ky, kx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    symbol = f4 - 2 * np.cos(2 * np.pi * kx / nx) - 2 * np.cos(2 * np.pi * ky / ny) - 2 * np.cos(2 * np.pi * kz / nz)
    w.field_fdtd_i(4)
    w.field_fdtd_i(5)
    df = w.arr("df")[inner]
    dB = []
    for n in range(3):
        b = f5 * gkl[n]
        exact = np.real(np.fft.ifftn(np.fft.fftn(b) / symbol))
        # the CG stops at |r| <= 1e-6 |b| (field.f90:461-462, 481, 520); the error is bounded by |r| / lambda_min = |r| / f5
        assert np.linalg.norm(df[..., n] - exact) <= 1e-6 * np.linalg.norm(b) / f5 * 1.5, n
        assert np.linalg.norm(df[..., n] - exact) > 0       # it IS an iterative result, not the same computation
        dB.append(df[..., n].copy())
    w.field_fdtd_i(6)
    w.field_fdtd_i(7)
    df = w.arr("df")[inner]
    dE = [
        f1 * (gfac * (-dB[2] + sh(dB[2], 0, 1, 0) + dB[1] - sh(dB[1], 1, 0, 0)) + (-B[2] + sh(B[2], 0, 1, 0) + B[1] - sh(B[1], 1, 0, 0))) - 4e0 * np.pi * delt * J[0],
        f1 * (gfac * (-dB[0] + sh(dB[0], 1, 0, 0) + dB[2] - sh(dB[2], 0, 0, 1)) + (-B[0] + sh(B[0], 1, 0, 0) + B[2] - sh(B[2], 0, 0, 1))) - 4e0 * np.pi * delt * J[1],
        f1 * (gfac * (-dB[1] + sh(dB[1], 0, 0, 1) + dB[0] - sh(dB[0], 0, 1, 0)) + (-B[1] + sh(B[1], 0, 0, 1) + B[0] - sh(B[0], 0, 1, 0))) - 4e0 * np.pi * delt * J[2],
    ]
    for n in range(3):
        assert np.abs(df[..., 3 + n] - dE[n]).max() <= 1e-13 * max(np.abs(dE[n]).max(), 1e-300), n
    w.field_fdtd_i(8)
    new = w.arr("uf")[inner]
    assert np.abs(new[..., :3] - (uf[..., :3] + np.stack(dB, axis=-1))).max() <= 1e-15
    w.close()


def numpy_deposit_one_2d(x0, x1, u1, cell, q, delx, delt, c):
    """pjx, pjy, pjz(-2:2,-2:2) of ONE particle of the 2-D code, 2d/common/field.f90:225-298 (Jz from the pushed momentum)"""
    fac = 1.0 / 3.0
    s0 = np.zeros((5, 2)); ds = np.zeros((5, 2))
    for a in range(2):
        dh = x0[a] / delx - 0.5 - cell[a]
        s0[1, a], s0[2, a], s0[3, a] = _shape(dh)
        i2 = int(x1[a] / delx)
        dh = x1[a] / delx - 0.5 - i2
        inc = i2 - cell[a]
        s1_1, s1_2, s1_3 = _shape(dh)
        smo_1 = -(inc - abs(inc)) * 0.5 + 0
        smo_2 = -abs(inc) + 1
        smo_3 = (inc + abs(inc)) * 0.5 + 0
        ds[:, a] = [s1_1 * smo_1, s1_1 * smo_2 + s1_2 * smo_1, s1_2 * smo_2 + s1_3 * smo_1 + s1_1 * smo_3,
                    s1_3 * smo_2 + s1_2 * smo_3, s1_3 * smo_3]
    ds = ds - s0
    gvz = u1[2] / np.sqrt(1. + (+u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]) / (c * c))
    pjx = np.zeros((5, 5)); pjy = np.zeros((5, 5)); pjz = np.zeros((5, 5))      # [ip+2, jp+2]
    for jp in range(5):
        for ip in range(4):
            pjx[ip + 1, jp] = pjx[ip, jp] - q * delx / delt * ds[ip, 0] * (s0[jp, 1] + 0.5 * ds[jp, 1])
    for jp in range(4):
        for ip in range(5):
            pjy[ip, jp + 1] = pjy[ip, jp] - q * delx / delt * ds[jp, 1] * (s0[ip, 0] + 0.5 * ds[ip, 0])
    for jp in range(5):
        for ip in range(5):
            pjz[ip, jp] = q * gvz * (+s0[ip, 0] * s0[jp, 1] + 0.5 * ds[ip, 0] * s0[jp, 1] + 0.5 * s0[ip, 0] * ds[jp, 1] + fac * ds[ip, 0] * ds[jp, 1])
    return np.stack([pjx.T, pjy.T, pjz.T], axis=-1)      # [jp, ip, comp]


def test_2d_deposit_matches_an_independent_numpy_restatement():
    from oracle.pyoracle import World2
    rng = np.random.default_rng(9)
    q, r, _ = weibel_constants(1)
    for trial in range(12):
        w = World2(8, 8, 8 * 3, q=q, r=r)
        up, gp, np2, cc = w.arr("up"), w.arr("gp"), w.arr("np2"), w.arr("cumcnt")
        np2[...] = 0
        cc[...] = 0
        isp = trial % 2
        np2[isp, 3] = 1
        cc[isp, 3, 4:] = 1                           # x-cell 5 of row j = 5
        x0 = np.array([5.0, 5.0]) + rng.random(2)
        x1 = x0 + rng.uniform(-0.95, 0.95, 2)
        u1 = rng.standard_normal(3)
        up[isp, 3, 0, :2] = x0
        gp[isp, 3, 0, :2] = x1
        gp[isp, 3, 0, 2:5] = u1
        w.field_fdtd_i(1)
        uj = w.arr("uj")                             # (j, i, 3)
        mine = numpy_deposit_one_2d(x0, x1, u1, (5, 5), q[isp], w.delx, w.delt, w.c)
        got = uj[3:8, 3:8]
        assert np.abs(got - mine).max() <= 4e-16 * max(np.abs(mine).max(), 1e-300), trial
        outside = uj.copy()
        outside[3:8, 3:8] = 0
        assert not outside.any()
        w.close()


def test_migration_and_sort_match_a_numpy_key_computation():
    """boundary_periodic__particle_x (3d/common/boundary_periodic.f90:86-92), __particle_yz (:153-171: wrap + destination pencil from
    the PRE-wrap integer cell) and sort__bucket (3d/common/sort.f90:60-86: key int(x)) restated as one numpy key computation on the
    pushed set: every particle's (species, k, j, i) destination, its wrapped coordinates and the resulting np2 / cumcnt must equal
    what the oracle's three procedures produce (records compared as ID-keyed sets: the order inside a cell is not defined)."""
    w = make_world3(10, 6, 5, 5, steps=2)
    nx, ny, nz = w.nx, w.ny, w.nz
    w.particle_solv()
    w.field_fdtd_i()
    np2_0 = w.arr("np2").copy()
    m = active_mask(np2_0, w.np)
    isp = np.nonzero(m)[0]
    g = w.arr("gp")[m].copy()
    w.bc_particle_x(); w.bc_particle_yz(); w.sort_bucket()
    assert w.error() == 0
    # --- numpy restatement ---
    x, y, z = g[:, 0].copy(), g[:, 1].copy(), g[:, 2].copy()
    ipos = np.trunc(x / w.delx).astype(int)
    x = np.where(ipos < 2, x + nx * w.delx, np.where(ipos >= nx + 2, x - nx * w.delx, x))
    jpos = np.trunc(y / w.delx).astype(int); kpos = np.trunc(z / w.delx).astype(int)
    y = np.where(jpos <= 1, y + ny * w.delx, np.where(jpos >= ny + 2, y - ny * w.delx, y))
    z = np.where(kpos <= 1, z + nz * w.delx, np.where(kpos >= nz + 2, z - nz * w.delx, z))
    jd = np.where(jpos <= 1, jpos + ny, np.where(jpos >= ny + 2, jpos - ny, jpos))     # the pencil that receives it (periodic neighbour)
    kd = np.where(kpos <= 1, kpos + nz, np.where(kpos >= nz + 2, kpos - nz, kpos))
    icell = np.trunc(x).astype(int)                                                    # sort.f90:65,77: int(x), no d_delx
    ids = g[:, 6].view(np.int64)
    expect = {}
    for n in range(len(ids)):
        expect[(int(isp[n]), int(ids[n]))] = (int(kd[n]), int(jd[n]), int(icell[n]), x[n], y[n], z[n], g[n, 3], g[n, 4], g[n, 5])
    up, np2, cc = w.arr("up"), w.arr("np2"), w.arr("cumcnt")
    assert int(np2.sum()) == len(ids)
    seen = 0
    for s in range(2):
        for k in range(nz):
            for j in range(ny):
                n = np2[s, k, j]
                rec = up[s, k, j, :n]
                cell = np.searchsorted(cc[s, k, j], np.arange(n), side="right") - 1 + 2
                for t in range(n):
                    e = expect[(s, int(rec[t, 6].view(np.int64)))]
                    assert e[:3] == (k + 2, j + 2, int(cell[t]))
                    assert np.array_equal(rec[t, :6], np.array(e[3:]))      # wrapped coordinates bit-exact
                    seen += 1
                counts = np.bincount(cell - 2, minlength=nx)
                assert np.array_equal(np.diff(cc[s, k, j]), counts[:nx])
    assert seen == len(ids)
    w.close()


@pytest.mark.parametrize("vay", [False, True], ids=["buneman-boris", "vay"])
def test_2d_push_matches_an_independent_numpy_restatement(vay):
    """2d/common/particle.f90:70-82 (staging: Ez is NOT averaged in the 2-D code), :97-133 (9-point gather), :135-171 (update)"""
    from tests.util import make_world2
    w = make_world2(14, 9, 5, steps=2)
    rng = np.random.default_rng(13)
    uf = w.arr("uf")                              # (j, i, 6) on the box
    uf[...] = 5.0 * rng.standard_normal(uf.shape)
    (w.particle_solv_vay if vay else w.particle_solv)()
    f = uf
    tm = np.zeros_like(uf)
    tm[:-1, :, 0] = 0.5 * (+f[:-1, :, 0] + f[1:, :, 0])
    tm[:, :-1, 1] = 0.5 * (+f[:, :-1, 1] + f[:, 1:, 1])
    tm[:-1, :-1, 2] = 0.25 * (+f[:-1, :-1, 2] + f[:-1, 1:, 2] + f[1:, :-1, 2] + f[1:, 1:, 2])
    tm[:, :-1, 3] = 0.5 * (+f[:, :-1, 3] + f[:, 1:, 3])
    tm[:-1, :, 4] = 0.5 * (+f[:-1, :, 4] + f[1:, :, 4])
    tm[:, :, 5] = f[:, :, 5]
    np2, cc, up = w.arr("np2"), w.arr("cumcnt"), w.arr("up")
    m = active_mask(np2, w.np)
    ii = np.broadcast_to(np.arange(w.np), m.shape)
    ci = ((ii[..., None] >= cc[..., None, :]).sum(axis=-1) - 1 + 2)[m]
    isp, jj, _ = np.nonzero(m)
    p = up[m]
    i, j = ci, jj + 2
    c, delt, d_delx = w.c, w.delt, 1.0 / w.delx
    q, r = w.q[isp], w.r[isp]
    fac1 = q / r * 0.5 * delt; txxx = fac1 * fac1; fac2 = q * delt / r
    shx = _shape(p[:, 0] * d_delx - 0.5 - i)
    shy = _shape(p[:, 1] * d_delx - 0.5 - j)
    fl = []
    for comp in range(6):
        tot = None
        for dj in (-1, 0, 1):
            row = (+tm[j + dj, i - 1, comp] * shx[0] + tm[j + dj, i, comp] * shx[1] + tm[j + dj, i + 1, comp] * shx[2]) * shy[dj + 1]
            tot = row if tot is None else tot + row
        fl.append(tot)
    bpx, bpy, bpz, epx, epy, epz = fl
    g = np.empty_like(p)
    if not vay:
        uvm1 = p[:, 2] + fac1 * epx; uvm2 = p[:, 3] + fac1 * epy; uvm3 = p[:, 4] + fac1 * epz
        gam = np.sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3)
        igam = 1. / gam
        fac1r = fac1 * igam
        fac2r = fac2 / (gam + txxx * (bpx * bpx + bpy * bpy + bpz * bpz) * igam)
        uvm4 = uvm1 + fac1r * (+uvm2 * bpz - uvm3 * bpy)
        uvm5 = uvm2 + fac1r * (+uvm3 * bpx - uvm1 * bpz)
        uvm6 = uvm3 + fac1r * (+uvm1 * bpy - uvm2 * bpx)
        uvm1 = uvm1 + fac2r * (+uvm5 * bpz - uvm6 * bpy)
        uvm2 = uvm2 + fac2r * (+uvm6 * bpx - uvm4 * bpz)
        uvm3 = uvm3 + fac2r * (+uvm4 * bpy - uvm5 * bpx)
        g[:, 2] = uvm1 + fac1 * epx; g[:, 3] = uvm2 + fac1 * epy; g[:, 4] = uvm3 + fac1 * epz
        gam = 1. / np.sqrt(1.0 + (+g[:, 2] * g[:, 2] + g[:, 3] * g[:, 3] + g[:, 4] * g[:, 4]) / (c * c))
    else:
        uvm1, uvm2, uvm3 = p[:, 2], p[:, 3], p[:, 4]
        gam = np.sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3)
        fac1r = fac1 / gam
        uvm4 = uvm1 + fac2 * epx + fac1r * (+uvm2 * bpz - uvm3 * bpy)
        uvm5 = uvm2 + fac2 * epy + fac1r * (+uvm3 * bpx - uvm1 * bpz)
        uvm6 = uvm3 + fac2 * epz + fac1r * (+uvm1 * bpy - uvm2 * bpx)
        taux, tauy, tauz = fac1 * bpx / c, fac1 * bpy / c, fac1 * bpz / c
        tau2 = taux * taux + tauy * tauy + tauz * tauz
        ua = (uvm4 * taux + uvm5 * tauy + uvm6 * tauz) / c
        sigma = 1.0 + (uvm4 * uvm4 + uvm5 * uvm5 + uvm6 * uvm6) / (c * c) - tau2
        gam2 = 0.5 * (sigma + np.sqrt(sigma * sigma + 4.0 * (tau2 + ua * ua)))
        gam = np.sqrt(gam2)
        s_ = 1.0 / (tau2 + gam2)
        g[:, 2] = s_ * (gam2 * uvm4 + c * ua * taux + gam * (uvm5 * tauz - uvm6 * tauy))
        g[:, 3] = s_ * (gam2 * uvm5 + c * ua * tauy + gam * (uvm6 * taux - uvm4 * tauz))
        g[:, 4] = s_ * (gam2 * uvm6 + c * ua * tauz + gam * (uvm4 * tauy - uvm5 * taux))
        gam = 1.0 / gam
    g[:, 0] = p[:, 0] + g[:, 2] * delt * gam
    g[:, 1] = p[:, 1] + g[:, 3] * delt * gam
    ref = w.arr("gp")[m]
    scale = np.abs(ref[:, :5]).max(axis=0)
    assert (np.abs(ref[:, :5] - g[:, :5]).max(axis=0) / scale).max() < 5e-15
    assert np.array_equal(ref[:, 5].view(np.int64), p[:, 5].view(np.int64))
    w.close()


@pytest.mark.parametrize("bc", [1, 2], ids=["reconnection", "shock"])
def test_wall_cg_matches_a_dense_numpy_solve(bc):
    """cgm of the 2-D wall set-ups (2d/common/field.f90:319-461 with boundary_reconnection__phi, 2d/proj/reconnection/
    boundary_reconnection.f90:557-577, or boundary_shock__phi, 2d/proj/shock/boundary_shock.f90:603-623) against a dense numpy solve
    of the same operator: (f4 - neighbours) phi = f5 gkl on cells nxs..nxe, periodic in y, with the wall rule of component l folded
    into the matrix (l = 1: phi(nxs-1) = -phi(nxs), phi(nxe+1) = -phi(nxe-2) [reconnection] / 0 [shock]; l = 2, 3:
    phi(nxs-1) = phi(nxs+1), phi(nxe+1) = phi(nxe-1) / 0), then the wall ghost fill of boundary_*__dfield (:350-359 / :396-405)."""
    from tests.util import make_world2
    order = bc
    w = make_world2(14, 8, 6, steps=3, bc=bc, order=order, u0=0.3 if bc == 2 else 0.0)
    nx, ny = w.nx, w.ny
    nxs, nxe = 2, nx + 1
    f5 = (w.delx / (w.c * w.delt * w.gfac)) ** 2
    f4 = 4.0 + f5
    w.particle_solv()
    if bc == 1:
        w.bc_particle_x()
    else:
        w.bc_injection(0.3)
    for st in (1, 2, 3):
        w.field_fdtd_i(st)
    gkl = w.arr("gkl").copy()                      # (j, i, 3) interior
    w.field_fdtd_i(4)
    df = w.arr("df").copy()                        # (j, i, 6) on the box
    nxr = nxe - nxs + 1
    idx = lambda i, j: (j % ny) * nxr + (i - nxs)  # noqa: E731
    for l in (1, 2, 3):
        A = np.zeros((nxr * ny, nxr * ny))
        for j in range(ny):
            for i in range(nxs, nxe + 1):
                row = idx(i, j)
                A[row, idx(i, j)] += f4
                A[row, idx(i, j - 1)] -= 1.0
                A[row, idx(i, j + 1)] -= 1.0
                # x neighbours with the wall rules
                if i > nxs:
                    A[row, idx(i - 1, j)] -= 1.0
                elif l == 1:
                    A[row, idx(nxs, j)] -= -1.0
                else:
                    A[row, idx(nxs + 1, j)] -= 1.0
                if i < nxe:
                    A[row, idx(i + 1, j)] -= 1.0
                elif bc == 1:
                    if l == 1:
                        A[row, idx(nxe - 2, j)] -= -1.0
                    else:
                        A[row, idx(nxe - 1, j)] -= 1.0
                # shock: phi(nxe+1) = 0 -> nothing
        b = (f5 * gkl[:, :, l - 1]).reshape(-1)
        exact = np.linalg.solve(A, b).reshape(ny, nxr)
        got = df[2:-2, 2:-2, l - 1]
        lam_min = np.linalg.eigvalsh(0.5 * (A + A.T)).min()
        assert np.linalg.norm(got - exact) <= 1.5e-6 * np.linalg.norm(b) / lam_min, (bc, l)
    w.field_fdtd_i(5)
    d5 = w.arr("df")
    X = lambda i: i                                 # box x index = global index (two ghosts, nxgs = 2)  # noqa: E731
    assert np.array_equal(d5[:, X(nxs - 1), 0], -d5[:, X(nxs), 0])
    assert np.array_equal(d5[:, X(nxs - 1), 1:4], d5[:, X(nxs + 1), 1:4])
    assert np.array_equal(d5[:, X(nxs - 1), 4:6], -d5[:, X(nxs), 4:6])
    if bc == 1:
        assert np.array_equal(d5[:, X(nxe), 0], -d5[:, X(nxe - 1), 0])
        assert np.array_equal(d5[:, X(nxe + 1), 1:4], d5[:, X(nxe - 1), 1:4])
        assert np.array_equal(d5[:, X(nxe), 4:6], -d5[:, X(nxe - 1), 4:6])
    else:
        assert not d5[:, X(nxe + 1), :].any()
    w.close()


def test_moments_match_a_periodic_numpy_cic_sum():
    """mom_calc__nvt (3d/common/mom_calc.f90:218-332: CIC weights about the cell centres, N, V = u/gamma, T = u^2/gamma) followed by
    boundary_periodic__mom (3d/common/boundary_periodic.f90:1102-1235: ghost layer folded into the periodic image, x then y then z)
    equals, on the interior cells, one CIC sum with periodic index arithmetic.  The momenta are the re-centred ones mom_calc__accl
    leaves in gp (its gather and half-step rotation are the push's, checked above)."""
    w = make_world3(10, 6, 5, 5, steps=2)
    nx, ny, nz = w.nx, w.ny, w.nz
    w.mom_calc()
    np2 = w.arr("np2")
    m = active_mask(np2, w.np)
    isp = np.nonzero(m)[0]
    g = w.arr("gp")[m]
    c = w.c
    ih = np.floor(g[:, 0] / w.delx - 5e-1).astype(int); jh = np.floor(g[:, 1] / w.delx - 5e-1).astype(int)
    kh = np.floor(g[:, 2] / w.delx - 5e-1).astype(int)
    dx = g[:, 0] - 5e-1 - ih; dy = g[:, 1] - 5e-1 - jh; dz = g[:, 2] - 5e-1 - kh
    gam = 1e0 / np.sqrt(1e0 + (g[:, 3] ** 2 + g[:, 4] ** 2 + g[:, 5] ** 2) / (c * c))
    vals = [np.ones_like(gam), g[:, 3] * gam, g[:, 4] * gam, g[:, 5] * gam, g[:, 3] ** 2 * gam, g[:, 4] ** 2 * gam, g[:, 5] ** 2 * gam]
    M = np.zeros((2, nz, ny, nx, 7))
    for ok, wz in ((0, 1 - dz), (1, dz)):
        for oj, wy in ((0, 1 - dy), (1, dy)):
            for oi, wx in ((0, 1 - dx), (1, dx)):
                wgt = wx * wy * wz
                for l in range(7):
                    np.add.at(M[..., l], (isp, (kh + ok - 2) % nz, (jh + oj - 2) % ny, (ih + oi - 2) % nx), vals[l] * wgt)
    ref = w.arr("mom")[:, 1:-1, 1:-1, 1:-1, :]                 # (nsp, nz, ny, nx, 7) interior
    assert np.abs(ref - M).max() <= 1e-12 * np.abs(M).max()
    assert abs(M[..., 0].sum() - np2.sum()) < 1e-9              # every particle's weights sum to one
    w.close()


def test_2d_field_solve_matches_numpy_fft_restatement():
    """2d/common/field.f90:124-174 + cgm :319-461 on a periodic box: right-hand side, exact FFT inversion of (f4 - four shifts) with
    f4 = 4 + f5, explicit dE -- the 2-D counterpart of test_field_solve_matches_numpy_fft_restatement"""
    from tests.util import make_world2
    w = make_world2(16, 10, 5, steps=3)
    c, delt, delx, gfac = w.c, w.delt, w.delx, w.gfac
    f1 = c * delt / delx
    f2 = gfac * f1 * f1
    f3 = 4.0 * np.pi * delx / c
    f5 = (delx / (c * delt * gfac)) ** 2
    f4 = 4.0 + f5
    w.particle_solv()
    w.field_fdtd_i(1)
    w.field_fdtd_i(2)
    inner = (slice(2, -2),) * 2
    uf = w.arr("uf")[inner].copy()                 # (j, i, 6)
    uj = w.arr("uj")[inner].copy()
    sh = lambda a, dj, di: np.roll(a, (-dj, -di), axis=(0, 1))       # a(i+di, j+dj)  # noqa: E731
    lap = lambda a: sh(a, -1, 0) + sh(a, 0, -1) - 4. * a + sh(a, 0, 1) + sh(a, 1, 0)   # noqa: E731
    B = [uf[..., n] for n in range(3)]; E = [uf[..., 3 + n] for n in range(3)]; J = [uj[..., n] for n in range(3)]
    gkl = [
        f2 * (lap(B[0]) + f3 * (-sh(J[2], -1, 0) + J[2])) - f1 * (-sh(E[2], -1, 0) + E[2]),
        f2 * (lap(B[1]) - f3 * (-sh(J[2], 0, -1) + J[2])) + f1 * (-sh(E[2], 0, -1) + E[2]),
        f2 * (lap(B[2]) + f3 * (-sh(J[1], 0, -1) + J[1] + sh(J[0], -1, 0) - J[0])) - f1 * (-sh(E[1], 0, -1) + E[1] + sh(E[0], -1, 0) - E[0]),
    ]
    w.field_fdtd_i(3)
    got = w.arr("gkl")
    for n in range(3):
        assert np.abs(got[..., n] - gkl[n]).max() <= 1e-13 * np.abs(gkl[n]).max(), n
    ny, nx = gkl[0].shape
    ky, kx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    symbol = f4 - 2 * np.cos(2 * np.pi * kx / nx) - 2 * np.cos(2 * np.pi * ky / ny)
    w.field_fdtd_i(4)
    w.field_fdtd_i(5)
    df = w.arr("df")[inner]
    dB = []
    for n in range(3):
        b = f5 * gkl[n]
        exact = np.real(np.fft.ifft2(np.fft.fft2(b) / symbol))
        assert np.linalg.norm(df[..., n] - exact) <= 1.5e-6 * np.linalg.norm(b) / f5, n
        dB.append(df[..., n].copy())
    w.field_fdtd_i(6)
    w.field_fdtd_i(7)
    df = w.arr("df")[inner]
    dE = [
        +f1 * (gfac * (-dB[2] + sh(dB[2], 1, 0)) + (-B[2] + sh(B[2], 1, 0))) - 4. * np.pi * delt * J[0],
        -f1 * (gfac * (-dB[2] + sh(dB[2], 0, 1)) + (-B[2] + sh(B[2], 0, 1))) - 4. * np.pi * delt * J[1],
        +f1 * (gfac * (-dB[1] + sh(dB[1], 0, 1) + dB[0] - sh(dB[0], 1, 0)) + (-B[1] + sh(B[1], 0, 1) + B[0] - sh(B[0], 1, 0))) - 4. * np.pi * delt * J[2],
    ]
    for n in range(3):
        assert np.abs(df[..., 3 + n] - dE[n]).max() <= 1e-13 * max(np.abs(dE[n]).max(), 1e-300), n
    w.close()


@pytest.mark.parametrize("bc", [1, 2], ids=["reconnection", "shock"])
def test_3d_wall_cg_matches_a_dense_numpy_solve(bc):
    """the 3-D counterpart of test_wall_cg_matches_a_dense_numpy_solve: cgm (3d/common/field.f90:409-560) with
    boundary_reconnection__phi (3d/proj/reconnection/boundary_reconnection.f90:1094-1116) or boundary_shock__phi
    (3d/proj/shock/boundary_shock.f90:1086-1108) as a dense solve, and the x ghost fills of boundary_*__dfield (:672-682 / :674-686)"""
    w = make_world3(9, 5, 4, 5, steps=2, bc=bc, order=bc, u0=0.3 if bc == 2 else 0.0)
    nx, ny, nz = w.nx, w.ny, w.nz
    nxs, nxe = 2, nx + 1
    f5 = (w.delx / (w.c * w.delt * w.gfac)) ** 2
    f4 = 6.0 + f5
    w.particle_solv()
    if bc == 1:
        w.bc_particle_x()
    else:
        w.bc_injection(0.3)
    for st in (1, 2, 3):
        w.field_fdtd_i(st)
    gkl = w.arr("gkl").copy()                      # (k, j, i, 3)
    w.field_fdtd_i(4)
    df = w.arr("df").copy()
    nxr = nxe - nxs + 1
    idx = lambda i, j, k: ((k % nz) * ny + (j % ny)) * nxr + (i - nxs)   # noqa: E731
    n = nxr * ny * nz
    for l in (1, 2, 3):
        A = np.zeros((n, n))
        for k in range(nz):
            for j in range(ny):
                for i in range(nxs, nxe + 1):
                    row = idx(i, j, k)
                    A[row, row] += f4
                    for dj, dk in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                        A[row, idx(i, j + dj, k + dk)] -= 1.0
                    if i > nxs:
                        A[row, idx(i - 1, j, k)] -= 1.0
                    elif l == 1:
                        A[row, idx(nxs, j, k)] += 1.0          # phi(nxs-1) = -phi(nxs)
                    else:
                        A[row, idx(nxs + 1, j, k)] -= 1.0      # phi(nxs-1) = phi(nxs+1)
                    if i < nxe:
                        A[row, idx(i + 1, j, k)] -= 1.0
                    elif bc == 1:
                        if l == 1:
                            A[row, idx(nxe - 2, j, k)] += 1.0  # phi(nxe+1) = -phi(nxe-2)
                        else:
                            A[row, idx(nxe - 1, j, k)] -= 1.0  # phi(nxe+1) = phi(nxe-1)
        b = (f5 * gkl[..., l - 1]).reshape(-1)
        exact = np.linalg.solve(A, b).reshape(nz, ny, nxr)
        got = df[2:-2, 2:-2, 2:-2, l - 1]
        lam_min = np.linalg.eigvalsh(0.5 * (A + A.T)).min()
        assert np.linalg.norm(got - exact) <= 1.5e-6 * np.linalg.norm(b) / lam_min, (bc, l)
    w.field_fdtd_i(5)
    d5 = w.arr("df")                               # (k, j, i, 6), box x index = global index
    assert np.array_equal(d5[:, :, nxs - 1, 0], -d5[:, :, nxs, 0])
    assert np.array_equal(d5[:, :, nxs - 1, 1:4], d5[:, :, nxs + 1, 1:4])
    assert np.array_equal(d5[:, :, nxs - 1, 4:6], -d5[:, :, nxs, 4:6])
    if bc == 1:
        assert np.array_equal(d5[:, :, nxe, 0], -d5[:, :, nxe - 1, 0])
        assert np.array_equal(d5[:, :, nxe + 1, 1:4], d5[:, :, nxe - 1, 1:4])
        assert np.array_equal(d5[:, :, nxe, 4:6], -d5[:, :, nxe - 1, 4:6])
    else:
        assert not d5[:, :, nxe + 1, :].any()
    w.close()


def test_deposit_plus_current_fold_is_a_periodic_sum():
    """ele_cur over all particles (3d/common/field.f90:238-404) followed by boundary_periodic__curre (3d/common/boundary_periodic.f90:
    676-978: the two ghost layers folded onto their periodic images, x then y then z) equals, on the interior, the sum of the
    single-particle stencils of numpy_deposit_one placed with periodic index arithmetic."""
    w = make_world3(8, 4, 4, 3, steps=2)
    nx, ny, nz = w.nx, w.ny, w.nz
    w.particle_solv()
    w.field_fdtd_i(1)
    w.field_fdtd_i(2)
    np2, cc, up, gp = w.arr("np2"), w.arr("cumcnt"), w.arr("up"), w.arr("gp")
    J = np.zeros((nz, ny, nx, 3))
    m, ci = _cells_of(w)
    isp, kk, jj, ii = np.nonzero(m)
    for n in range(len(isp)):
        s, k, j, t = isp[n], kk[n], jj[n], ii[n]
        cell = (int(ci[s, k, j, t]), j + 2, k + 2)
        st = numpy_deposit_one(up[s, k, j, t, :3], gp[s, k, j, t, :3], cell, w.q[s] * w.delx / w.delt)     # [kp, jp, ip, comp]
        for kp in range(5):
            for jp in range(5):
                for ip in range(5):
                    J[(cell[2] + kp - 2 - 2) % nz, (cell[1] + jp - 2 - 2) % ny, (cell[0] + ip - 2 - 2) % nx] += st[kp, jp, ip]
    got = w.arr("uj")[2:-2, 2:-2, 2:-2]
    assert np.abs(got - J).max() <= 1e-12 * np.abs(J).max()
    w.close()


def test_wall_particle_rules_match_numpy():
    """boundary_reconnection__particle_x (3d/proj/reconnection/boundary_reconnection.f90:69-110: reflect about (nxs+1) delx and
    (nxe-1) delx, all three momenta negated on the left, ux,uy,uz on the right) and boundary_shock__injection
    (3d/proj/shock/boundary_shock.f90:424-469: left wall as above, moving right wall xend = nxe delx + v0 delt with ux -> 2 u0 - ux)"""
    for bc, u0 in ((1, 0.0), (2, 0.3)):
        w = make_world3(10, 4, 4, 6, steps=1, bc=bc, order=bc, u0=u0)
        nxs, nxe = 2, w.nx + 1
        w.particle_solv()
        m = active_mask(w.arr("np2"), w.np)
        # push some particles through both walls
        gp = w.arr("gp")
        rng = np.random.default_rng(2)
        sel = rng.random(gp.shape[:-1]) < 0.3
        beyond = (nxe - 1 + rng.random(gp.shape[:-1])) if bc == 1 else (nxe + 0.3 + 0.5 * rng.random(gp.shape[:-1]))
        gp[..., 0] = np.where(m & sel, np.where(rng.random(gp.shape[:-1]) < 0.5, nxs + 1 - rng.random(gp.shape[:-1]), beyond), gp[..., 0])
        before = gp[m].copy()
        if bc == 1:
            w.bc_particle_x()
        else:
            w.bc_injection(u0)
        after = w.arr("gp")[m]
        x, u = before[:, 0].copy(), before[:, 3:6].copy()
        ipos = np.trunc(x / w.delx).astype(int)
        left = ipos < nxs + 1
        if bc == 1:
            right = ~left & (ipos >= nxe - 1)
            xr = 2.0 * (nxe - 1) * w.delx - x
            ur = -u
        else:
            v0 = u0 / np.sqrt(1.0 + u0 * u0 / (w.c * w.c))
            xend = nxe * w.delx + v0 * w.delt
            right = ~left & (x > xend)
            xr = 2.0 * xend - x
            ur = np.stack([2.0 * u0 - u[:, 0], -u[:, 1], -u[:, 2]], axis=1)
        xn = np.where(left, 2.0 * (nxs + 1) * w.delx - x, np.where(right, xr, x))
        un = np.where(left[:, None], -u, np.where(right[:, None], ur, u))
        assert left.any() and right.any()
        assert np.array_equal(after[:, 0], xn) and np.array_equal(after[:, 3:6], un)
        assert np.array_equal(after[:, 1:3], before[:, 1:3])
        w.close()
