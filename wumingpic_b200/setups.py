"""Host-side set-up harness: the initial loads of the reference's reconnection and shock drivers (and the shock driver's per-step
injection bookkeeping) restated in numpy, so that BASELINE.json configs[2..4] can be started on this backend -- and on the oracle from
the very same arrays -- without the Fortran drivers (no Fortran compiler / MPI in this image).

  reconnection  2d/proj/reconnection/app.f90:286-307 (constants), :362-452 (Harris sheet + localised perturbation);
                3d/proj/reconnection/app.f90:298-322, :395-470
  shock         2d/proj/shock/app.f90:318-359 (constants, nominal cumcnt), :410-476 (uniform upstream flow), :697-785 (how many
                particles inject() adds per row), :883-893 (vprofile); 3d/proj/shock/app.f90:326-368, :425-500, :733-820

Everything here is rank-independent: the random draws of a pencil depend on (seed, species-independent pencil index) only -- numpy's
Philox bit generator keyed per pencil -- so every rank can build just its own slab and any slab count gives the same global state.  The
reference seeds its RNG from the clock (utils/wuming_utils.f90:48-53), so there is no reference stream to reproduce; the DISTRIBUTIONS
are the reference's, formula by formula, including its single-precision literals (`sqrt(2.)`, 2d/proj/reconnection/app.f90:410-411).

Arrays come back in the reference layout (C order = reversed Fortran shape), cell-sorted with their cumcnt where the driver sorts
after loading (reconnection: app.f90:336-340), or with the driver's nominal cumcnt where it does not (shock: app.f90:346-359 -- the
device treats the host-provided cumcnt as authoritative exactly like the reference's first push does).
"""
import math
from dataclasses import dataclass, field

import numpy as np

PI = 4.0 * math.atan(1.0)
WM_BC_PERIODIC, WM_BC_RECONNECTION, WM_BC_SHOCK = 0, 1, 2


@dataclass
class Setup:
    name: str
    dim: int
    nx: int
    ny: int
    nz: int
    np_cap: int                 # pencil capacity `np` of the host arrays
    nxs: int
    nxe: int
    bc: int                     # WM_BC_*
    order: int                  # WM_ORDER_* (same numbering)
    delt: float
    c: float
    gfac: float
    q: np.ndarray
    r: np.ndarray
    n0: int
    u0: float = 0.0
    extra: dict = field(default_factory=dict)

    @property
    def ndim(self):
        return 6 if self.dim == 2 else 7


def _rng(seed, pencil):
    return np.random.Generator(np.random.Philox(key=[seed & (2**64 - 1), pencil]))


def _normal(rng, n):
    """utils/wuming_utils.f90:72-90 (Box-Muller on two uniforms; the sin / cos pair in the order the reference hands them out)"""
    m = (n + 1) // 2
    x1, x2 = rng.random(m), rng.random(m)
    rr = np.sqrt(-2.0 * np.log(1.0 - x1) + 1.0e-30)
    out = np.empty(2 * m)
    out[0::2] = rr * np.sin(2.0 * PI * x2)
    out[1::2] = rr * np.cos(2.0 * PI * x2)
    return out[:n]


def _sort_rows(up, np2, nx, nxgs=2):
    """sort__bucket on freshly loaded pencils (2d/common/sort.f90:52-80): stable counting sort by int(x); returns cumcnt"""
    lead = np2.shape
    flat_up = up.reshape((-1,) + up.shape[-2:])
    cc = np.zeros((flat_up.shape[0], nx + 1), dtype=np.int32)
    for p, n in enumerate(np2.reshape(-1)):
        rec = flat_up[p, :n]
        cell = rec[:, 0].astype(np.int64) - nxgs
        o = np.argsort(cell, kind="stable")
        flat_up[p, :n] = rec[o]
        cc[p, 1:] = np.cumsum(np.bincount(np.clip(cell, 0, nx - 1), minlength=nx))
    return cc.reshape(lead + (nx + 1,))


# ---------------------------------------------------------------------------------------------------------------------------
# magnetic reconnection: Harris current sheet along y, localised X-point perturbation, conducting walls in x
# ---------------------------------------------------------------------------------------------------------------------------
def reconnection_constants(nx, ny, nz=None, mass_ratio=16.0, alpha=2.0, rtemp=0.2, lcs=0.5, nbg=50, ncs=250):
    """2d/proj/reconnection/app.f90:243-258, 286-307 [3d :253-322]"""
    dim = 2 if nz is None else 3
    c, delx, cfl, gfac = 1.0, 1.0, 0.5, 0.501
    r = np.array([mass_ratio, 1.0])
    delt = cfl * delx / c
    ldb = delx
    vte = math.sqrt(rtemp) * c / (math.sqrt(1 + rtemp) * alpha)
    vti = vte * math.sqrt(r[1] / r[0]) / math.sqrt(rtemp)
    wpe = vte / ldb / math.sqrt(2.0)
    wpi = wpe * math.sqrt(r[1] / r[0])
    wge = wpe / alpha
    wgi = wge / mass_ratio
    n0 = nbg + ncs
    q = np.array([+math.sqrt(r[0] / (4 * PI * n0 / delx**2)) * wpi, -math.sqrt(r[1] / (4 * PI * n0 / delx**2)) * wpe])
    b0 = r[0] * c / q[0] * wgi
    nxgs, nxge, nygs, nyge = 2, nx + 1, 2, ny + 1
    x0 = 0.5 * (nxge + nxgs) * delx
    y0 = 0.5 * (nyge - nygs) * delx            # sic (app.f90:301)
    lcs_len = lcs * c / wpi
    npr = int(nbg * (nxge - nxgs) + ncs * 2 * lcs_len)      # integer = real assignment truncates (app.f90:306)
    s = Setup("reconnection", dim, nx, ny, nz or 1, n0 * nx, nxgs, nxge, WM_BC_RECONNECTION, 1, delt, c, gfac, q, r, n0)
    s.extra = dict(vte=vte, vti=vti, b0=b0, x0=x0, y0=y0, lcs=lcs_len, np_row=npr, nbg=nbg, ncs=ncs, rtemp=rtemp,
                   ibg=nbg * (nxge - nxgs - 2))
    return s


def reconnection_slab(s, nys, nye, nzs=2, nze=2, seed=20240601):
    """set_initial_condition of the reconnection drivers for the pencils nys..nye (, nzs..nze), then sort__bucket:
    returns up, np2, cumcnt, uf of that slab"""
    e = s.extra
    b0, x0, y0, lcs, npr, ibg = e["b0"], e["x0"], e["y0"], e["lcs"], e["np_row"], e["ibg"]
    e1 = 0.12
    nxgs, nxge = 2, s.nx + 1
    nyl, nzl = nye - nys + 1, (nze - nzs + 1 if s.dim == 3 else 1)
    nd = s.ndim
    # --- field, ghost cells included, from the global coordinates (app.f90:386-399)
    ii = np.arange(nxgs - 2, nxge + 3, dtype=np.float64)
    jj = np.arange(nys - 2, nye + 3, dtype=np.float64)
    X, Y = np.meshgrid(ii, jj)                                       # (ny+4, nx+4)
    g = np.exp(-((X - x0)**2 + (Y - y0)**2) / (2 * lcs)**2)
    uf2 = np.zeros((nyl + 4, s.nx + 4, 6))
    uf2[..., 0] = +e1 * b0 * ((Y - y0) / lcs) * g
    uf2[..., 1] = b0 * np.tanh((X - x0) / lcs) + (-e1 * b0 * ((X - x0) / lcs) * g)
    uf = uf2 if s.dim == 2 else np.ascontiguousarray(np.broadcast_to(uf2, (nzl + 4,) + uf2.shape))

    def density(x):
        return e["ncs"] * np.cosh((x - x0) / lcs)**(-2) + e["nbg"]

    def jz(x, y):
        gg = np.exp(-((x - x0)**2 + (y - y0)**2) / (2 * lcs)**2)
        return (+b0 / (4 * PI * lcs) * np.cosh((x - x0) / lcs)**(-2)
                - 2 * e1 * b0 / (4 * PI * lcs) * (1.0 - ((x - x0)**2 + (y - y0)**2) / (2 * lcs)**2) * gg)

    f1 = 1.0 / ((1.0 + e["rtemp"]) * s.q[0])
    f2 = e["rtemp"] / ((1.0 + e["rtemp"]) * s.q[1])
    rt2 = float(np.sqrt(np.float32(2.0)))                           # `sqrt(2.)`: a single-precision sqrt promoted to double
    sdi, sde = e["vti"] / rt2, e["vte"] / rt2
    up = np.zeros((2, nzl, nyl, s.np_cap, nd))
    np2 = np.full((2, nzl, nyl), npr, dtype=np.int32)
    U = 2 if s.dim == 2 else 3                                       # first momentum column
    for kz in range(nzl):
        k = nzs + kz
        for jy in range(nyl):
            j = nys + jy
            pencil = (k - 2) * s.ny + (j - 2) if s.dim == 3 else (j - 2)
            rng = _rng(seed, pencil)
            x = np.empty(npr)
            x[:ibg] = (nxgs + 1) + rng.random(ibg) * (nxge - nxgs - 2)
            r1 = rng.random(npr - ibg)
            r1 = (2.0 * r1 - 1.0) * math.tanh(0.5 * (nxge - nxgs - 2) / lcs)
            x[ibg:] = lcs * 0.5 * (np.log(1.0 + r1) - np.log(1.0 - r1)) + x0
            y = j + rng.random(npr)
            z = k + rng.random(npr) if s.dim == 3 else None
            for isp, (sd, f) in enumerate(((sdi, f1), (sde, f2))):
                rec = up[isp, kz, jy, :npr]
                rec[:, 0], rec[:, 1] = x, y
                if s.dim == 3:
                    rec[:, 2] = z
                nrm = _normal(rng, 3 * npr).reshape(npr, 3)
                rec[:, U] = sd * nrm[:, 0]
                rec[:, U + 1] = sd * nrm[:, 1]
                rec[:, U + 2] = sd * nrm[:, 2] + f * jz(x, y) / density(x)
                pid = -(np.int64(pencil) * npr + np.arange(1, npr + 1, dtype=np.int64))
                rec[:, nd - 1] = pid.view(np.float64)
    cumcnt = _sort_rows(up, np2, s.nx)
    if s.dim == 2:
        return up[:, 0], np2[:, 0], cumcnt[:, 0], uf
    return up, np2, cumcnt, uf


# ---------------------------------------------------------------------------------------------------------------------------
# collisionless shock: cold upstream flow towards a reflecting wall at the left, injection wall on the right, growing box
# ---------------------------------------------------------------------------------------------------------------------------
def shock_constants(nx, n_x_ini, ny, nz=None, n_ppc=2, u_inject=40.0, mass_ratio=1.0, sigma_e=0.1, omega_pe=0.1, v_the=0.0, v_thi=0.0,
                    theta_bn=90.0, phi_bn=90.0, l_damp_ini=100.0):
    """2d/proj/shock/app.f90:263-281, 318-331 [3d :326-341]"""
    dim = 2 if nz is None else 3
    c, delx, cfl, gfac = 1.0, 1.0, 1.0, 0.501
    delt = cfl * delx / c
    u0 = -abs(u_inject)
    gam0 = math.sqrt(1 + (u0 * u0) / (c * c))
    v0 = u0 / gam0
    wpe = omega_pe
    wge = omega_pe * math.sqrt(sigma_e)
    wpi = wpe / math.sqrt(mass_ratio)
    wgi = wge / mass_ratio
    n0 = n_ppc
    r = np.array([mass_ratio, 1.0])
    q = np.array([+math.sqrt(gam0 * r[0] / (4 * PI * n0 / delx**2)) * wpi, -math.sqrt(gam0 * r[1] / (4 * PI * n0 / delx**2)) * wpe])
    b0 = r[0] * c / q[0] * wgi * gam0
    nxgs = 2
    s = Setup("shock", dim, nx, ny, nz or 1, n_ppc * nx * 5, nxgs, nxgs + n_x_ini, WM_BC_SHOCK, 2, delt, c, gfac, q, r, n0, u0=u0)
    s.extra = dict(v0=v0, gam0=gam0, b0=b0, v_thi=v_thi, v_the=v_the, theta_bn=theta_bn * PI / 180, phi_bn=phi_bn * PI / 180,
                   l_damp_ini=l_damp_ini)
    return s


def vprofile(s, x):
    """2d/proj/shock/app.f90:883-893"""
    x0 = s.extra["l_damp_ini"] + 2 * 1.0
    xs = s.extra["l_damp_ini"] * 0.1
    return 0.5 * s.extra["v0"] * (1 + np.tanh((x - x0) / xs))


def shock_slab(s, nys, nye, nzs=2, nze=2, seed=20240601):
    """set_initial_condition of the shock drivers (2d/proj/shock/app.f90:410-476) with the driver's NOMINAL cumcnt (:346-359: n0 per cell
    from cell nxs+1 on, entries above nxe left at zero) -- the particles are NOT sorted by the driver before the first step"""
    e = s.extra
    b0, th, ph, c = e["b0"], e["theta_bn"], e["phi_bn"], s.c
    nxgs, nxge, nxs, nxe = 2, s.nx + 1, s.nxs, s.nxe
    nyl, nzl = nye - nys + 1, (nze - nzs + 1 if s.dim == 3 else 1)
    nd, n0 = s.ndim, s.n0
    ii = np.arange(nxgs - 2, nxge + 3, dtype=np.float64)
    row = np.zeros((s.nx + 4, 6))
    row[:, 0] = b0 * math.cos(th)
    row[:, 1] = b0 * math.sin(th) * math.cos(ph)
    row[:, 2] = b0 * math.sin(th) * math.sin(ph)
    row[:, 4] = +vprofile(s, ii) * row[:, 2] / c
    row[:, 5] = -vprofile(s, ii) * row[:, 1] / c
    shape = (nyl + 4, s.nx + 4, 6) if s.dim == 2 else (nzl + 4, nyl + 4, s.nx + 4, 6)
    uf = np.ascontiguousarray(np.broadcast_to(row, shape))
    npr = n0 * (nxe - nxs - 1)
    if npr > s.np_cap:
        raise ValueError("Error: Too large number of particles")
    up = np.zeros((2, nzl, nyl, s.np_cap, nd))
    np2 = np.full((2, nzl, nyl), npr, dtype=np.int32)
    cumcnt = np.zeros((2, nzl, nyl, s.nx + 1), dtype=np.int32)
    cc = np.zeros(s.nx + 1, dtype=np.int32)
    for i in range(nxs + 2, nxe + 1):
        cc[i - nxgs] = cc[i - 1 - nxgs] + n0
    cumcnt[...] = cc
    U = 2 if s.dim == 2 else 3
    x = (nxs + (nxe - nxs) * (np.arange(1, npr + 1) - 0.5) / npr) * 1.0
    v1 = vprofile(s, x)
    gam1 = 1 / np.sqrt(1 - (v1 / c)**2)
    for kz in range(nzl):
        k = nzs + kz
        for jy in range(nyl):
            j = nys + jy
            pencil = (k - 2) * s.ny + (j - 2) if s.dim == 3 else (j - 2)
            rng = _rng(seed, pencil)
            y = j + rng.random(npr)
            z = k + rng.random(npr) if s.dim == 3 else None
            for isp, sd in enumerate((e["v_thi"], e["v_the"])):
                rec = up[isp, kz, jy, :npr]
                rec[:, 0], rec[:, 1] = x, y
                if s.dim == 3:
                    rec[:, 2] = z
                nrm = sd * _normal(rng, 3 * npr).reshape(npr, 3)
                gamp = np.sqrt(1 + (nrm**2).sum(axis=1) / c**2)
                rec[:, U] = gam1 * (nrm[:, 0] + v1 * gamp)          # Lorentz transform to the lab frame (app.f90:462-470)
                rec[:, U + 1], rec[:, U + 2] = nrm[:, 1], nrm[:, 2]
                pid = -(np.int64(pencil) * npr + np.arange(1, npr + 1, dtype=np.int64))
                rec[:, nd - 1] = pid.view(np.float64)
    if s.dim == 2:
        return up[:, 0], np2[:, 0], cumcnt[:, 0], uf
    return up, np2, cumcnt, uf


def shock_inject_counts(s, it, nproc=1, seed=20240601):
    """How many particles inject() adds behind every GLOBAL row this step (2d/proj/shock/app.f90:711-743): the flux n0 |v0| dt per row
    times the rows, its fractional part by a random draw, split evenly over the ranks and, inside a rank, over its rows, remainders to
    randomly chosen ranks / rows.  Returns int32[nrows_global] in global row order (row = (k - nzgs) ny + (j - nygs))."""
    nrows = s.ny * s.nz
    rng = _rng(seed ^ 0x5bd1e995, 1_000_000 + it)
    pflux = s.n0 * abs(s.extra["v0"]) * s.delt * 1.0 * nrows
    nginj = int(pflux)
    if rng.random() < pflux - int(pflux):
        nginj += 1
    per_proc = np.full(nproc, nginj // nproc, dtype=np.int64)
    per_proc[rng.permutation(nproc)[:nginj % nproc]] += 1
    out = np.zeros(nrows, dtype=np.int32)
    rows_per = nrows // nproc
    for p in range(nproc):
        lo = p * rows_per
        n = rows_per if p < nproc - 1 else nrows - lo
        out[lo:lo + n] = per_proc[p] // n
        out[lo + rng.permutation(n)[:per_proc[p] % n]] += 1
    return out


def shock_params(s, seed=20240601):
    """wm_shock_params (include/wuming_b200.h) of this set-up, for wm_shock_inject / wm_shock_relocate"""
    from .backend import ShockParams
    e = s.extra
    return ShockParams(n0=s.n0, v0=e["v0"], v_thi=e["v_thi"], v_the=e["v_the"], b0=e["b0"], theta_bn=e["theta_bn"], phi_bn=e["phi_bn"],
                       l_damp_ini=e["l_damp_ini"], seed=seed)
