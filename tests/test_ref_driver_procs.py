"""The DRIVER-level procedures of the reference (`proj/*/app.f90`) through the translator, against the oracle's and
wumingpic_b200/setups.py's restatements of them.

oracle/f2cxx/app_harness.py extracts `set_initial_condition`, `set_particle_ids`, `inject`, `relocate`, `get_global_cumsum`,
`vprofile` (shock), the Harris-sheet loader (reconnection), the Weibel loaders and `energy_history` verbatim from the reference files
where they lie -- plus the statement blocks of `load_config` / `init` that derive sizes and physical constants -- and f2cxx.py
translates them.  Fortran's `random_number` stream is not reproducible (the reference seeds it from the clock), so the random
inputs (`uniform_rand`, `normal_rand`, `shuffle` of utils/wuming_utils.f90) are handed to the translated code IN CALL ORDER: the very
values the restatement under test gives the same particle (the oracle's keyed Philox streams; setups.py's per-pencil generators).
What is compared is therefore every deterministic statement of these procedures: placement, Lorentz boost, velocity profile, ID
numbering, np2 / cumcnt bookkeeping, the upstream field columns, the injection counts, the Harris-sheet fields and drifts."""
import numpy as np
import pytest

from oracle import pyoracle
from oracle.f2cxx import app_harness, pyref
from oracle.pyoracle import ShockPrm, World2, World3
from tests.util import active_mask
from wumingpic_b200 import setups

SEED = 20240601


def have(name):
    return app_harness.build(name) is not None and pyref.available(2) and pyref.available(3)


@pytest.fixture(autouse=True)
def one_thread():
    before = pyoracle.num_threads()
    pyoracle.set_num_threads(1)
    yield
    pyoracle.set_num_threads(before)


# ---------------------------------------------------------------------------------------------------------------------------
# shock
# ---------------------------------------------------------------------------------------------------------------------------
SHOCK_CFG = dict(num_process=1, n_ppc=4, n_x=24, n_x_ini=12, n_y=6, u_inject=0.3, mass_ratio=4.0, sigma_e=0.1, omega_pe=0.1,
                 v_the=0.03, v_thi=0.02, theta_bn=90.0, phi_bn=60.0, l_damp_ini=4.0)


def shock_app(dim, nz=4):
    cfg = dict(SHOCK_CFG)
    if dim == 3:
        cfg.update(num_process_j=1, n_z=nz)
    A = pyref.RefApp(f"shock{dim}d")
    ny = cfg["n_y"]
    A.configure([0, 2, ny + 1] if dim == 2 else [0, 2, ny + 1, 2, nz + 1, 0, 0], **cfg)
    return A, cfg


def rows_of(dim, ny, nz):
    """(j, k, global row) in the reference's loop order (k outer, j inner)"""
    if dim == 2:
        return [(j, None, j - 2) for j in range(2, ny + 2)]
    return [(j, k, (k - 2) * ny + (j - 2)) for k in range(2, nz + 2) for j in range(2, ny + 2)]


@pytest.mark.parametrize("dim", [2, 3])
def test_shock_loader_equals_setups_py(dim):
    """set_initial_condition + set_particle_ids + the constants and the nominal cumcnt of init (2d/proj/shock/app.f90:318-359,
    410-476; 3d :326-368, 425-500) against wumingpic_b200.setups.shock_constants / shock_slab"""
    if not have(f"shock{dim}d"):
        pytest.skip("the translated driver procedures cannot be built here")
    nz = 3
    A, cfg = shock_app(dim, nz)
    ny, n0 = cfg["n_y"], cfg["n_ppc"]
    s = setups.shock_constants(cfg["n_x"], cfg["n_x_ini"], ny, nz if dim == 3 else None, n_ppc=n0, u_inject=cfg["u_inject"],
                               mass_ratio=cfg["mass_ratio"], sigma_e=cfg["sigma_e"], omega_pe=cfg["omega_pe"], v_the=cfg["v_the"],
                               v_thi=cfg["v_thi"], theta_bn=cfg["theta_bn"], phi_bn=cfg["phi_bn"], l_damp_ini=cfg["l_damp_ini"])
    npr = n0 * (cfg["n_x_ini"] - 1)
    # replay setups.py's draws per pencil and arrange them in the reference's call order: positions (all rows), then per species
    # (all rows) three normals per particle
    uni, nrm = [], {0: [], 1: []}
    for j, k, row in rows_of(dim, ny, nz):
        rng = setups._rng(SEED, row)
        y = rng.random(npr)
        z = rng.random(npr) if dim == 3 else None
        uni.append(np.stack([y, z], axis=1).ravel() if dim == 3 else y)      # 3-D: y and z of a particle are drawn back to back
        for isp in range(2):
            nrm[isp].append(setups._normal(rng, 3 * npr))
    A.feed(uniform=np.concatenate(uni), normal=np.concatenate(nrm[0] + nrm[1]))
    A.call("harness__init")
    assert A.leftover() == (0, 0, 0)
    # constants
    assert A.scalar("np").value == s.np_cap and A.scalar("nxe").value == s.nxe and A.scalar("n0").value == s.n0
    for name, want in (("delt", s.delt), ("u0", s.u0), ("v0", s.extra["v0"]), ("gam0", s.extra["gam0"]), ("b0", s.extra["b0"]),
                       ("theta_bn", s.extra["theta_bn"]), ("phi_bn", s.extra["phi_bn"])):
        assert A.scalar(name).value == pytest.approx(want, rel=1e-15, abs=0), name
    assert np.allclose(A.array("q"), s.q, rtol=1e-15, atol=0) and np.array_equal(A.array("r"), s.r)
    up, np2, cc, uf = setups.shock_slab(s, 2, ny + 1, 2, nz + 1 if dim == 3 else 2, seed=SEED)
    assert np.array_equal(A.array("np2", np.int32), np2) and np.array_equal(A.array("cumcnt", np.int32), cc)
    assert np.abs(A.array("uf") - uf).max() <= 4e-16 * np.abs(uf).max()            # tanh of vprofile: numpy vs libm
    m = active_mask(np2, s.np_cap)
    got, want = A.array("up")[m], up[m]
    assert np.array_equal(got[:, -1].view(np.int64), want[:, -1].view(np.int64))      # set_particle_ids
    assert np.array_equal(got[:, :dim], want[:, :dim])                                # placement: bit for bit
    assert np.abs(got[:, dim:-1] - want[:, dim:-1]).max() <= 1e-15                    # boosted momenta (tanh again)
    assert np.array_equal(A.array("gp")[m].view(np.int64), A.array("up")[m].view(np.int64))


def oracle_world_from_app(A, dim, cfg, nz):
    """an oracle world holding the translated driver's state"""
    nx, ny = cfg["n_x"], cfg["n_y"]
    q, r = A.array("q").copy(), A.array("r").copy()
    W = World2 if dim == 2 else World3
    args = (nx, ny, A.scalar("np").value) if dim == 2 else (nx, ny, nz, A.scalar("np").value)
    w = W(*args, q=q, r=r, bc=2, delt=A.scalar("delt").value)
    w.set_xrange(2, A.scalar("nxe").value)
    for name in ("up", "gp", "uf"):
        w.arr(name)[...] = A.array(name)
    w.arr("np2")[...] = A.array("np2", np.int32)
    w.arr("cumcnt")[...] = A.array("cumcnt", np.int32)
    return w


def same_state(A, w, what):
    assert np.array_equal(w.arr("np2"), A.array("np2", np.int32)), what
    m = active_mask(w.arr("np2"), w.np)
    assert np.array_equal(w.arr("up")[m].view(np.int64), A.array("up")[m].view(np.int64)), what
    assert np.array_equal(w.arr("uf"), A.array("uf")), what
    nxe = A.scalar("nxe").value
    # the reference leaves cumcnt above nxe stale; compare the entries the next push reads (nxs .. nxe)
    assert np.array_equal(w.arr("cumcnt")[..., :nxe - 2 + 1], A.array("cumcnt", np.int32)[..., :nxe - 2 + 1]), what


@pytest.mark.parametrize("dim", [2, 3])
def test_shock_time_loop_with_inject_and_relocate(dim):
    """the whole shock driver loop -- five procedure calls, inject(), relocate() every step (intvl_expand = 1), the box growing
    (2d/proj/shock/app.f90:112-124, 615-852) -- translated reference against the oracle, bit for bit after every stage"""
    if not have(f"shock{dim}d"):
        pytest.skip("the translated driver procedures cannot be built here")
    nz = 3
    A, cfg = shock_app(dim, nz)
    ny, n0, nx = cfg["n_y"], cfg["n_ppc"], cfg["n_x"]
    rows = rows_of(dim, ny, nz)
    npr = n0 * (cfg["n_x_ini"] - 1)
    rng = np.random.default_rng(5)
    A.feed(uniform=rng.random(len(rows) * npr * (dim - 1)), normal=rng.standard_normal(2 * len(rows) * npr * 3))
    A.call("harness__init")
    w = oracle_world_from_app(A, dim, cfg, nz)
    prm = ShockPrm(n0=n0, v0=A.scalar("v0").value, v_thi=cfg["v_thi"], v_the=cfg["v_the"], b0=A.scalar("b0").value,
                   theta_bn=A.scalar("theta_bn").value, phi_bn=A.scalar("phi_bn").value, l_damp_ini=cfg["l_damp_ini"], seed=SEED)
    R = pyref.RefWorld(dim, nx, ny, nz, A.scalar("np").value, q=A.array("q"), r=A.array("r"), delt=A.scalar("delt").value, bc=2, bounds=True)
    for name in ("up", "gp", "uf"):                                      # the translated procedures work on the driver's own arrays
        R.a[0][name] = A.array(name)
    R.a[0]["np2"], R.a[0]["cumcnt"] = A.array("np2", np.int32), A.array("cumcnt", np.int32)
    u0, v0, delt = A.scalar("u0").value, A.scalar("v0").value, A.scalar("delt").value
    pflux = n0 * abs(v0) * delt * 1.0 * (ny if dim == 2 else ny * nz)
    grew = 0
    for it in range(1, 11):
        nxe = A.scalar("nxe").value
        R.set_xrange(2, nxe)
        R.step(pyref.ORDER_SHOCK, u0)
        w.step(2, u0)
        assert w.error() == 0
        same_state(A, w, f"step {it}")
        # ---- inject(): the fraction draw, the two shuffles (identity), then y (and z) per new particle, then three normals per
        # particle and species -- the values of the oracle's keyed streams
        frac = 0.37
        nginj = int(pflux) + (1 if frac < pflux - int(pflux) else 0)
        counts = np.array([nginj // len(rows) + (1 if i < nginj % len(rows) else 0) for i in range(len(rows))], dtype=np.int32)
        if dim == 3:
            # 3d/proj/shock/app.f90: the per-cell counts run over the local (j, k) grid in the same (k outer, j inner) order
            pass
        uni, nrm = [frac], {1: [], 2: []}
        for (j, k, row), n in zip(rows, counts):
            for ii in range(1, n + 1):
                a, b = pyoracle.philox_uniform2(SEED, row, ii, 0, it)
                uni += [a] if dim == 2 else [a, b]
                for isp in (1, 2):
                    nrm[isp] += list(pyoracle.keyed_normals(SEED, row, ii, isp, 0, it))
        A.feed(uniform=uni, normal=nrm[1] + nrm[2], shuffles=[[0], list(range(len(rows)))])
        A.call("inject")
        assert A.leftover() == (0, 0, 0), "inject() drew fewer values than predicted"
        w.shock_inject(prm, counts, it)
        assert w.error() == 0
        same_state(A, w, f"inject {it}")
        # ---- relocate()
        uni, nrm = [], {1: [], 2: []}
        if nxe < nx + 1:
            for j, k, row in rows:
                for ii in range(1, n0 + 1):
                    a, b = pyoracle.philox_uniform2(SEED, row, ii, 16, it)
                    uni += [a] if dim == 2 else [a, b]
                    for isp in (1, 2):
                        nrm[isp] += list(pyoracle.keyed_normals(SEED, row, ii, isp, 16, it))
            grew += 1
        A.feed(uniform=uni, normal=nrm[1] + nrm[2])
        A.call("relocate")
        assert A.leftover() == (0, 0, 0)
        w.shock_relocate(prm, it)
        assert w.error() == 0 and w.nxe_now == A.scalar("nxe").value
        same_state(A, w, f"relocate {it}")
    assert grew >= 5 and int(A.array("np2", np.int32).sum()) > 2 * len(rows) * npr


# ---------------------------------------------------------------------------------------------------------------------------
# Weibel: loader, IDs, energy_history
# ---------------------------------------------------------------------------------------------------------------------------
WEIBEL_CFG = dict(num_process=1, n_ppc=5, n_x=10, n_y=6, mass_ratio=4.0, sigma_e=0.04, omega_pe=0.1, v_the=0.1, v_thi=0.05, t_ani=5.0)


def weibel_app(dim, nz=4):
    cfg = dict(WEIBEL_CFG)
    if dim == 3:
        cfg.update(num_process_j=1, n_z=nz)
    A = pyref.RefApp(f"weibel{dim}d")
    ny = cfg["n_y"]
    A.configure([0, 2, ny + 1] if dim == 2 else [0, 2, ny + 1, 2, nz + 1, 0, 0], **cfg)
    return A, cfg


@pytest.mark.parametrize("dim", [2, 3])
def test_weibel_loader_ids_and_energy_history(dim):
    """init constants + set_initial_condition + set_particle_ids (3d/proj/weibel/app.f90:298-338, 391-504; 2d :292-328, 380-474)
    against the oracle's load_weibel, and energy_history (3d :509-577) against the oracle's energy()"""
    if not have(f"weibel{dim}d"):
        pytest.skip("the translated driver procedures cannot be built here")
    nz = 4
    A, cfg = weibel_app(dim, nz)
    nx, ny, n0 = cfg["n_x"], cfg["n_y"], cfg["n_ppc"]
    rows = rows_of(dim, ny, nz)
    npr = n0 * nx
    uni, nrm = [], {1: [], 2: []}
    for j, k, row in rows:
        for ii in range(1, npr + 1):
            a, b = pyoracle.philox_uniform2(SEED, row, ii, 0)
            uni += [a] if dim == 2 else [a, b]
            for isp in (1, 2):
                nrm[isp] += list(pyoracle.keyed_normals(SEED, row, ii, isp, 0))
    A.feed(uniform=uni, normal=nrm[1] + nrm[2])
    A.call("harness__init")
    assert A.leftover() == (0, 0, 0)
    q, r, b0 = pyoracle.weibel_constants(n0, mass_ratio=cfg["mass_ratio"], sigma_e=cfg["sigma_e"], omega_pe=cfg["omega_pe"])
    assert np.allclose(A.array("q"), q, rtol=2e-16, atol=0) and np.array_equal(A.array("r"), r)
    assert A.scalar("b0").value == pytest.approx(b0, rel=2e-16) and A.scalar("np").value == n0 * nx * (5 if dim == 2 else 3)
    W = World2 if dim == 2 else World3
    w = W(*((nx, ny) if dim == 2 else (nx, ny, nz)), A.scalar("np").value, q=A.array("q").copy(), r=A.array("r").copy())
    w.load_weibel(n0, v_thi=cfg["v_thi"], v_the=cfg["v_the"], t_ani=cfg["t_ani"], b0=A.scalar("b0").value, seed=SEED)
    assert np.array_equal(w.arr("np2"), A.array("np2", np.int32)) and np.array_equal(w.arr("cumcnt"), A.array("cumcnt", np.int32))
    assert np.array_equal(w.arr("uf"), A.array("uf"))
    m = active_mask(w.arr("np2"), w.np)
    got, want = A.array("up")[m], w.arr("up")[m]
    assert np.array_equal(got[:, :-1], want[:, :-1])                 # positions and momenta: bit for bit
    # IDs: the reference numbers every species from 1 (set_particle_ids: -(particles before this pencil + i)); the oracle's
    # loader adds (isp - 1) x the species population so that IDs are unique across species -- a label convention of the oracle
    ntot = len(rows) * npr
    ref_ids = -got[:, -1].view(np.int64)
    orc_ids = -want[:, -1].view(np.int64)
    assert np.array_equal(ref_ids[:ntot], orc_ids[:ntot]) and np.array_equal(ref_ids[ntot:], orc_ids[ntot:] - ntot)
    assert np.array_equal(ref_ids[:ntot], np.arange(1, ntot + 1))
    # energy_history: the record written to energy.dat = (it delt, kinetic ion, kinetic electron, E^2/8pi, B^2/8pi, total)
    it = 3
    A.call("energy_history", A.array("up"), A.array("uf"), A.array("np2", np.int32), it)
    rec = (pyref.C.c_double * 16)()
    n = A.L.f90rt_captured(rec, 16)
    e = w.energy()
    total = ((e[0] + e[1]) + e[2]) + e[3]
    assert rec[0] == it * A.scalar("delt").value
    if dim == 3:      # 3d/proj/weibel/app.f90:566-571: both species separately
        assert n == 6 and [rec[1], rec[2], rec[3], rec[4], rec[5]] == [e[0], e[1], e[2], e[3], total]
    else:             # 2d/proj/weibel/app.f90:539-541: the species summed
        assert n == 5 and [rec[1], rec[2], rec[3], rec[4]] == [e[0] + e[1], e[2], e[3], total]
    w.close()


# ---------------------------------------------------------------------------------------------------------------------------
# reconnection: Harris sheet
# ---------------------------------------------------------------------------------------------------------------------------
REC_CFG = dict(num_process=1, n_x=33, n_y=6, mass_ratio=16.0, alpha=2.0, rtemp=0.2, lcs=0.5, nbg=6, ncs=30)


@pytest.mark.parametrize("dim", [2, 3])
def test_harris_sheet_loader_equals_setups_py(dim):
    """init constants + set_initial_condition with its statement functions (b_harris, bx_pert, by_pert, density, jz) and the
    single-precision `sqrt(2.)` (2d/proj/reconnection/app.f90:286-307, 362-452; 3d :298-322, 395-470), then the driver's
    sort__bucket -- against wumingpic_b200.setups.reconnection_constants / reconnection_slab"""
    if not have(f"reconnection{dim}d"):
        pytest.skip("the translated driver procedures cannot be built here")
    nz = 3
    cfg = dict(REC_CFG)
    if dim == 3:
        cfg.update(num_process_j=1, n_z=nz)
    A = pyref.RefApp(f"reconnection{dim}d")
    nx, ny = cfg["n_x"], cfg["n_y"]
    A.configure([0, 2, ny + 1] if dim == 2 else [0, 2, ny + 1, 2, nz + 1, 0, 0], **cfg)
    s = setups.reconnection_constants(nx, ny, nz if dim == 3 else None, mass_ratio=cfg["mass_ratio"], alpha=cfg["alpha"],
                                      rtemp=cfg["rtemp"], lcs=cfg["lcs"], nbg=cfg["nbg"], ncs=cfg["ncs"])
    npr, ibg = s.extra["np_row"], s.extra["ibg"]
    # setups.py draws per pencil: x of the background, x of the sheet, y, (z,) then 3 normals per particle for the ions and for the
    # electrons; the reference interleaves them per particle (x, y[, z]; ion normals, electron normals)
    uni, nrm = [], []
    for j, k, row in rows_of(dim, ny, nz):
        rng = setups._rng(SEED, row)
        xbg, xcs, y = rng.random(ibg), rng.random(npr - ibg), rng.random(npr)
        z = rng.random(npr) if dim == 3 else None
        n1, n2 = setups._normal(rng, 3 * npr).reshape(npr, 3), setups._normal(rng, 3 * npr).reshape(npr, 3)
        x = np.concatenate([xbg, xcs])
        uni.append(np.stack([x, y] + ([z] if dim == 3 else []), axis=1).ravel())
        nrm.append(np.concatenate([n1, n2], axis=1).ravel())
    A.feed(uniform=np.concatenate(uni), normal=np.concatenate(nrm))
    A.call("harness__init")
    assert A.leftover() == (0, 0, 0)
    assert A.scalar("np").value == s.np_cap and int(A.array("np2", np.int32)[0].flat[0]) == npr
    for name, want in (("delt", s.delt), ("b0", s.extra["b0"]), ("vte", s.extra["vte"]), ("vti", s.extra["vti"]), ("x0", s.extra["x0"]),
                       ("y0", s.extra["y0"]), ("lcs", s.extra["lcs"])):
        assert A.scalar(name).value == pytest.approx(want, rel=4e-16, abs=0), name
    assert np.allclose(A.array("q"), s.q, rtol=4e-16, atol=0) and np.array_equal(A.array("r"), s.r)
    # the driver's `call sort__bucket(gp, up, cumcnt, np2, nxs, nxe); up = gp` with the translated sort
    R = pyref.RefWorld(dim, nx, ny, nz, s.np_cap, q=s.q, r=s.r, delt=s.delt, bc=1, bounds=True)
    R.ranks[0].call("sort__bucket", A.array("gp"), A.array("up"), A.array("cumcnt", np.int32), A.array("np2", np.int32), 2, nx + 1)
    up, np2, cc, uf = setups.reconnection_slab(s, 2, ny + 1, 2, nz + 1 if dim == 3 else 2, seed=SEED)
    assert np.array_equal(A.array("np2", np.int32), np2) and np.array_equal(A.array("cumcnt", np.int32), cc)
    assert np.abs(A.array("uf") - uf).max() <= 1e-15 * np.abs(uf).max()               # tanh / exp: numpy vs libm
    m = active_mask(np2, s.np_cap)
    got, want = A.array("gp")[m], up[m]
    assert np.array_equal(got[:, -1].view(np.int64), want[:, -1].view(np.int64))       # same particles in the same sorted order
    assert np.abs(got[:, :dim] - want[:, :dim]).max() <= 4e-15 * nx                    # logistic placement: log / tanh
    assert np.abs(got[:, dim:-1] - want[:, dim:-1]).max() <= 1e-15                     # Maxwellian + drift  f jz(x, y) / density(x)


# ---------------------------------------------------------------------------------------------------------------------------
# paraio: get_particle_count (what io__ptcl / io__orb pack)
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [2, 3])
def test_get_particle_count_equals_the_oracle_packing(dim):
    """3d/common/paraio.f90:1007-1085 (2d/common/paraio.f90): mode 0 packs every active particle, mode 1 the tracers (positive
    64-bit ID bit-cast in the last slot), both in (species, k, j, particle) order, and the per-species counts -> cumsum"""
    if not have(f"pack{dim}d"):
        pytest.skip("the translated procedure cannot be built here")
    from tests.util import make_world2, make_world3
    w = make_world2(9, 6, 5, steps=2) if dim == 2 else make_world3(8, 5, 4, 4, steps=2)
    up, np2 = w.arr("up"), w.arr("np2")
    m = active_mask(np2, w.np)
    ids = up[..., -1].view(np.int64)
    tracer = m & (np.abs(ids) % 7 == 3)
    ids[tracer] = np.abs(ids[tracer])                               # tracers carry positive IDs (app.f90 set_particle_ids)
    A = pyref.RefApp(f"pack{dim}d")
    g = w.geom(0)
    geo = [g["nys"], g["nye"]] + ([g["nzs"], g["nze"]] if dim == 3 else [])
    A.call("harness__set", w.ndim, w.np, 2, *geo, 1)
    for mode in (0, 1):
        buf = np.zeros(int(np2.sum()) * w.ndim + 1)
        cumsum = np.zeros((2, 2), np.int64)                         # cumsum(nproc+1, nsp), nproc = 1
        A.call("get_particle_count", up, np2, buf, cumsum, mode, len(buf))
        rec, lcount = w.pack_particles(mode)
        assert np.array_equal(cumsum[:, 0], [0, 0]) and np.array_equal(cumsum[:, 1], lcount)
        n = int(lcount.sum())
        assert n == (int(tracer.sum()) if mode else int(np2.sum())) and n > 0
        assert np.array_equal(buf[:n * w.ndim].view(np.int64), rec.reshape(-1).view(np.int64))
    w.close()
