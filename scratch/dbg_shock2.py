import sys, numpy as np
sys.path.insert(0, '.')
import wumingpic_b200 as wm
from wumingpic_b200 import setups
from tests.setup_util import world_for
from tests.util import backend_for, rel_err, active_mask
s = setups.shock_constants(1000, 500, 8, None, n_ppc=4, v_the=0.05, v_thi=0.05)
w = world_for(s)
w.arr("gp")[...] = w.arr("up"); w.sort_bucket(); w.arr("gp")[...] = w.arr("up")
b = backend_for(w); b.set_fused(False)
b.upload(w.arr("up"), w.arr("np2"), w.arr("cumcnt"), w.arr("uf"))
w.particle_solv(); b.particle__solv(s.nxs, s.nxe)
gp = b.empty("gp"); b.download(gp=gp)
m = active_mask(w.arr("np2"), w.np)
ref, got = w.arr("gp")[m], gp[m]
up = w.arr("up")[m]
d = np.abs(got[:, :5] - ref[:, :5]).max(axis=1)
print('push max abs diff', d.max(), 'at', np.argmax(d), ref[np.argmax(d)], got[np.argmax(d)], up[np.argmax(d)])
dx = ref[:, 0] - up[:, 0]
print('oracle dx range', dx.min(), dx.max(), ' gpu dx range', (got[:,0]-up[:,0]).min(), (got[:,0]-up[:,0]).max())
cell_old = up[:, 0].astype(int); cell_new = ref[:, 0].astype(int)
print('oracle inc range', (cell_new - cell_old).min(), (cell_new - cell_old).max())
w.bc_injection(s.u0); b.bc__injection(s.nxs, s.nxe, s.u0)
b.download(gp=gp); got = gp[m]; ref = w.arr("gp")[m]
print('after injection diff', np.abs(got[:, :5] - ref[:, :5]).max(), 'x range', ref[:,0].min(), ref[:,0].max())
cell_new = ref[:, 0].astype(int)
print('oracle inc range after wall', (cell_new - cell_old).min(), (cell_new - cell_old).max(), 'dy range', (ref[:,1].astype(int)-up[:,1].astype(int)).min(), (ref[:,1].astype(int)-up[:,1].astype(int)).max())
try:
    w.field_fdtd_i(); b.field__fdtd_i(s.nxs, s.nxe); b.sync(); print('field ok', b.stats())
    w.bc_particle_y(); b.bc__particle_yz(); b.sync(); print('yz ok')
    w.sort_bucket(); b.sort__bucket(s.nxs, s.nxe); b.sync(); print('sort ok')
except Exception as e:
    print('FAILED', e, b.stats())
