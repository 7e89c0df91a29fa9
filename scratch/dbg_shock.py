import sys, numpy as np
sys.path.insert(0, '.')
import wumingpic_b200 as wm
from wumingpic_b200 import setups
from tests.setup_util import world_for, oracle_shock_prm, id_first_inject, id_first_relocate
from tests.util import backend_for, rel_err
def run(tag, presort, **kw):
    s = setups.shock_constants(kw.pop('nx'), kw.pop('nxi'), kw.pop('ny'), None, **kw)
    w = world_for(s)
    if presort:
        w.arr("gp")[...] = w.arr("up"); w.sort_bucket(); w.arr("gp")[...] = w.arr("up")
    b = backend_for(w)
    b.upload(w.arr("up"), w.arr("np2"), w.arr("cumcnt"), w.arr("uf"))
    if not presort:
        w.arr("gp")[...] = w.arr("up"); w.sort_bucket(); w.arr("gp")[...] = w.arr("up")
    prm_o, prm_c = oracle_shock_prm(s), setups.shock_params(s)
    rows = np.arange(s.ny); nxe = s.nxe
    try:
        for it in range(1, 9):
            w.step(2, s.u0); b.step(s.nxs, nxe, 1, 2, s.u0); b.sync()
            counts = setups.shock_inject_counts(s, it)
            nptotal = w.arr("np2").reshape(2, -1).sum(axis=1)
            w.shock_inject(prm_o, counts, it); b.shock_inject(prm_c, nxe, counts, id_first_inject(rows, counts, nptotal), it)
            nptotal = w.arr("np2").reshape(2, -1).sum(axis=1)
            w.shock_relocate(prm_o, it); nxe += 1
            b.shock_relocate(prm_c, nxe, id_first_relocate(rows, s.n0, nptotal), it); b.sync()
            np2 = b.empty("np2"); uf = b.empty("uf"); b.download(np2=np2, uf=uf)
            print(tag, it, 'np2 equal', np.array_equal(np2, w.arr("np2")), 'uf', rel_err(uf, w.arr("uf")), flush=True)
    except Exception as e:
        print(tag, 'FAILED at', it, e, flush=True)
    b.close(); w.close()
run('u40-presort', True, nx=1000, nxi=500, ny=8, n_ppc=4, v_the=0.05, v_thi=0.05)
run('u40-repair', False, nx=1000, nxi=500, ny=8, n_ppc=4, v_the=0.05, v_thi=0.05)
run('u3-repair', False, nx=128, nxi=24, ny=8, n_ppc=4, u_inject=3.0, v_the=0.05, v_thi=0.05, l_damp_ini=6.0)
run('u40-cold-presort', True, nx=1000, nxi=500, ny=8, n_ppc=4)
