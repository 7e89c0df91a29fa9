"""Struct layouts of the C ABI (include/wuming_b200.h) against their ctypes mirrors (wumingpic_b200.backend) and the oracle's
ShockPrm: a tiny C program compiled with the system compiler prints sizeof / offsetof, which must equal ctypes' view.  Catches a
field added on one side only -- the Fortran shim's bind(c) types follow the same header."""
import ctypes as C
import os
import subprocess

from oracle.pyoracle import ShockPrm
from wumingpic_b200 import backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r'''
#include <stddef.h>
#include <stdio.h>
#include "wuming_b200.h"
#define F(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  printf("wm_params %zu\n", sizeof(wm_params));
  F(wm_params, dim); F(wm_params, ndim); F(wm_params, np); F(wm_params, nsp); F(wm_params, nxgs); F(wm_params, nzge);
  F(wm_params, nys); F(wm_params, nze); F(wm_params, nproc_j); F(wm_params, rank_k); F(wm_params, bc_kind); F(wm_params, device);
  F(wm_params, delx); F(wm_params, gfac); F(wm_params, q); F(wm_params, r);
  printf("wm_stats %zu\n", sizeof(wm_stats));
  F(wm_stats, cg_iterations); F(wm_stats, n_particles); F(wm_stats, max_np2); F(wm_stats, error_flags); F(wm_stats, timed_steps);
  F(wm_stats, ms_push); F(wm_stats, ms_sort);
  printf("wm_shock_params %zu\n", sizeof(wm_shock_params));
  F(wm_shock_params, n0); F(wm_shock_params, v0); F(wm_shock_params, l_damp_ini); F(wm_shock_params, seed);
  return 0;
}
'''


def test_struct_layouts_match(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text(PROG)
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirrors = {"wm_params": backend._Params, "wm_stats": backend._Stats, "wm_shock_params": backend.ShockParams}
    for name, cls in mirrors.items():
        assert int(out[name]) == C.sizeof(cls), name
    for key, val in out.items():
        if "." in key:
            t, f = key.split(".")
            assert getattr(mirrors[t], f).offset == int(val), key
    # the oracle's copy of the shock parameter block has the same layout as the product's
    assert C.sizeof(ShockPrm) == C.sizeof(backend.ShockParams)
    for (n1, _), (n2, _) in zip(ShockPrm._fields_, backend.ShockParams._fields_):
        assert n1 == n2 and getattr(ShockPrm, n1).offset == getattr(backend.ShockParams, n2).offset
