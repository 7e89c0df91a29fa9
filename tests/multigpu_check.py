"""Multi-GPU parity check, launched by torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [--fused 0|1] [--steps 6]

Every rank builds the SAME in-process oracle world with nproc_k = WORLD_SIZE z-slabs (the oracle emulates the
MPI ranks of the reference, 3d/common/mpi_set.f90:45-76, and its MPI_SENDRECV / MPI_ALLREDUCE call sites), uploads
its own slab to its GPU, steps both, and compares its slab: np2 / cumcnt / per-cell particle records (IDs bit-exact),
uf, and the Gauss residual.  Exit code 0 = parity."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--nx", type=int, default=16)
    ap.add_argument("--ny", type=int, default=12)
    ap.add_argument("--nz", type=int, default=10)
    ap.add_argument("--n0", type=int, default=8)
    ap.add_argument("--dim", type=int, default=3, help="3: z-slabs; 2: y-slabs (2d/common/mpi_set.f90:36-47)")
    ap.add_argument("--bc", type=int, default=0, help="0 periodic (Weibel loop), 1 reconnection walls, 2 shock walls")
    ap.add_argument("--yslab", type=int, default=0, help="1 (3-D): decompose along y (nproc_j = WORLD_SIZE, nproc_k = 1) like the reference's "
                                                           "shipped 3-D samples, instead of z-slabs")
    ap.add_argument("--source", type=int, default=0, help="1 (with --bc 2): the shock driver's inject()/relocate() run on the device "
                                                            "after every step (wm_shock_inject / wm_shock_relocate), box grows from nx-6")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tests.util import backend_for, canonical_cells, make_world2, make_world3, rel_err, upload_from_world

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    order, u0 = args.bc, (0.3 if args.bc == 2 else 0.0)     # each boundary module with its own time loop
    if args.source:
        return shock_source_check(args, rank, world, local)
    if args.dim == 3 and args.yslab:
        w = make_world3(args.nx, args.ny, args.nz, args.n0, steps=2, nproc_j=world, nproc_k=1, bc=args.bc, order=order, u0=u0)
        b = backend_for(w, rank=rank, device=local, nproc_j=world, nproc_k=1)
    elif args.dim == 3:
        w = make_world3(args.nx, args.ny, args.nz, args.n0, steps=2, nproc_k=world, bc=args.bc, order=order, u0=u0)
        b = backend_for(w, rank=rank, device=local, nproc_k=world)
    else:
        w = make_world2(args.nx, args.ny, args.n0, steps=2, nproc=world, bc=args.bc, order=order, u0=u0)
        b = backend_for(w, rank=rank, device=local)
    box = [b.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    b.comm_init(world, rank, box[0])
    b.set_fused(bool(args.fused))
    upload_from_world(b, w, rank)
    ntot0 = sum(int(w.arr("np2", r).sum()) for r in range(world))
    worst_uf = 0.0
    for it in range(args.steps):
        w.step(order, u0)
        b.step(2, args.nx + 1, 1, order, u0)
        uf = b.empty("uf")
        b.download(uf=uf)
        worst_uf = max(worst_uf, rel_err(uf, w.arr("uf", rank)))
        res, rho = b.gauss()
        assert res < 1e-13 * max(rho, 1.0), f"rank {rank}: Gauss residual {res} at step {it}"
    assert w.error() == 0
    up, np2, cc = b.empty("up"), b.empty("np2"), b.empty("cumcnt")
    b.download(up, np2, cc)
    assert np.array_equal(np2, w.arr("np2", rank)), f"rank {rank}: np2 differs"
    assert np.array_equal(cc, w.arr("cumcnt", rank)), f"rank {rank}: cumcnt differs"
    worst = 0.0
    for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc),
                                  canonical_cells(w.arr("up", rank), w.arr("np2", rank), w.arr("cumcnt", rank))):
        assert np.array_equal(cg, cr)
        assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64)), f"rank {rank}: particle ID sets differ"
        if len(rg):
            worst = max(worst, np.abs(rg[:, :-1] - rr[:, :-1]).max())
    assert worst < 1e-9 and worst_uf < 1e-8, (worst, worst_uf)
    n = torch.tensor([b.stats()["n_particles"]], device="cuda", dtype=torch.int64)
    dist.all_reduce(n)
    assert int(n.item()) == ntot0, "global particle count not conserved"
    assert b.stats()["error_flags"] == 0
    # moments across slabs (bc__mom folds the slab ghost rows through the same exchanges)
    got = b.mom_calc(2, args.nx + 1)
    w.mom_calc()
    ref = w.arr("mom", rank)
    inner = (slice(None),) + (slice(1, -1),) * args.dim
    assert rel_err(got[inner], ref[inner]) < 1e-9, "moments differ"
    print(f"rank {rank}/{world} ok: dim={args.dim}{' y-slabs' if args.yslab else ''} bc={args.bc} fused={args.fused} np2/cumcnt/IDs exact, max|dx| {worst:.2e}, uf rel {worst_uf:.2e}",
          flush=True)
    b.close()
    dist.destroy_process_group()


def shock_source_check(args, rank, world, local):
    """the shock time loop with the particle source on the device, across slabs"""
    import torch
    import torch.distributed as dist
    import wumingpic_b200 as wm
    from tests.shock_util import (U0, id_first_inject, id_first_relocate, local_rows, make_shock_world, monotone, row_counts,
                                  shock_prm)
    from tests.util import backend_for, canonical_cells, rel_err, upload_from_world
    n0, nxe = args.n0, args.nx - 5
    w = make_shock_world(args.dim, args.nx, args.ny, args.nz, n0, nxe, nproc=world)
    prm = shock_prm(n0)
    prm_c = wm.ShockParams(n0=prm.n0, v0=prm.v0, v_thi=prm.v_thi, v_the=prm.v_the, b0=prm.b0, theta_bn=prm.theta_bn,
                           phi_bn=prm.phi_bn, l_damp_ini=prm.l_damp_ini, seed=prm.seed)
    b = backend_for(w, rank=rank, device=local, nproc_k=world) if args.dim == 3 else backend_for(w, rank=rank, device=local)
    box = [b.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    b.comm_init(world, rank, box[0])
    b.set_fused(bool(args.fused))
    upload_from_world(b, w, rank)
    rows = local_rows(w, rank, args.dim)
    nrows_global = args.ny * (args.nz if args.dim == 3 else 1)
    worst = 0.0
    for it in range(1, args.steps + 1):
        w.step(2, U0)
        b.step(2, nxe, 1, 2, U0)
        assert w.error() == 0
        counts = row_counts(nrows_global, it, n0)
        nptotal = sum(w.arr("np2", r).reshape(2, -1).sum(axis=1) for r in range(world))
        w.shock_inject(prm, counts, it)
        b.shock_inject(prm_c, nxe, counts[rows], id_first_inject(rows, counts, nptotal), it)
        if it % 2 == 0:
            nptotal = sum(w.arr("np2", r).reshape(2, -1).sum(axis=1) for r in range(world))
            w.shock_relocate(prm, it)
            nxe += 1
            b.shock_relocate(prm_c, nxe, id_first_relocate(rows, n0, nptotal), it)
        up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
        b.download(up, np2, cc, uf)
        assert np.array_equal(np2, w.arr("np2", rank)), f"rank {rank}: np2 differs at step {it}"
        cc_ref = monotone(w.arr("cumcnt", rank))
        assert np.array_equal(cc[..., :nxe], cc_ref[..., :nxe]), f"rank {rank}: cumcnt differs at step {it}"
        assert rel_err(uf, w.arr("uf", rank)) < 1e-8
        for (cg, rg), (cr, rr) in zip(canonical_cells(up, np2, cc), canonical_cells(w.arr("up", rank), w.arr("np2", rank), cc_ref)):
            assert np.array_equal(cg, cr)
            assert np.array_equal(rg[:, -1].view(np.int64), rr[:, -1].view(np.int64)), f"rank {rank}: particle ID sets differ"
            if len(rg):
                worst = max(worst, np.abs(rg[:, :-1] - rr[:, :-1]).max())
    assert worst < 1e-9, worst
    assert b.stats()["error_flags"] == 0
    print(f"rank {rank}/{world} ok: dim={args.dim} shock source on the device, np2/cumcnt/IDs exact, max|dx| {worst:.2e}", flush=True)
    b.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
