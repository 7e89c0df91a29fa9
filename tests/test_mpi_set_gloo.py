"""world_size-2 gloo test of the host-side slab logic: every rank derives its slab and neighbours with
SlabLayout (the mpi_set mirror), exchanges them, and the result must tile the domain and agree with
the oracle's in-process rank table."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ny, nz):
    sys.path.insert(0, ROOT)
    from wumingpic_b200 import SlabLayout
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lay = SlabLayout(2, ny + 1, 2, nz + 1, 1, world, rank)
    mine = torch.tensor([lay.nys, lay.nye, lay.nzs, lay.nze, lay.jup, lay.jdown, lay.kup, lay.kdown])
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    # ring exchange of a token along kup: what I receive must come from my kdown
    tok = torch.tensor([rank])
    got = torch.zeros_like(tok)
    ops = [dist.P2POp(dist.isend, tok, lay.kup), dist.P2POp(dist.irecv, got, lay.kdown)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    assert int(got) == lay.kdown
    if rank == 0:
        from oracle.pyoracle import World3
        w = World3(4, ny, nz, 8, nproc_j=1, nproc_k=world)
        zs = []
        for r in range(world):
            g = w.geom(r)
            assert [g[k] for k in ("nys", "nye", "nzs", "nze", "jup", "jdown", "kup", "kdown")] == allv[r].tolist()
            zs += list(range(g["nzs"], g["nze"] + 1))
        assert zs == list(range(2, nz + 2))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slabs_gloo():
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 6, 7), nprocs=2, join=True)
