"""World-size-2 gloo test of the host-side bookkeeping of the shock driver's inject() (2d/proj/shock/app.f90:711-781): rank 0
draws the per-rank totals and broadcasts them (MPI_Bcast :729), every rank derives its row counts and the ID offsets from the
global cumulative sums (get_global_cumsum :855-880) -- the integers wm_shock_inject takes."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["WM_ROOT"])
from wumingpic_b200 import SlabLayout, inject_counts
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n0, v0, ny = 7, -0.29, 11
lay = [SlabLayout(2, ny + 1, 0, 0, world, 1, r) for r in range(world)]
rows = [l.nye - l.nys + 1 for l in lay]
# (1)-(2) on rank 0, then MPI_Bcast
if rank == 0:
    per_rank, _ = inject_counts(n0, v0, 1.0, 1.0, ny, 1, world, rows, np.random.default_rng(5))
    t = torch.tensor(per_rank)
else:
    t = torch.zeros(world, dtype=torch.int64)
dist.broadcast(t, src=0)
per_rank = t.numpy()
pflux = n0 * abs(v0) * ny
assert per_rank.sum() in (int(pflux), int(pflux) + 1)
assert per_rank.max() - per_rank.min() <= 1
# (3) local rows
rng = np.random.default_rng(100 + rank)
mine = np.full(rows[rank], per_rank[rank] // rows[rank], dtype=np.int32)
mine[rng.permutation(rows[rank])[: per_rank[rank] % rows[rank]]] += 1
assert mine.sum() == per_rank[rank] and mine.max() - mine.min() <= 1
# ID offsets: nptotal from the global cumulative sum of np2, ncinj from the cumulative per-rank totals
np2_local = torch.tensor([[100 + rank, 90 + rank]])          # (1, nsp) particles of this rank
gathered = [torch.zeros_like(np2_local) for _ in range(world)]
dist.all_gather(gathered, np2_local)
nptotal = torch.cat(gathered).sum(dim=0).numpy()
ncinj0 = per_rank[:rank].sum()
id_first = np.stack([ncinj0 + np.concatenate([[0], np.cumsum(mine)[:-1]]) + nptotal[isp] for isp in range(2)])
# the ID ranges of all ranks tile [nptotal+1, nptotal+total] without gaps or overlap
lo = torch.tensor([int(id_first[0, 0]) + 1]); hi = torch.tensor([int(id_first[0, -1]) + int(mine[-1])])
los = [torch.zeros_like(lo) for _ in range(world)]; his = [torch.zeros_like(hi) for _ in range(world)]
dist.all_gather(los, lo); dist.all_gather(his, hi)
los = [int(x) for x in los]; his = [int(x) for x in his]
assert los[0] == nptotal[0] + 1 and his[-1] == nptotal[0] + per_rank.sum()
for r in range(1, world):
    assert los[r] == his[r - 1] + 1
print(f"rank {rank} counts ok", flush=True)
dist.destroy_process_group()
'''


def test_inject_bookkeeping_two_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, WM_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("counts ok") == 2
