"""Slab decomposition and neighbour table -- the host-side logic of the reference's mpi_set
(3d/common/mpi_set.f90:21-97, 2d/common/mpi_set.f90:21-51), without MPI.

Pure Python on purpose: it is exercised on CPU (gloo, world_size 2) and must agree with both the
oracle's in-process rank emulation and the C library's wm_para_range.
"""
from dataclasses import dataclass


def para_range(n1, n2, isize, irank):
    """start/end of a 1-D block decomposition (mpi_set.f90:81-94)."""
    iwork1 = (n2 - n1 + 1) // isize
    iwork2 = (n2 - n1 + 1) % isize
    ns = irank * iwork1 + n1 + min(irank, iwork2)
    ne = ns + iwork1 - 1
    if iwork2 > irank:
        ne += 1
    return ns, ne


@dataclass
class SlabLayout:
    """rank -> (rank_j, rank_k), slab ranges and periodic neighbours.

    3-D: rank = j*nproc_k + k (mpi_set.f90:45-60); 2-D: nproc_k = 1, rank = j.
    """
    nygs: int
    nyge: int
    nzgs: int
    nzge: int
    nproc_j: int
    nproc_k: int
    rank: int

    def __post_init__(self):
        nproc = self.nproc_j * self.nproc_k
        if not 0 <= self.rank < nproc:
            raise ValueError("error in proc no.")  # mpi_set.f90:34-43
        self.rank_j, self.rank_k = divmod(self.rank, self.nproc_k)
        self.nys, self.nye = para_range(self.nygs, self.nyge, self.nproc_j, self.rank_j)
        self.nzs, self.nze = para_range(self.nzgs, self.nzge, self.nproc_k, self.rank_k)
        pj, pk = self.nproc_j, self.nproc_k
        rk = lambda j, k: (j % pj) * pk + (k % pk)  # noqa: E731  periodic in y and z (mpi_set.f90:69-76)
        self.jup, self.jdown = rk(self.rank_j + 1, self.rank_k), rk(self.rank_j - 1, self.rank_k)
        self.kup, self.kdown = rk(self.rank_j, self.rank_k + 1), rk(self.rank_j, self.rank_k - 1)

    @property
    def nyl(self):
        return self.nye - self.nys + 1

    @property
    def nzl(self):
        return self.nze - self.nzs + 1
