"""Multi-GPU (z-slab) parity through NCCL: runs tests/multigpu_check.py under torchrun when the box has >= 2 GPUs.
On a 1-GPU box the test is skipped (NCCL cannot place two ranks on one device); the slab logic itself is also
covered without GPUs by tests/test_mpi_set_gloo.py and by the oracle's multi-rank tests."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _keep_log(name, text):
    """the per-rank result lines of a multi-GPU parity run, kept under gpurun_out/ (copied to profiles/ by the builder: the driver's
    round-end GPU test box has one GPU and skips these tests)"""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, f"multigpu_parity_{_ngpu()}gpu_box.log"), "a") as f:
            f.write(f"== {name}\n" + "\n".join(l for l in text.splitlines() if " ok: " in l) + "\n")
    except OSError:
        pass


@pytest.mark.parametrize("fused", [1, 0], ids=["fused", "per-procedure"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_slab_parity(nranks, fused):
    if _ngpu() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(29520 + nranks + fused),
           os.path.join(ROOT, "tests", "multigpu_check.py"), "--fused", str(fused), "--nz", "12" if nranks < 8 else "20"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: ") == nranks
    _keep_log(f"slab_parity_{nranks}ranks_{'fused' if fused else 'per_procedure'}", r.stdout)


@pytest.mark.parametrize("dim,bc", [(2, 0), (2, 1), (3, 1), (3, 2)], ids=["2d-periodic", "2d-reconnection", "3d-reconnection", "3d-shock"])
def test_slab_parity_variants(dim, bc):
    """y-slabs in 2-D and the wall set-ups across slabs (BASELINE.json configs 3-5), 2 ranks, fused step."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + 3 * dim + bc),
           os.path.join(ROOT, "tests", "multigpu_check.py"), "--fused", "1", "--dim", str(dim), "--bc", str(bc),
           "--nx", "18", "--ny", "12", "--nz", "8"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: ") == 2
    _keep_log(f"slab_parity_variant_dim{dim}_bc{bc}_2ranks", r.stdout)


@pytest.mark.parametrize("dim", [3, 2])
def test_slab_shock_source(dim):
    """the shock loop with inject()/relocate() on the device across two slabs"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29590 + dim),
           os.path.join(ROOT, "tests", "multigpu_check.py"), "--fused", "1", "--dim", str(dim), "--bc", "2", "--source", "1",
           "--nx", "22", "--ny", "12", "--nz", "8", "--n0", "6"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: ") == 2


def test_slab_parity_uneven_slabs():
    """nz = 9 over two ranks (slabs of 5 and 4 planes, para_range's remainder rule, 3d/common/mpi_set.f90:81-94): the peer-memory
    cgm addresses a neighbour arena of a different size and shifts its ghost-plane stores by the NEIGHBOUR's slab thickness"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29587",
           os.path.join(ROOT, "tests", "multigpu_check.py"), "--fused", "1", "--nz", "9"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: ") == 2


@pytest.mark.parametrize("bc", [0, 1], ids=["periodic", "reconnection"])
def test_yslab_parity(bc):
    """3-D y-slabs (nproc_j = 2, nproc_k = 1: how every shipped 3-D sample of the reference decomposes, 3d/proj/weibel/config_sample.json
    :12-13) against the oracle's nproc_j = 2 emulation; the device runs them as z-slabs of the relabelled system (tests/test_gpu_yslab.py)"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29620 + bc),
           os.path.join(ROOT, "tests", "multigpu_check.py"), "--fused", "1", "--yslab", "1", "--bc", str(bc),
           "--nx", "18", "--ny", "12", "--nz", "6"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok: ") == 2
    _keep_log(f"yslab_parity_bc{bc}_2ranks", r.stdout)
