#!/usr/bin/env python
"""bench.py -- particle-updates/s of the whole PIC step (push + deposit + field solve + boundary/migration
+ sort) on N B200s of one node, plus the roofline of the dominant kernel, the CPU baseline and the
end-to-end (host-buffer) number.  Contract: see the repo's task description; one JSON line on rank 0.

  python bench.py [--gpus N --steps K --warmup W]          this backend (N > 1: launched by torchrun)
  python bench.py --impl reference [--steps K --warmup W]  the reference's own source on the host cores: oracle/_ref (the Fortran
                                                           hot path translated to C++, flat MPI with one rank per host thread); the
                                                           OpenMP oracle port where oracle/_ref is absent

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): 3-D Weibel, uniform Maxwellian pair plasma, 64 particles per cell
and species, z-slabs across GPUs.  Default = STRONG scaling, the north star's target: the fixed 256 x 256 x 128 box
(1.07 G particles; 120 GB double-buffered, fits one B200) split over the N GPUs as the reference splits a fixed box over its
ranks (3d/common/mpi_set.f90:45-94).  --weak: 256 x 256 x 64 cells PER GPU (536.9 M particles per GPU, nz = 64 N).
Inputs are synthetic (device-side Philox Maxwellian load).  Before the timed region every rank runs the committed golden
case of tests/golden/bench_parity3d.npz (oracle output; no oracle code is imported) through the same wm_step path and reports
the comparison in `checks.parity`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/sec (push+deposit+solve+migrate+sort), 3-D Weibel"
UNIT = "particle-updates/s"
BYTES_PER_UPDATE_STEP = 224       # 4 * ndim * 8 B (push r+w, sort r+w), 3-D
BYTES_PER_UPDATE_PUSH = 112       # the push(+deposit) launch: read up + write gp, 2 * ndim * 8 B


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=256)
    ap.add_argument("--nz", type=int, default=64, help="z cells PER GPU (with --weak)")
    ap.add_argument("--ppc", type=int, default=64, help="particles per cell and species")
    ap.add_argument("--strong-nz", type=int, default=128,
                    help="strong scaling (default): TOTAL z cells of the fixed box split over the GPUs (128 = the 256x256x128 box of "
                         "SURVEY.md 8d)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --nz cells per GPU, the box grows with N")
    ap.add_argument("--no-parity", action="store_true", help="skip the golden-vector parity run before the timed region")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-nz", type=int, default=4, help="z cells of the bounded CPU sample")
    ap.add_argument("--unfused", "--five-calls", dest="unfused", action="store_true",
                    help="drive the reference's five procedure calls per step (particle__solv, field__fdtd_i, bc__particle_x, "
                         "bc__particle_yz, sort__bucket: what an unmodified driver does) instead of the one-call wm_step")
    ap.add_argument("--slow-kernels", action="store_true",
                    help="wm_set_fused(0): the separate per-procedure kernels (push, RED deposit, classify, eager sort)")
    ap.add_argument("--setup", default="weibel", choices=["weibel", "shock", "reconnection"],
                    help="weibel (default): the metric's workload.  shock / reconnection: BASELINE.json configs[2..4] -- the drivers' own loads "
                         "(wumingpic_b200/setups.py) on y-slabs (2-D) / z-slabs (3-D), sizes --nx (shock: box capacity) --ny --nz PER GPU, --ppc "
                         "(shock: n_ppc; reconnection: scales nbg = ppc, ncs = 5 ppc); not the metric's bench line")
    ap.add_argument("--dim", type=int, default=3, choices=[2, 3],
                    help="3 (default): the C2 3-D Weibel workload of the metric; 2: a 2-D Weibel sheet nx x (ny per GPU) for the "
                         "2-D code path (not a bench line of BASELINE.json; ndim = 6 -> 192 B per particle-step)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  nvidia-smi needs up to a second
    before its first line, and a strong-scaling step on 8 GPUs is 18 ms: the sampler is therefore started BEFORE the warm-up and
    every row is stamped on arrival; begin() / end() bracket the timed region and stop() reports the rows that fell inside it.
    If the region was shorter than one sampling period, the rows closest to it (the warm-up steps right before it: same kernels,
    same load) are reported instead, and `window` says so."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    PERIOD_MS = 100

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        # nvidia-smi numbers the physical devices; CUDA's ordinal goes through CUDA_VISIBLE_DEVICES (indices or UUIDs)
        self.index = ids[index] if index < len(ids) else str(index)
        self.rows, self.proc, self.t0, self.t1 = [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.index, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.PERIOD_MS)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [x.strip() for x in line.split(",")]))

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        t0 = self.t0 if self.t0 is not None else float("-inf")
        t1 = self.t1 if self.t1 is not None else time.monotonic()
        if t1 - t0 < 2.5 * self.PERIOD_MS * 1e-3:      # a very short region: wait for the row that was sampled at its end
            time.sleep(1.5 * self.PERIOD_MS * 1e-3)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        ok = [(t, r) for t, r in list(self.rows) if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        rows = [r for t, r in ok if t0 <= t <= t1]
        window = "timed region"
        if not rows and ok:
            # no row landed inside the region: the rows sampled within one period of its ends, else the three closest (warm-up)
            near = [r for t, r in ok if t0 - self.PERIOD_MS * 1e-3 <= t <= t1 + self.PERIOD_MS * 1e-3]
            if near:
                rows, window = near, f"timed region +- {self.PERIOD_MS} ms (the region is shorter than the sampling period)"
            else:
                ok.sort(key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))
                rows, window = [r for _, r in ok[:3]], "closest samples: warm-up steps (no sample fell into the timed region)"
        sm = sorted(float(r[1]) for r in rows)
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window, "period_ms": self.PERIOD_MS}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_memory_available():
    """bytes of host memory this job may still take: min(MemAvailable, cgroup limit - usage); None if unknown"""
    avail = None
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    avail = int(line.split()[1]) * 1024
    except OSError:
        pass
    for lim, use in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                     ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            v = open(lim).read().strip()
            if v != "max":
                room = int(v) - int(open(use).read().strip())
                avail = room if avail is None else min(avail, room)
        except (OSError, ValueError):
            pass
    return avail


def host_threads():
    """host threads this job may really use: the affinity mask, capped by a cgroup CPU quota where one is set (a container that sees 64
    CPUs but may burn 16 CPU-seconds per second would only throttle a 64-thread team)"""
    n = len(os.sched_getaffinity(0))
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            quota, period = open(path).read().split()[:2]
            if quota != "max":
                n = max(1, min(n, int(-(-int(quota) // int(period)))))
        except (OSError, ValueError):
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        p_ = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0 and p_ > 0:
            n = max(1, min(n, -(-q // p_)))
    except (OSError, ValueError):
        pass
    return n


def rank_grid(cores, ny, nz):
    """nproc_j x nproc_k for a flat-MPI run with one rank per host thread: y first (what every shipped 3-D sample does,
    3d/proj/*/config_sample.json), z as well once the y-slabs would get thinner than 2 cells -- the two ghost layers of the
    reference's exchanges need two cells per rank and direction (a 1-cell slab runs but computes something else:
    tests/test_ref_flat_mpi.py) -- and never more ranks than that allows"""
    for nk in (1, 2, 4, 8, 16):
        if cores % nk == 0 and nz % nk == 0 and nz // nk >= 2 and ny // (cores // nk) >= 2:
            return cores // nk, nk
    nj = max(1, min(cores, ny // 2))
    return nj, 1


def cpu_port(args, steps=2, warmup=1):
    """The oracle (a C++ port of the reference loop nests, -O3 -march=native -fopenmp) on the host cores,
    on a bounded sample of the same workload: same nx, ny, ppc and physics, only nz reduced."""
    from oracle import pyoracle
    from oracle.pyoracle import World3, weibel_constants
    import numpy as np
    nx, ny, nz, n0 = args.nx, args.ny, args.cpu_nz, args.ppc
    # every host thread this process may use -- set explicitly: launchers (torch.distributed.run) export OMP_NUM_THREADS=1 and
    # the OpenMP runtime obeys it.  `cores` below is the team a parallel region of the oracle library REALLY gets.
    want = host_threads()
    team = pyoracle.set_num_threads(want, fast=True)
    if team != want or pyoracle.num_threads(fast=True) != want:
        raise RuntimeError(f"oracle OpenMP team is {team} threads, {want} host threads are available: refusing to report a CPU baseline")
    q, r, _ = weibel_constants(n0)
    w = World3(nx, ny, nz, int(n0 * nx * 1.5), q=q, r=r, fast=True)
    w.load_weibel(n0)
    npart = int(w.arr("np2").sum())
    for _ in range(warmup):
        w.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step()
    dt = time.perf_counter() - t0
    assert w.error() == 0
    cores = team
    uf = np.array(w.arr("uf"))
    w.close()
    return {"value": npart * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"3-D Weibel {nx}x{ny}x{nz}, {n0} ppc x 2 species = {npart} particles, {steps} steps "
                      f"({dt:.1f} s) with OMP threads = {cores}; oracle = C++ restatement of the Fortran loop nests "
                      "(no Fortran compiler / MPI in this image)",
            "seconds": dt, "ms_per_step": dt / steps * 1e3, "particles": npart, "parallelism": f"openmp{cores}", "uf": uf}


def cpu_reference(args, steps=2, warmup=1):
    """THE REFERENCE'S OWN SOURCE on the host cores: oracle/_ref = the 3-D hot-path Fortran files translated statement by statement
    to C++ (oracle/f2cxx, pinned bit for bit against the oracle: DESIGN.md 2.1), compiled -O3 -march=native, run as the flat-MPI job
    the reference is (one rank per host thread on the reference's own y x z rank grid, every rank a private copy of the library on
    its own thread; MPI_SENDRECV / MPI_ALLREDUCE rendezvous in oracle/f2cxx/mpi_threads.cpp), on the same bounded sample as
    cpu_port().  The start state is the oracle's Weibel load cut into the ranks' blocks."""
    from oracle import pyoracle
    from oracle.f2cxx import pyref
    from oracle.pyoracle import World3, weibel_constants
    import numpy as np
    nx, ny, nz, n0 = args.nx, args.ny, args.cpu_nz, args.ppc
    cores = host_threads()
    nj, nk = rank_grid(cores, ny, nz)
    q, r, _ = weibel_constants(n0)
    cap = int(n0 * nx * 1.5)
    pyoracle.set_num_threads(cores, fast=True)
    w = World3(nx, ny, nz, cap, nproc_j=nj, nproc_k=nk, q=q, r=r, fast=True)
    w.load_weibel(n0)
    R = pyref.RefWorld(3, nx, ny, nz, cap, nproc_j=nj, nproc_k=nk, q=q, r=r, fast=True, native_mpi=nj * nk > 1)
    # one rank per host CPU, bound to it, and every rank's arrays first touched by that rank (what `mpiexec --bind-to core` gives)
    R.pin_ranks(sorted(os.sched_getaffinity(0)))

    def seed(rk):
        for k in ("up", "uf", "np2", "cumcnt"):
            R.arr(k, rk)[...] = w.arr(k, rk)
        R.arr("gp", rk)[...] = 0.0
    R._all(seed)
    npart = sum(int(w.arr("np2", rk).sum()) for rk in range(nj * nk))
    w.close()
    R.run_steps(warmup)
    t0 = time.perf_counter()
    R.run_steps(steps)
    dt = time.perf_counter() - t0
    after = sum(int(R.arr("np2", rk).sum()) for rk in range(nj * nk))
    if after != npart:
        raise RuntimeError(f"the translated reference lost particles: {npart} -> {after}")
    # the global E, B of the interior cells, for the cross-check against the port (ghost layers differ in what they hold)
    uf = np.zeros((nz + 4, ny + 4, nx + 4, 6))
    for rk in range(nj * nk):
        g = R.geom(rk)
        uf[g["nzs"]:g["nze"] + 1, g["nys"]:g["nye"] + 1, 2:nx + 2] = \
            R.arr("uf", rk)[2:2 + g["nze"] - g["nzs"] + 1, 2:2 + g["nye"] - g["nys"] + 1, 2:nx + 2]
    stats = R.mpi_stats()
    R.close()
    return {"value": npart * steps / dt, "unit": UNIT, "cores": nj * nk, "kind": "reference",
            "sample": f"3-D Weibel {nx}x{ny}x{nz}, {n0} ppc x 2 species = {npart} particles, {steps} steps ({dt:.1f} s): the "
                      f"reference's own Fortran source (3d/common/*.f90) translated to C++ by oracle/f2cxx and compiled with g++ -O3 "
                      f"-march=native (no Fortran compiler / MPI in this image), run as flat MPI on a {nj} x {nk} (y x z) rank grid, one "
                      f"rank per host thread ({cores} available), in-process MPI_SENDRECV / MPI_ALLREDUCE"
                      + (f" ({stats[0] // (steps + warmup) // (nj * nk)} + {stats[1] // (steps + warmup) // (nj * nk)} calls per rank and step)"
                         if stats else ""),
            "seconds": dt, "ms_per_step": dt / steps * 1e3, "particles": npart, "parallelism": f"mpi{nj}x{nk}", "uf": uf}


def cpu_baseline(args, steps=2, warmup=1):
    """The CPU baseline of this workload: the translated reference (kind "reference") where oracle/_ref is present, with the OpenMP
    port timed beside it on the same sample and the two end states compared; the port alone (kind "port") where it is not."""
    import numpy as np
    port = cpu_port(args, steps=min(steps, 2), warmup=1)
    try:
        ref = cpu_reference(args, steps=steps, warmup=warmup)
    except Exception as ex:  # noqa: BLE001  -- no prebuilt oracle/_ref, or g++ failed on this host
        port.pop("uf")
        port["sample"] += f"; the translated reference was not available here ({type(ex).__name__}: {str(ex)[:200]})"
        return port
    nx, ny, nz = args.nx, args.ny, args.cpu_nz
    ref["port"] = {"value": port["value"], "cores": port["cores"], "parallelism": port["parallelism"]}
    if min(steps, 2) + 1 == steps + warmup:         # both legs took the same number of steps from the same load: compare E, B
        a = ref.pop("uf")[2:nz + 2, 2:ny + 2, 2:nx + 2]
        b = port.pop("uf")[2:nz + 2, 2:ny + 2, 2:nx + 2]
        ref["port"]["uf_rel_diff"] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    ref.pop("uf", None)
    return ref


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": (f"3-D Weibel {args.nx}x{args.ny}x{args.nz} cells/GPU" if args.weak else
                                    f"3-D Weibel, fixed {args.nx}x{args.ny}x{args.strong_nz} box") +
                                   f", {args.ppc} ppc x 2 species; the CPU arm runs the bounded sample nz={args.cpu_nz} "
                                   "(same nx, ny, ppc, physics: a per-particle rate)",
                       "parallelism": cb["parallelism"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "port") if k in cb},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def golden_parity(wm, dist, world, rank, local_rank, fused=True, five_calls=False):
    """Every rank runs its z-slab of the committed golden case (tests/golden/bench_parity3d.npz: oracle output, 8 x 6 x 32 cells,
    9216 particles, 3 steps) through the path that is timed below -- wm_step: fused kernel, lazy sort, peer-memory cgm, NCCL halos
    and migration -- and compares with the oracle's end state: fields (relative to the max-norm), np2 / cumcnt and the particle-ID
    sets per cell (exact).  No oracle code runs here: the golden file is its committed output."""
    import numpy as np
    import torch
    from tests import fixture_slabs as fs
    fx = fs.load()
    nx, ny, nz = int(fx["nx"]), int(fx["ny"]), int(fx["nz"])
    lay = wm.SlabLayout(2, ny + 1, 2, nz + 1, 1, world, rank)
    b = wm.Backend(3, int(fx["np_cap"]), 2, nx + 1, 2, ny + 1, 2, nz + 1, nys=lay.nys, nye=lay.nye, nzs=lay.nzs, nze=lay.nze,
                   q=fx["q"], r=fx["r"], nproc_j=1, nproc_k=world, rank_j=0, rank_k=rank, device=local_rank)
    if world > 1:
        box = [b.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        b.comm_init(world, rank, box[0])
    b.set_fused(fused)
    st = fs.slab_state(fx, lay.nzs, lay.nze)
    b.upload(st["up"], st["np2"], st["cumcnt"], st["uf"])
    b.upload_work("df", st["df"])
    worst_gauss = 0.0
    for _ in range(int(fx["steps"])):
        (b.time_loop if five_calls else b.step)(2, nx + 1, 1, wm.backend.WM_ORDER_WEIBEL)
        res, rho = b.gauss()
        worst_gauss = max(worst_gauss, res / max(rho, 1.0))
    up, np2, cc, uf = b.empty("up"), b.empty("np2"), b.empty("cumcnt"), b.empty("uf")
    b.download(up, np2, cc, uf)
    r = fs.compare_slab(fx, lay.nzs, lay.nze, up, np2, cc, uf)
    cg = b.stats()["cg_iterations"]
    flags = b.stats()["error_flags"]
    b.close()
    ok_local = r["np2_equal"] and r["cumcnt_equal"] and r["ids_equal"] and flags == 0
    t = torch.tensor([r["uf_rel_err"], worst_gauss, 0.0 if ok_local else 1.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.tolist()
    return {"case": f"golden 3-D Weibel {nx}x{ny}x{nz}, {len(fx['rec0'])} particles, {int(fx['steps'])} steps, {world} z-slab(s), "
                    f"{'wm_step (fused, lazy sort)' if fused else 'per-procedure kernels'}; oracle output committed under tests/golden/",
            "uf_rel_err_max_over_ranks": t[0], "gauss_rel_residual_max": t[1],
            "np2_cumcnt_id_sets_exact_all_ranks": t[2] == 0.0,
            "cg_iterations": cg, "cg_iterations_oracle": [int(v) for v in fx["cg_1"]],
            "pass": bool(t[2] == 0.0 and t[0] < 1e-9 and t[1] < 1e-13 and list(cg) == [int(v) for v in fx["cg_1"]])}


def run_setup(args):
    """BASELINE.json configs[2..4]: the shock and reconnection set-ups end to end on N GPUs -- the drivers' initial loads (every rank
    builds its own slab, wumingpic_b200/setups.py), then the set-up's own time loop on device-resident state: reconnection = wm_step with
    the reflecting walls before the field solve; shock = wm_step with the injection wall + wm_shock_inject + wm_shock_relocate every
    step (intvl_expand = 1 as in the sample config), the host keeping only the integer bookkeeping of inject()."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import wumingpic_b200 as wm
    from wumingpic_b200 import setups

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dim = args.dim
    ny = args.ny * world if dim == 2 else args.ny
    nz = None if dim == 2 else args.nz * world
    if args.setup == "reconnection":
        s = setups.reconnection_constants(args.nx, ny, nz, nbg=args.ppc, ncs=5 * args.ppc)
        load, cap = setups.reconnection_slab, int(s.extra["np_row"] * 1.3) + 64
    else:
        s = setups.shock_constants(args.nx, args.nx // 2, ny, nz, n_ppc=args.ppc)   # cold upstream, as config_sample.json
        load, cap = setups.shock_slab, int(args.ppc * args.nx * 1.3) + 64
    s.np_cap = cap                     # pencil capacity of the HOST arrays: the drivers allocate n0 nx (x5); only the populated part travels
    if dim == 2:
        lay = wm.SlabLayout(2, ny + 1, 0, 0, world, 1, rank)
        b = wm.Backend(2, cap, 2, s.nx + 1, 2, ny + 1, nys=lay.nys, nye=lay.nye, delt=s.delt, c=s.c, gfac=s.gfac, q=s.q, r=s.r,
                       bc_kind=s.bc, nproc_j=world, nproc_k=1, rank_j=rank, rank_k=0, device=local_rank)
        nzs = nze = 2
    else:
        lay = wm.SlabLayout(2, ny + 1, 2, nz + 1, 1, world, rank)
        b = wm.Backend(3, cap, 2, s.nx + 1, 2, ny + 1, 2, nz + 1, nys=lay.nys, nye=lay.nye, nzs=lay.nzs, nze=lay.nze, delt=s.delt,
                       c=s.c, gfac=s.gfac, q=s.q, r=s.r, bc_kind=s.bc, nproc_j=1, nproc_k=world, rank_j=0, rank_k=rank, device=local_rank)
        nzs, nze = lay.nzs, lay.nze
    if world > 1:
        box = [b.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        b.comm_init(world, rank, box[0])
    t_load = time.perf_counter()
    up, np2, cc, uf = load(s, lay.nys, lay.nye, nzs, nze)
    b.upload(up, np2, cc, uf)
    del up
    t_load = time.perf_counter() - t_load
    stream = torch.cuda.ExternalStream(b.stream(), device=torch.device("cuda", local_rank))
    nrows_glob = s.ny * s.nz
    if dim == 2:
        rows = np.arange(lay.nys - 2, lay.nye - 1)
    else:
        rows = ((np.arange(nzs, nze + 1) - 2)[:, None] * s.ny + (np.arange(lay.nys, lay.nye + 1) - 2)[None, :]).ravel()
    prm = setups.shock_params(s) if args.setup == "shock" else None
    state = {"nxe": s.nxe, "it": 0}

    def count_global():
        t = torch.tensor(b_counts(), device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(t)
        return t.tolist()

    def b_counts():
        n = b.empty("np2")
        b.download(np2=n)
        return n.reshape(2, -1).sum(axis=1).astype(np.int64)

    def one_step():
        state["it"] += 1
        it = state["it"]
        if args.setup == "reconnection":
            b.step(s.nxs, s.nxe, 1, s.order)
            return
        b.step(s.nxs, state["nxe"], 1, s.order, s.u0)
        # inject(): the host's integer bookkeeping (2d/proj/shock/app.f90:711-781); the particle totals it needs for the IDs are
        # tracked from the counts themselves (every injected / relocated particle is known to the host), no read-back per step
        counts = setups.shock_inject_counts(s, it, nproc=world)
        excl = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        idf = np.stack([excl[rows] + state["ntot"][0], excl[rows] + state["ntot"][1]])
        b.shock_inject(prm, state["nxe"], counts[rows], idf, it)
        state["ntot"] = state["ntot"] + int(counts.sum())
        if state["nxe"] < s.nx + 1:
            state["nxe"] += 1
            g = np.asarray(rows, dtype=np.int64)
            b.shock_relocate(prm, state["nxe"], np.stack([g * s.n0 + state["ntot"][0], g * s.n0 + state["ntot"][1]]), it)
            state["ntot"] = state["ntot"] + nrows_glob * s.n0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state["ntot"] = np.array(count_global(), dtype=np.int64)
    n_start = int(state["ntot"].sum())
    e0 = b.energy().sum()
    sampler = ClockSampler(local_rank)
    sampler.start()                       # before the warm-up: nvidia-smi is up and printing when the timed region starts
    for _ in range(args.warmup):
        one_step()
    b.sync()
    b.set_timing(True)
    launches0 = b.launch_count()
    n_before = sum(count_global())
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    ev0.record(stream)
    for _ in range(args.steps):
        one_step()
    b.settle()
    ev1.record(stream)
    barrier()
    sampler.end()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = b.launch_count() - launches0
    st = b.stats()
    b.set_timing(False)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    n_after = sum(count_global())
    per_rank = torch.tensor([float(st["n_particles"])], device="cuda", dtype=torch.float64)
    pr = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(pr, per_rank)
    else:
        pr = [per_rank]
    pr = [float(x.item()) for x in pr]
    res, rho = b.gauss()
    e1 = b.energy().sum()
    et = torch.tensor([e0, e1, res / max(rho, 1.0)], device="cuda", dtype=torch.float64)
    if world > 1:
        tmp = et.clone()
        dist.all_reduce(tmp[:2])
        gm = et[2:].clone()
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        et = torch.cat([tmp[:2], gm])
    et = et.tolist()
    updates = 0.5 * (n_before + n_after) * args.steps            # the shock box fills up while it is timed: mean population
    peak, peak_src = measured_peak()
    bytes_step = 192 if dim == 2 else 224
    k_ms = (st["ms_push"] + st["ms_deposit"]) / max(1, st["timed_steps"])
    npart_rank = max(pr)
    if rank == 0:
        line = {"metric": f"particle-updates/sec (push+deposit+solve+migrate+sort), {dim}-D {args.setup}", "value": updates / (ms * 1e-3),
                "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{dim}-D {args.setup} (BASELINE.json configs[{2 if args.setup == 'shock' else (3 if dim == 2 else 4)}]): "
                                       f"{s.nx} x {s.ny}" + (f" x {s.nz}" if dim == 3 else "") + f" cells, {n_start} particles at the start, "
                                       + (f"Harris sheet, mass ratio {s.r[0]:g}, nbg {s.extra['nbg']}, ncs {s.extra['ncs']}, cfl 0.5"
                                          if args.setup == "reconnection" else
                                          f"n_ppc {s.n0}, u_inject {abs(s.u0):g}, injection wall + inject() + relocate() every step, "
                                          f"box {s.nxe - s.nxs} -> {state['nxe'] - s.nxs} cells"),
                           "parallelism": f"{'y' if dim == 2 else 'z'}-slabs x{world}", "particles": int(n_after),
                           "l2": "particle arrays exceed the 126 MB L2", "loader_s": t_load,
                           "note": "not the metric's bench line (that is --setup weibel); the drivers' loads restated in wumingpic_b200/setups.py"},
                "roofline": {"bound": "hbm", "kernel": "push+deposit", "kernel_ms": k_ms, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "achieved": npart_rank * (bytes_step / 2) / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None,
                             "frac": npart_rank * (bytes_step / 2) / (k_ms * 1e-3) / 1e9 / peak if k_ms > 0 else None, "traffic": None,
                             "step": {"bytes_per_update": bytes_step, "frac": updates / args.steps / world * bytes_step / (ms / args.steps * 1e-3) / 1e9 / peak},
                             "phases_ms": {k: st[k] / max(1, st["timed_steps"]) for k in ("ms_push", "ms_deposit", "ms_field", "ms_sort")}},
                "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
                "checks": {"gauss_rel_residual_max": et[2], "energy_start": et[0], "energy_end": et[1], "cg_iterations": st["cg_iterations"],
                           "error_flags": st["error_flags"], "particles_per_rank": pr,
                           "imbalance_max_over_mean": max(pr) / (sum(pr) / len(pr)),
                           "particles_before_after_timed_region": [int(n_before), int(n_after)]}}
        emit(line)
    b.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else written to fd 1 meanwhile (NCCL's version banner,
    library chatter) was redirected to stderr by quiet_stdout()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.setup != "weibel":
        return run_setup(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import wumingpic_b200 as wm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nx, ny, n0 = args.nx, args.ny, args.ppc
    strong = not args.weak and args.dim == 3 and args.strong_nz > 0
    if strong:                                          # strong scaling: the box is fixed, the slabs shrink
        if args.strong_nz % world:
            raise SystemExit(f"--strong-nz {args.strong_nz} is not a multiple of {world} GPUs")
        args.nz = args.strong_nz // world
    nz_glob = args.nz * world                           # weak scaling: z-slabs, per-GPU work fixed
    parity = None
    if not args.no_parity and args.dim == 3:
        parity = golden_parity(wm, dist, world, rank, local_rank, fused=not args.slow_kernels, five_calls=args.unfused)
    q, r, _ = wm.weibel_constants(n0)
    np_cap = int(n0 * nx * 1.25)
    if args.dim == 3:
        lay = wm.SlabLayout(2, ny + 1, 2, nz_glob + 1, 1, world, rank)
        b = wm.Backend(3, np_cap, 2, nx + 1, 2, ny + 1, 2, nz_glob + 1, nys=lay.nys, nye=lay.nye, nzs=lay.nzs, nze=lay.nze,
                       q=q, r=r, nproc_j=1, nproc_k=world, rank_j=0, rank_k=rank, device=local_rank)
    else:
        ny_glob = ny * world                            # 2-D: y-slabs
        lay = wm.SlabLayout(2, ny_glob + 1, 0, 0, world, 1, rank)
        b = wm.Backend(2, np_cap, 2, nx + 1, 2, ny_glob + 1, nys=lay.nys, nye=lay.nye, q=q, r=r, nproc_j=world, nproc_k=1,
                       rank_j=rank, rank_k=0, device=local_rank)
    if world > 1:
        box = [b.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        b.comm_init(world, rank, box[0])
    b.load_weibel(n0)
    b.sync()
    npart_rank = b.stats()["n_particles"]
    npart = npart_rank * world
    stream = torch.cuda.ExternalStream(b.stream(), device=torch.device("cuda", local_rank))
    order = wm.backend.WM_ORDER_WEIBEL
    if args.slow_kernels:
        b.set_fused(False)      # the separate per-procedure kernels: nothing is deferred or fused
    if args.unfused:            # the driver's five calls per step; on resident state they reach the same fused kernel + lazy sort
        step = lambda n: b.time_loop(2, nx + 1, n, order)  # noqa: E731
    else:
        step = lambda n: b.step(2, nx + 1, n, order)  # noqa: E731

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steady state: W warm-up steps, then exactly K timed steps --------------------
    sampler = ClockSampler(local_rank)
    sampler.start()             # before the warm-up: nvidia-smi is up and printing when the timed region starts
    step(args.warmup)
    b.sync()
    b.set_timing(True)
    launches0 = b.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    ev0.record(stream)
    step(args.steps)
    b.settle()      # the last sort's permutation is applied inside the timed region (wm_step leaves it pending for the next step)
    ev1.record(stream)
    barrier()
    sampler.end()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = b.launch_count() - launches0
    st = b.stats()
    b.set_timing(False)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = npart * args.steps / (ms * 1e-3)
    res, rho = b.gauss()

    # ---- roofline of the dominant kernel (push+deposit), timed live with CUDA events on the library stream ----
    peak, peak_src = measured_peak()
    bytes_push = BYTES_PER_UPDATE_PUSH if args.dim == 3 else 96      # 2 * ndim * 8 B
    bytes_step = BYTES_PER_UPDATE_STEP if args.dim == 3 else 192     # 4 * ndim * 8 B
    k_ms = (st["ms_push"] + st["ms_deposit"]) / max(1, st["timed_steps"])
    achieved = npart_rank * bytes_push / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            # measured DRAM bytes per particle of the kernel (ncu, one capture at C2) x the particles this launch processes
            traffic = tj["push_deposit_dram_bytes_per_particle"] * npart_rank if args.dim == 3 else None
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "push+deposit", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": npart_rank * bytes_push,
                "co_limiters": "ncu (profiles/r01_s4_fused_v6.md; round-2 restructuring experiments: profiles/r02_fused_experiments.md): "
                               "shared-memory wavefronts (LSU data pipe) 77 %, fp64 pipe 48 %, issue slots 52 %, DRAM 24 % -- the kernel is "
                               "bound by shared-memory wavefronts and fp64 issue, not by HBM; measured DFMA peak 1.71e13/s "
                               "(profiles/r01_fp64_peak.json)",
                "kernel_ms": k_ms,
                "step": {"bytes_per_update": bytes_step,
                         "achieved": npart_rank * bytes_step / (ms_per_step * 1e-3) / 1e9,
                         "frac": npart_rank * bytes_step / (ms_per_step * 1e-3) / 1e9 / peak},
                "phases_ms": {k: st[k] / max(1, st["timed_steps"]) for k in ("ms_push", "ms_deposit", "ms_field", "ms_sort")}}

    # ---- end to end through the host-buffer C ABI call: pinned host state in, one step, host state out ----
    e2e = None
    if not args.no_e2e and args.dim == 2:
        args.no_e2e = True   # the end-to-end leg is defined on the 3-D workload of the metric
    if not args.no_e2e:
        # the host-side copy of the state is pinned: all local ranks together must fit the box's host memory
        shp = b.shapes()
        need = world * 8 * (int(np.prod(shp["up"])) + int(np.prod(shp["uf"])))
        avail = host_memory_available()
        if avail is not None and need > 0.6 * avail:
            args.no_e2e = True
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "skipped": f"pinned host state of {world} ranks = {need / 2**30:.0f} GiB exceeds 60 % of the {avail / 2**30:.0f} GiB "
                              "of host memory available on this box"}
    if not args.no_e2e:
        try:
            shp = b.shapes()
            up_t = torch.empty(shp["up"], dtype=torch.float64).pin_memory()
            uf_t = torch.empty(shp["uf"], dtype=torch.float64).pin_memory()
            up, uf = up_t.numpy(), uf_t.numpy()
            np2, cc = b.empty("np2"), b.empty("cumcnt")
            b.download(up, np2, cc, uf)
            h2d = int(np2.sum()) * 7 * 8 + uf.nbytes + np2.nbytes + cc.nbytes
            d2h = h2d
            b.h_step(up, uf, np2, cc, 2, nx + 1, order)           # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                b.h_step(up, uf, np2, cc, 2, nx + 1, order)
            barrier()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            e2e = {"value": npart * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
                   "how": "wm_h_step: full particle+field state H2D from pinned host arrays in the reference layout, "
                          "one step, full state D2H, every step (worst-case drop-in; a driver that only syncs at its "
                          "output cadence runs at `value`)"}
            del up_t, uf_t
        except Exception as ex:  # noqa: BLE001
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "error": f"{type(ex).__name__}: {ex}"}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu and args.dim == 3:
        try:
            cb = cpu_baseline(args)
            cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "port") if k in cb}
        except Exception as ex:  # noqa: BLE001
            cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": (f"3-D Weibel, fixed {nx}x{ny}x{nz_glob} box ({npart} particles) split into {world} z-slab(s) of "
                                        f"{args.nz} planes, {n0} ppc x 2 species ({npart_rank} particles/GPU), periodic, cfl=1, gfac=0.501"
                                        if strong else
                                        f"3-D Weibel {nx}x{ny}x{args.nz} cells/GPU, {n0} ppc x 2 species "
                                        f"({npart_rank} particles/GPU), periodic, cfl=1, gfac=0.501") if args.dim == 3 else
                                       (f"2-D Weibel {nx}x{ny} cells/GPU, {n0} ppc x 2 species ({npart_rank} particles/GPU), "
                                        "periodic (NOT the metric's workload: 2-D code path check)"),
                           "parallelism": f"{'z' if args.dim == 3 else 'y'}-slabs x{world}",
                           "l2": "inputs (particle arrays, >= 30 GB) far exceed the 126 MB L2; no flush needed",
                           "particles": npart, "path": ("five reference procedure calls per step" if args.unfused else "wm_step") +
                                   (", per-procedure kernels (wm_set_fused 0)" if args.slow_kernels else "")},
                "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "checks": {"gauss_residual": res, "max_4pi_rho": rho, "cg_iterations": st["cg_iterations"],
                           "error_flags": st["error_flags"], "parity": parity}}
        emit(line)
    b.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
