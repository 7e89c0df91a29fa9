"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the translated reference as a
flat-MPI job on the host cores; the oracle port where oracle/_ref is absent) prints exactly one JSON line on stdout with the
agreed keys, and the B200 arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--nx", "32", "--ny", "16", "--cpu-nz", "4", "--ppc", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == ("reference" if _ref_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    if d["cpu_baseline"]["kind"] == "reference":
        # the OpenMP port is timed beside the translated reference on the same sample, and the two end states agree
        assert d["cpu_baseline"]["port"]["value"] > 0 and d["cpu_baseline"]["port"]["uf_rel_diff"] < 1e-10
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and "workload" in d["config"]


def _ref_available():
    sys.path.insert(0, ROOT)
    from oracle.f2cxx import build_ref
    return build_ref.build(3) is not None and os.path.exists(os.path.join(build_ref.OUT, "ref3d.cpp"))


def test_reference_arm_falls_back_to_the_port_without_oracle_ref(tmp_path):
    """where oracle/_ref does not exist (and cannot be made: no reference tree), the arm still prints its line, labelled `port`"""
    import shutil
    root = tmp_path / "repo"
    shutil.copytree(ROOT, root, ignore=shutil.ignore_patterns("_ref", ".git", "gpurun_out", "profiles", "*.o", "libwuming_b200.so",
                                                              "golden", "__pycache__", ".pytest_cache"))
    env = dict(os.environ, WUMING_REFERENCE=str(tmp_path / "no_such_tree"))
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--nx", "32", "--ny", "16", "--cpu-nz", "4", "--ppc", "4"], capture_output=True, text=True, timeout=600,
                       cwd=str(root), env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["kind"] == "port" and d["config"]["parallelism"].startswith("openmp")
    assert "translated reference was not available" in d["cpu_baseline"]["sample"]


def test_reference_arm_nonzero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--nx", "16", "--ny", "8",
                        "--nz", "4", "--ppc", "2", "--no-e2e", "--no-cpu"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert r.stdout.strip() == ""            # no JSON line from a run that could not use the GPU


def test_reference_arm_uses_all_host_threads_under_a_launcher():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the reference arm must size its team itself -- one MPI rank per host thread
    for the translated reference, the OpenMP team for the port -- and report what it really ran (an N > 1
    reference arm ran single-threaded while labelled with the core count)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--nx", "32", "--ny", "16", "--cpu-nz", "4", "--ppc", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    sys.path.insert(0, ROOT)
    import bench
    cores = bench.host_threads()
    assert 1 <= cores <= len(os.sched_getaffinity(0))
    if d["cpu_baseline"]["kind"] == "reference":
        nj, nk = bench.rank_grid(cores, 16, 4)
        assert d["cpu_baseline"]["cores"] == nj * nk and d["config"]["parallelism"] == f"mpi{nj}x{nk}"
        assert d["cpu_baseline"]["port"]["cores"] == cores and d["cpu_baseline"]["port"]["parallelism"] == f"openmp{cores}"
    else:
        assert d["cpu_baseline"]["cores"] == cores and d["config"]["parallelism"] == f"openmp{cores}"


def test_rank_grid_uses_every_host_thread_where_the_sample_allows():
    sys.path.insert(0, ROOT)
    import bench
    for cores in (1, 2, 8, 16, 32, 48, 64, 96, 128, 192, 256, 384):
        nj, nk = bench.rank_grid(cores, 256, 4)
        assert nj * nk <= cores and 256 // nj >= 2 and 4 // nk >= 2, (cores, nj, nk)
        if cores <= 256:
            assert nj * nk == cores, (cores, nj, nk)


def _fake_nvidia_smi(tmp_path, startup_s):
    """an `nvidia-smi` that needs `startup_s` before its first row, then prints one row every 100 ms like `-lms 100`"""
    p = tmp_path / "nvidia-smi"
    p.write_text("#!/bin/sh\nsleep %g\nwhile true; do echo '0, 1965, 1965, 612.3, 0x0000000000000000, Not Active, Not Active, "
                 "Not Active, Active'; sleep 0.1; done\n" % startup_s)
    p.chmod(0o755)
    return str(tmp_path)


def test_clock_sampler_windows_rows_on_the_timed_region(tmp_path, monkeypatch):
    """The sampler is started before the warm-up and reports the rows that arrived between begin() and end(); a region shorter
    than the sampling period (18 ms steps on 8 GPUs) falls back to the closest rows and says so -- never `samples: 0` while
    nvidia-smi is printing."""
    import importlib
    import time
    monkeypatch.setenv("PATH", _fake_nvidia_smi(tmp_path, 0.3) + os.pathsep + os.environ["PATH"])
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.8)                 # "warm-up": nvidia-smi starts up meanwhile
    s.begin()
    time.sleep(0.45)
    s.end()
    c = s.stop()
    assert c["window"] == "timed region" and 3 <= c["samples"] <= 6, c
    assert c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.8)
    s.begin()
    time.sleep(0.005)               # a timed region far shorter than the sampling period
    s.end()
    c = s.stop()
    assert c["samples"] >= 1 and c["sm_mhz"] == 1965.0 and c["window"] != "timed region", c


def test_clock_sampler_without_nvidia_smi(tmp_path, monkeypatch):
    import importlib
    monkeypatch.setenv("PATH", str(tmp_path))
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    s = bench.ClockSampler(0)
    s.start()
    s.begin()
    s.end()
    c = s.stop()
    assert c["sm_mhz"] is None and c["samples"] == 0 and c["reasons"] == ["nvidia-smi unavailable"]
