"""The build-system side of the drop-in: tools/make_reference_patch.py derives, from a WumingPIC checkout, the patch that swaps the
five common/*.f90 files (and the projects' boundary_*.f90) for the ISO_C_BINDING shim, adds libwuming_b200.so to every project's
link line and the one wm_shim_comm_init call to every driver.  Here the patch is made from /root/reference, applied to a copy, and
the result is checked: it applies cleanly, touches nothing else, and every module / procedure the drivers and the umbrella
modules `use` from the replaced files is provided by the shim under the same name (the check a linker would make)."""
import filecmp
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "f2cxx"))

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "common.mk")), reason="/root/reference is absent")


@pytest.fixture(scope="module")
def patched(tmp_path_factory):
    import make_reference_patch as mp
    work = tmp_path_factory.mktemp("ref")
    tree = work / "WumingPIC"
    shutil.copytree(REF, tree, ignore=shutil.ignore_patterns("*.png", "*.ipynb", ".git"))
    for d, _, fs in os.walk(tree):
        os.chmod(d, 0o755)
        for f in fs:
            os.chmod(os.path.join(d, f), 0o644)
    patch = work / "wuming_b200.patch"
    patch.write_text(mp.make_patch(REF))
    r = subprocess.run(["patch", "-p1", "-d", str(tree), "-i", str(patch)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAILED" not in r.stdout and "fuzz" not in r.stdout
    return tree, [rel for rel, _ in mp.plan(REF)]


def test_patch_touches_only_build_files_and_one_line_per_driver(patched):
    tree, touched = patched
    changed = []
    for d, _, fs in os.walk(tree):
        for f in fs:
            rel = os.path.relpath(os.path.join(d, f), tree)
            if not filecmp.cmp(os.path.join(d, f), os.path.join(REF, rel), shallow=False):
                changed.append(rel)
    assert sorted(changed) == sorted(touched)
    assert all(rel == "common.mk" or rel.endswith("Makefile") or rel.endswith("app.f90") for rel in changed)
    for rel in changed:
        if rel.endswith("app.f90"):
            old, new = open(os.path.join(REF, rel)).read().splitlines(), open(tree / rel).read().splitlines()
            added = [l.strip() for l in new if l not in old]
            assert len(new) == len(old) + 2 and len(added) == 2, rel
            assert added[0] == "use wuming_b200_c, only: wm_shim_comm_init"
            assert re.fullmatch(r"call wm_shim_comm_init\(nproc,(nproc_j,nproc_k|nproc,1),nrank,ncomw\)", added[1]), added[1]
            i = new.index(next(l for l in new if "call wm_shim_comm_init" in l))
            assert re.match(r"\s*call mpi_set__init\(", new[i - 1])                 # right after mpi_set__init, before the __init calls
            assert not any("call wm_shim_comm_init" in l for l in new[:i]) and any("bc__init" in l for l in new[i:])


def test_makefiles_build_the_shim_instead(patched):
    tree, _ = patched
    for dim in (2, 3):
        mk = open(tree / f"{dim}d" / "common" / "Makefile").read()
        srcs = re.search(r"^SRCS\s*=\s*\\\n\t(.*)$", mk, re.M).group(1).split()
        assert f"wuming_b200_shim{dim}d.f90" in srcs and "wuming_b200_c.f90" in srcs and f"wuming{dim}d.f90" in srcs
        assert not {"particle.f90", "field.f90", "sort.f90", "boundary_periodic.f90", "mom_calc.f90"} & set(srcs)
        assert {"mpi_set.f90", "paraio.f90"} <= set(srcs)
        assert srcs.index("wuming_b200_c.f90") < srcs.index(f"wuming_b200_shim{dim}d.f90")
        mods = re.search(r"^MODS\s*=\s*(.*)$", mk, re.M).group(1).split()
        for m in ("particle", "field", "sort", "boundary_periodic", "boundary_reconnection", "boundary_shock", "mom_calc", "wuming_b200_c"):
            assert m + ".mod" in mods
        assert f"wuming_b200_shim{dim}d.o: wuming_b200_c.o" in mk
        for name in os.listdir(tree / f"{dim}d" / "proj"):
            p = open(tree / f"{dim}d" / "proj" / name / "Makefile").read()
            link = next(l for l in p.splitlines() if "$(FC) -o" in l)
            assert link.rstrip().endswith("$(WM_B200_LIBS)") and link.index("_common") < link.index("$(WM_B200_LIBS)")
            assert "boundary_" not in p
    cm = open(tree / "common.mk").read()
    assert "WM_B200_LIBS  = -L$(WM_B200)/wumingpic_b200/lib -lwuming_b200 -Wl,-rpath,$(WM_B200)/wumingpic_b200/lib" in cm


def test_everything_the_drivers_use_is_provided(patched):
    """the link check: every `use <module>[, local => name | only: ...]` of the drivers and of wuming{2,3}d.f90 that points at a
    module the shim replaces finds that module in the shim, and every procedure it names is one of the module's procedures"""
    import f2cxx
    tree, _ = patched
    for dim in (2, 3):
        provided = {}
        for src in ("wuming_b200_c.f90", f"wuming_b200_shim{dim}d.f90"):
            for part in f2cxx.split_modules(open(os.path.join(ROOT, "fortran", src)).read()):
                m = f2cxx.parse_module(part, src, skip=("wm_check",))
                provided[m.name] = {s.name for s in m.subs} | set(m.cfuncs) | set(m.syms)
        replaced = {"particle", "field", "sort", "boundary_periodic", "boundary_reconnection", "boundary_shock", "mom_calc"}
        users = [tree / f"{dim}d" / "common" / f"wuming{dim}d.f90"]
        users += [tree / f"{dim}d" / "proj" / n / "app.f90" for n in os.listdir(tree / f"{dim}d" / "proj")]
        seen = set()
        for u in users:
            for _, st in f2cxx.logical_lines(open(u).read()):
                mu = re.match(r"use\s+([a-z_]\w*)\s*(?:,\s*(.*))?$", st)
                if not mu or mu.group(1) not in replaced | {"wuming_b200_c"}:
                    continue
                mod, rest = mu.group(1), mu.group(2) or ""
                assert mod in provided, (str(u), mod)
                seen.add(mod)
                rest = re.sub(r"^only\s*:", "", rest.strip())
                for item in [x.strip() for x in rest.split(",") if x.strip()]:
                    name = item.split("=>")[-1].strip()
                    assert name in provided[mod], (str(u), mod, name)
        assert {"particle", "field", "sort", "mom_calc", "boundary_periodic", "wuming_b200_c"} <= seen
        assert {"boundary_reconnection", "boundary_shock"} <= seen


def test_resident_patch_puts_the_sync_points_where_the_drivers_read_or_edit_the_host_arrays(tmp_path):
    import make_reference_patch as mp
    tree = tmp_path / "WumingPIC"
    shutil.copytree(REF, tree, ignore=shutil.ignore_patterns("*.png", "*.ipynb", ".git"))
    for d, _, fs in os.walk(tree):
        os.chmod(d, 0o755)
        for f in fs:
            os.chmod(os.path.join(d, f), 0o644)
    patch = tmp_path / "resident.patch"
    patch.write_text(mp.make_patch(REF, resident=True))
    r = subprocess.run(["patch", "-p1", "-d", str(tree), "-i", str(patch)], capture_output=True, text=True)
    assert r.returncode == 0 and "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout + r.stderr
    sync = "call wm_shim_sync_to_host(up,uf,np2,cumcnt)"
    n_drivers = 0
    for dim in (2, 3):
        for name in os.listdir(tree / f"{dim}d" / "proj"):
            n_drivers += 1
            new = [l.strip() for l in open(tree / f"{dim}d" / "proj" / name / "app.f90").read().splitlines()]
            old = [l.strip() for l in open(os.path.join(REF, f"{dim}d", "proj", name, "app.f90")).read().splitlines()]
            a, b = new.index("subroutine app__main()"), new.index("end subroutine app__main")
            main = new[a:b]
            assert main[main.index("call init()") + 1] == "call wm_shim_set_mode(WM_SHIM_RESIDENT)"
            # every reader of the host arrays inside the loop is preceded by the sync, every writer is bracketed
            for i, l in enumerate(main):
                if re.match(r"call (io__ptcl|io__orb|mom_calc__accl|save_restart)\(", l):
                    assert main[i - 1] == sync, (name, l)
                if l in ("call inject()", "call relocate()"):
                    assert main[i - 1] == sync and main[i + 1] == "call wm_shim_host_modified()", (name, l)
            assert sum(l == sync for l in main) == sum(bool(re.match(r"call (io__ptcl|io__orb|mom_calc__accl|save_restart)\(|"
                                                                         r"call (inject|relocate)\(\)", l)) for l in main)
            assert ("call inject()" in main) == (name == "shock")
            # the five calls of the time loop are untouched, and outside app__main only the `use` and the comm_init line were added
            removed = [l for l in old if l not in new]
            assert removed == [], (name, removed)
            outside = new[:a] + new[b:]
            extra = [l for l in outside if l not in old]
            assert len(extra) == 3 and extra[0].startswith("use wuming_b200_c, only:") and extra[1] == "WM_SHIM_RESIDENT" \
                and extra[2].startswith("call wm_shim_comm_init("), (name, extra)
    assert n_drivers == 7
