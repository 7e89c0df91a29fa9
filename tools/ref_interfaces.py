#!/usr/bin/env python
"""Extracts the PUBLIC procedure interfaces (name, dummy-argument list, and for every dummy its declared type / intent / shape) of the
reference modules the shim replaces, and of the shim itself, as plain data.  Used by tests/test_shim_interfaces.py to diff the two.

  tools/ref_interfaces.py /root/reference > tests/golden/ref_interfaces.json      (interface facts only: names and declarations)
"""
import json
import os
import re
import sys

FILES = {
    "2d": ["2d/common/particle.f90", "2d/common/field.f90", "2d/common/sort.f90", "2d/common/boundary_periodic.f90", "2d/common/mom_calc.f90",
           "2d/proj/reconnection/boundary_reconnection.f90", "2d/proj/shock/boundary_shock.f90"],
    "3d": ["3d/common/particle.f90", "3d/common/field.f90", "3d/common/sort.f90", "3d/common/boundary_periodic.f90", "3d/common/mom_calc.f90",
           "3d/proj/reconnection/boundary_reconnection.f90", "3d/proj/shock/boundary_shock.f90"],
}


def logical_lines(text):
    """free-form Fortran: strip comments, join '&' continuations, lower-case"""
    out, cur = [], ""
    for raw in text.splitlines():
        line = raw.split("!")[0].rstrip()
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1]
            continue
        out.append((cur + s).lower())
        cur = ""
    return out


def split_top(s):
    """split on commas that are not inside parentheses"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def parse_module(text):
    """-> {module_name: {proc_name: {"args": [...], "decl": {arg: "type|intent|shape"}}}} for module-level PUBLIC subroutines"""
    lines = logical_lines(text)
    mods, mod, public, depth, proc = {}, None, set(), 0, None
    for ln in lines:
        m = re.match(r"module\s+(\w+)$", ln)
        if m and not ln.startswith("module procedure"):
            mod, public, depth = m.group(1), set(), 0
            mods[mod] = {}
            continue
        if mod is None:
            continue
        if re.match(r"end\s+module", ln):
            mods[mod] = {k: v for k, v in mods[mod].items() if k in public or not public}
            mod = None
            continue
        m = re.match(r"public\s*::\s*(.*)$", ln)
        if m and depth == 0:
            public.update(x.strip() for x in m.group(1).split(","))
            continue
        if re.match(r"interface\b", ln):
            depth += 100          # interface blocks: their subroutines are dummy-procedure declarations, not module procedures
            continue
        if re.match(r"end\s+interface", ln):
            depth -= 100
            continue
        m = re.match(r"(?:recursive\s+)?subroutine\s+(\w+)\s*\((.*)\)\s*$", ln) or re.match(r"(?:recursive\s+)?subroutine\s+(\w+)\s*$", ln)
        if m:
            depth += 1
            if depth == 1:
                args = [a.strip() for a in (m.group(2).split(",") if m.lastindex and m.lastindex >= 2 and m.group(2).strip() else [])]
                proc = {"args": args, "decl": {}}
                mods[mod][m.group(1)] = proc
            continue
        if re.match(r"end\s+subroutine", ln):
            depth -= 1
            if depth == 0:
                proc = None
            continue
        if proc is not None and depth == 1 and "::" in ln:
            left, right = ln.split("::", 1)
            attrs = [a.strip() for a in split_top(left)]
            typ = attrs[0].replace(" ", "")
            intent = next((a.replace(" ", "") for a in attrs if a.startswith("intent")), "")
            for item in split_top(right):
                m2 = re.match(r"(\w+)\s*(\(.*\))?", item)
                if m2 and m2.group(1) in proc["args"]:
                    proc["decl"][m2.group(1)] = f"{typ}|{intent}|{(m2.group(2) or '').replace(' ', '')}"
        if proc is not None and depth == 1 and re.match(r"external\b", ln):
            for name in re.split(r"[,\s:]+", ln[len("external"):]):
                if name in proc["args"]:
                    proc["decl"][name] = "external||"
    return mods


def extract(root, files):
    out = {}
    for f in files:
        with open(os.path.join(root, f), errors="replace") as fh:
            for mod, procs in parse_module(fh.read()).items():
                out[mod] = {"file": f, "procs": procs}
    return out


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    json.dump({dim: extract(root, fl) for dim, fl in FILES.items()}, sys.stdout, indent=1, sort_keys=True)
