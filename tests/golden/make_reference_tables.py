"""Extracts, from the reference tree (read-only /root/reference, present in the build container only), the TABLES the hot
path is built on, into tests/golden/reference_tables.json -- reference-held data the CPU tests pin the oracle and the
product's host logic to (tests/test_reference_tables.py):

  constants   share/module_model_constants.F: the REAL PARAMETERs the path uses, as source text and evaluated in real*4
              with Fortran's left-to-right rules
  halos       Registry/Registry.EM_COMMON: every `halo` / `period` line of dyn_em (name -> [[stencil points | width, fields]])
  rconfig     Registry/Registry.EM_COMMON: default of every namelist variable wrf_b200/shipped.py knows
  flux_operators  dyn_em/module_advect_em.F:202-213: the statement functions flux3 / flux4 / flux5 / flux6 as source text

    python tests/golden/make_reference_tables.py        # rewrites the JSON; commit it together with this script
"""
import json
import os
import re

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
WANT = ["g", "r_d", "cp", "r_v", "cv", "cvpm", "rcp", "p1000mb", "t0", "p0", "cpovcv", "rvovrd", "reradius", "cb", "xlv", "rhowater",
        "svp1", "svp2", "svp3", "svpt0", "ep_2"]


def f32_eval(expr, env, funcs=None):
    """evaluate a Fortran real*4 expression (+ - * / parentheses, names, literals, calls of `funcs`) operation by operation in
    float32, left to right within a precedence level as the standard prescribes"""
    funcs = funcs or {}
    toks = re.findall(r"[A-Za-z_]\w*|\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|[-+*/(),]", expr)
    pos = [0]

    def peek():
        return toks[pos[0]] if pos[0] < len(toks) else None

    def take():
        pos[0] += 1
        return toks[pos[0] - 1]

    def atom():
        t = take()
        if t == "(":
            v = addsub()
            take()
            return v
        if t == "-":
            return np.float32(-atom())
        if t == "+":
            return atom()
        if re.match(r"[A-Za-z_]", t):
            if t.lower() in funcs and peek() == "(":
                take()
                args = [addsub()]
                while peek() == ",":
                    take()
                    args.append(addsub())
                take()
                return np.float32(funcs[t.lower()](*args))
            return env[t.lower()]
        return np.float32(float(t))

    def muldiv():
        v = atom()
        while peek() in ("*", "/"):
            op = take()
            r = atom()
            v = np.float32(v * r) if op == "*" else np.float32(v / r)
        return v

    def addsub():
        v = muldiv()
        while peek() in ("+", "-"):
            op = take()
            r = muldiv()
            v = np.float32(v + r) if op == "+" else np.float32(v - r)
        return v

    return addsub()


def constants():
    env, out = {}, {}
    for ln in open(os.path.join(REF, "share", "module_model_constants.F")):
        m = re.match(r"\s*REAL\s*,\s*PARAMETER\s*::\s*(\w+)\s*=\s*([^!]+)", ln, re.I)
        if not m:
            continue
        name, expr = m.group(1).lower(), m.group(2).strip()
        try:
            env[name] = f32_eval(expr, env)
        except (KeyError, IndexError):
            continue
        if name in WANT:
            out[name] = {"source": expr, "value": float(env[name]), "bits": int(np.float32(env[name]).view(np.uint32))}
    return out


def flux_operators():
    """the statement functions flux3 ... flux6 of advect_u (dyn_em/module_advect_em.F:202-213), continuation lines joined"""
    src = open(os.path.join(REF, "dyn_em", "module_advect_em.F")).read().split("\n")
    out, n = {}, 0
    while n < len(src) and len(out) < 4:
        m = re.match(r"\s*(flux[3456])\(([^)]*)\)\s*=\s*&?\s*$|\s*(flux[3456])\(([^)]*)\)\s*=(.*)$", src[n])
        if m:
            name = m.group(1) or m.group(3)
            args = [a.strip() for a in (m.group(2) or m.group(4)).split(",")]
            text = (m.group(5) or "")
            while text.rstrip().endswith("&") or not text.strip():
                n += 1
                text = text.rstrip().rstrip("&") + " " + src[n].strip()
            out[name] = {"args": args, "expr": " ".join(text.replace("&", " ").split())}
        n += 1
    return out


def halos():
    out = {}
    for ln in open(os.path.join(REF, "Registry", "Registry.EM_COMMON")):
        m = re.match(r"(halo|period)\s+(\w+)\s+dyn_em\s+(\S.*)$", ln)
        if not m:
            continue
        groups = []
        for part in m.group(3).strip().split(";"):
            if ":" not in part:
                continue
            n, fields = part.split(":", 1)
            groups.append([int(n), [f.strip() for f in fields.split(",") if f.strip()]])
        out[m.group(2)] = {"kind": m.group(1), "groups": groups}
    return out


def rconfig(names):
    out = {}
    for ln in open(os.path.join(REF, "Registry", "Registry.EM_COMMON")):
        if not ln.startswith("rconfig"):
            continue
        t = ln.split()
        if len(t) >= 6 and t[2] in names:
            out[t[2]] = {"type": t[1], "how_set": t[3], "nentries": t[4], "default": t[5]}
    return out


if __name__ == "__main__":
    import sys
    sys.path.insert(0, ROOT)
    from wrf_b200 import shipped
    tables = {"_source": "wrf-model/WRF, files share/module_model_constants.F, Registry/Registry.EM_COMMON, dyn_em/module_advect_em.F", "constants": constants(),
              "halos": halos(), "rconfig": rconfig(set(shipped.REGISTRY_DEFAULTS)), "flux_operators": flux_operators()}
    with open(os.path.join(ROOT, "tests", "golden", "reference_tables.json"), "w") as f:
        json.dump(tables, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in tables.items() if isinstance(v, dict)})
