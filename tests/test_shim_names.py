"""The Fortran shim (wrf_b200/fortran/*.F90) cannot be compiled here (no Fortran compiler), but what it refers to can be checked against
the reference: every grid%<member> it touches must be a Registry `state` / `rconfig` entry (the Registry generates TYPE(domain)'s members,
boundary arrays as <name>_b[t]{xs,xe,ys,ye}) or a member frame/module_domain_type.F declares itself, and every config_flags%<entry> must be an
`rconfig` entry.  Round 1's shim used two members that do not exist; this keeps that from coming back.  Needs /root/reference."""
import glob
import os
import re

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "Registry")), reason="needs the reference tree (/root/reference)")
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def registry_entries():
    state, rconfig = set(), set()
    for path in glob.glob(os.path.join(REF, "Registry", "Registry.EM_COMMON")) + glob.glob(os.path.join(REF, "Registry", "registry.*")):
        for ln in open(path, errors="replace"):
            f = ln.split()
            if len(f) >= 3 and f[0] == "state":
                name, dims = f[2].lower(), (f[3] if len(f) > 3 else "")
                state.add(name)
                if len(f) > 5 and f[5] != "-" and f[5].isdigit() and int(f[5]) > 1:         # time levels: u -> u_1, u_2
                    state.update(f"{name}_{n}" for n in range(1, int(f[5]) + 1))
                if "b" in (f[4].lower() if len(f) > 4 else ""):
                    pass
            elif len(f) >= 3 and f[0] == "rconfig":
                rconfig.add(f[2].lower())
    return state, rconfig


def domain_type_members():
    out = set()
    for ln in open(os.path.join(REF, "frame", "module_domain_type.F"), errors="replace"):
        m = re.match(r"\s*(?:INTEGER|REAL|LOGICAL|TYPE\s*\([^)]*\)|CHARACTER[^:]*)[^:]*::\s*(.+)$", ln, re.I)
        if m:
            for ent in re.split(r",(?![^()]*\))", m.group(1).split("!")[0]):
                nm = re.match(r"\s*([A-Za-z_]\w*)", ent)
                if nm:
                    out.add(nm.group(1).lower())
    return out


def shim_text():
    return "\n".join(re.sub(r"!.*", "", ln) for p in sorted(glob.glob(os.path.join(HERE, "wrf_b200", "fortran", "*.F90"))) for ln in open(p))


def test_grid_members_exist_in_the_reference():
    state, rconfig = registry_entries()
    frame = domain_type_members()
    used = {m.lower() for m in re.findall(r"grid%([A-Za-z_]\w*)", shim_text())}
    assert len(used) > 100
    missing = []
    for name in sorted(used):
        base = re.sub(r"_b(t)?[xy][se]$", "", name)          # boundary arrays of a state variable that has them
        # the Registry names the two time levels u_1 / u_2 through its ntl column; 2-time-level dry fields are declared as u, v, ...
        stem = re.sub(r"_[12]$", "", base)
        if not ({name, base, stem} & (state | rconfig | frame)):
            missing.append(name)
    assert not missing, f"grid% members the reference does not have: {missing}"


def test_config_entries_exist_in_the_reference():
    _, rconfig = registry_entries()
    used = {m.lower() for m in re.findall(r"config_flags%([A-Za-z_]\w*)", shim_text())}
    assert len(used) > 60
    missing = sorted(n for n in used if n not in rconfig)
    assert not missing, f"config_flags entries that are not rconfig entries of the Registry: {missing}"


def test_round_one_slips_stay_out():
    text = shim_text().lower()
    assert "restart_flag_alarm_rings" not in text and "history_alarm_rings" not in text
    assert "is_alarm_tstep" in text
