#!/usr/bin/env python
"""compare_wrfout.py -- field-by-field comparison of two wrfout / wrfrst files (netCDF-3 classic or 64-bit offset).

The north-star check against the reference: run the same namelist.input + input_sounding through the reference's
ideal.exe + wrf.exe (GNU stanza, real*4; recipe below) and through this repo (tools/run_shipped_case.py writes a wrfout in the
same layout), then

    python tools/compare_wrfout.py wrfout_reference wrfout_ours [--time N] [--tol 1e-5]

prints, per prognostic field, the relative L2 difference measured the way tests/parity.py measures it (perturbation
fields against the full field the model uses: P vs P+PB, PH vs PH+PHB, MU vs MU+MUB) and exits non-zero when a field
exceeds the tolerance (north star: <= 1e-5 after one step; bounded drift reported over 100 steps).

Reference build recipe (needs gfortran, netCDF-Fortran, MPI is optional; none of them exist in this repo's build image):
    export NETCDF=/path/to/netcdf NETCDF_classic=1            # classic files: this script reads netCDF-3 with scipy
    ./configure        # option 32 (serial) or 35 (dm+sm) "GNU (gfortran/gcc)", nesting 0
    sed -i 's/^FCOPTIM.*/FCOPTIM = -O2 -ftree-vectorize -funroll-loops -ffp-contract=off/' configure.wrf   # no FMA contraction,
                                                              # like oracle/Makefile and wrf_b200/csrc/Makefile (-fmad=false)
    ./compile em_quarter_ss                                   # or em_hill2d_x, em_b_wave, em_les
    cd test/em_quarter_ss && ./ideal.exe && ./wrf.exe         # namelist.input as shipped; history_interval in minutes
The same directory then feeds this repo:  python tools/run_shipped_case.py test/em_quarter_ss --steps N --out wrfout_ours
"""
from __future__ import annotations

import argparse
import sys

import numpy as np
from scipy.io import netcdf_file

# prognostic / diagnostic fields of the path and the base field a perturbation is measured against
FIELDS = {"U": None, "V": None, "W": None, "T": 300.0, "PH": "PHB", "MU": "MUB", "P": "PB", "QVAPOR": None, "QCLOUD": None, "QRAIN": None, "TKE": None}


def rel_l2(a, b, base=None):
    a, b = a.astype(np.float64), b.astype(np.float64)
    ref = b if base is None else b + base
    den = np.sqrt((ref * ref).sum())
    num = np.sqrt(((a - b) ** 2).sum())
    return 0.0 if num == 0.0 else (num / den if den > 0 else float("inf"))


def compare(path_ref, path_ours, time=-1, tol=1.0e-5, fields=None, out=sys.stdout):
    """returns {field: (rel_l2, max_abs)} for the fields both files hold; prints a table"""
    fa, fb = netcdf_file(path_ref, "r", mmap=False), netcdf_file(path_ours, "r", mmap=False)
    res, worst = {}, 0.0
    try:
        for name in (fields or FIELDS):
            if name not in fa.variables or name not in fb.variables:
                continue
            a, b = np.array(fa.variables[name][time]), np.array(fb.variables[name][time])
            if a.shape != b.shape:
                raise SystemExit(f"{name}: shapes differ {a.shape} vs {b.shape}")
            base = FIELDS.get(name)
            if isinstance(base, str):
                base = np.array(fa.variables[base][time]).astype(np.float64) if base in fa.variables else None
            r = rel_l2(b, a, base)
            res[name] = (r, float(np.abs(a.astype(np.float64) - b).max()))
            worst = max(worst, r)
            print(f"{name:8s} rel-L2 {r:10.3e}  max|diff| {res[name][1]:10.3e}  {'ok' if r <= tol else 'EXCEEDS ' + str(tol)}", file=out)
    finally:
        fa.close(); fb.close()
    if not res:
        raise SystemExit("no common fields")
    print(f"worst rel-L2 {worst:.3e} (tolerance {tol:g})", file=out)
    return res


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("reference")
    ap.add_argument("ours")
    ap.add_argument("--time", type=int, default=-1, help="record to compare (default: the last one)")
    ap.add_argument("--tol", type=float, default=1.0e-5)
    a = ap.parse_args()
    res = compare(a.reference, a.ours, a.time, a.tol)
    sys.exit(0 if all(r <= a.tol for r, _ in res.values()) else 1)


if __name__ == "__main__":
    main()
