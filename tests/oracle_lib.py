"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def load():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8"], stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_last_error.restype = ctypes.c_char_p
        _LIB.oracle_field_size.restype = ctypes.c_size_t
        _LIB.oracle_field_name.restype = ctypes.c_char_p
    return _LIB


class Oracle:
    PHASES = ("rk_step_prep", "rk_phys_bc_dry_1", "init_zero_tendency", "diffusion_prep", "rk_tendency", "small_step_prep",
              "advance_uv", "advance_mu_t", "advance_w", "sumflux", "calc_p_rho", "small_step_finish",
              "rk_scalar_step", "calc_p_rho_phi", "rk_phys_bc_dry_2", "microphysics", "end_of_step")

    def __init__(self, case):
        self.lib = load()
        self.h = ctypes.c_void_p()
        self.case = case
        self._chk(self.lib.oracle_create(ctypes.byref(case.dims), ctypes.byref(case.config), ctypes.byref(self.h)))
        for k, a in case.fields.items():
            self.set(k, a)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.oracle_last_error().decode())

    def names(self):
        return [self.lib.oracle_field_name(self.h, i).decode() for i in range(self.lib.oracle_num_fields(self.h))]

    def set(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self._chk(self.lib.oracle_set_field(self.h, name.encode(), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.size)))

    def get(self, name):
        n = self.lib.oracle_field_size(self.h, name.encode())
        if n == 0:
            raise KeyError(name)
        a = np.empty(n, np.float32)
        self._chk(self.lib.oracle_get_field(self.h, name.encode(), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n)))
        d = self.case.dims
        ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
        if n == ni * nk * nj:
            return a.reshape(nj, nk, ni)
        if n == ni * nj:
            return a.reshape(nj, ni)
        return a

    def call(self, phase, *args):
        self._chk(getattr(self.lib, "oracle_" + phase)(self.h, *[ctypes.c_int(int(x)) for x in args]))

    def max_cfl(self, reset=False):
        mv, mh = ctypes.c_float(), ctypes.c_float()
        self._chk(self.lib.oracle_max_cfl(self.h, ctypes.byref(mv), ctypes.byref(mh), ctypes.c_int(int(reset))))
        return mv.value, mh.value

    def set_dtbc(self, dtbc):
        self.lib.oracle_set_dtbc(self.h, ctypes.c_float(dtbc))

    def dtbc(self):
        self.lib.oracle_get_dtbc.restype = ctypes.c_float
        return float(self.lib.oracle_get_dtbc(self.h))

    def set_threads(self, n):
        return int(self.lib.oracle_set_threads(self.h, ctypes.c_int(int(n))))

    def step(self, n=1):
        self._chk(self.lib.oracle_step(self.h, ctypes.c_int(n)))

    def bc(self, name, var):
        self._chk(self.lib.oracle_set_physical_bc(self.h, name.encode(), ctypes.c_char(var.encode())))

    def start_domain(self):
        self._chk(self.lib.oracle_start_domain(self.h))

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_destroy(self.h)
        except Exception:
            pass
