"""History / restart output straight from the device-resident state.

The reference writes `wrfout_d<nn>_<date>` at the history alarm from the host grid% arrays
(frame/module_integrate.F:406 -> share/mediation_integrate.F med_before_solve_io -> med_hist_out ->
share/module_io_domain.F output_history -> share/output_wrf.F), one netCDF variable per Registry state entry
with an `h` in its I/O string (Registry/Registry.EM_COMMON:158-457), named by the upper-cased dataname, patch
interior only, dimension order (Time, bottom_top[_stag], south_north[_stag], west_east[_stag]).  With io_form 102
every rank writes its own patch to `<name>_<rank:04d>` and records the patch bounds as global attributes
(external/io_netcdf/wrf_io.F90 ext_ncd_open_for_write_begin, frame/module_io_quilt / module_io.F).

Here the state lives in HBM.  `HistoryStream.after_step` asks the solver for a snapshot of the history fields at
the alarm (wrfgpu_snapshot_begin: a device-to-device copy, then the time loop continues while the device-to-host
copy overlaps the next steps) and writes the file when the snapshot has landed -- no synchronous download of
the model state ever sits in the time loop.  Files are netCDF-3 (64-bit offset, what io_form 2 writes) produced
with scipy.io.netcdf_file; layout, names, `MemoryOrder`/`stagger`/`FieldType` attributes and the io_form 102 patch
attributes follow the reference so that its post-processing reads them.

`T` (theta - t0, Registry `th_phy_m_t0`) is a diagnostic the reference computes in phy_prep; it is derived here
from THM and QVAPOR on the host (module_big_step_utilities_em.F:4919-4933) and may differ from the reference's
in the last bit.
"""
from __future__ import annotations

import datetime as _dt

import numpy as np

# (device field, wrfout name, stagger, description, units, streams)  -- Registry.EM_COMMON rows with h / r flags
_STATE = [
    ("u_2", "U", "X", "x-wind component", "m s-1", "hr"),
    ("v_2", "V", "Y", "y-wind component", "m s-1", "hr"),
    ("w_2", "W", "Z", "z-wind component", "m s-1", "hr"),
    ("ph_2", "PH", "Z", "perturbation geopotential", "m2 s-2", "hr"),
    ("phb", "PHB", "Z", "base-state geopotential", "m2 s-2", "hr"),
    ("t_2", "THM", "", "either 1) pert moist pot temp=(1+Rv/Rd Qv)*(theta)-T0, or 2) pert dry pot temp=theta-T0; "
     "based on use_theta_m setting", "K", "hr"),
    ("mu_2", "MU", "", "perturbation dry air mass in column", "Pa", "hr"),
    ("mub", "MUB", "", "base state dry air mass in column", "Pa", "hr"),
    ("p", "P", "", "perturbation pressure", "Pa", "hr"),
    ("pb", "PB", "", "BASE STATE PRESSURE", "Pa", "hr"),
    ("al", "AL", "", "inverse perturbation density", "m3 kg-1", "r"),
    ("alb", "ALB", "", "inverse base density", "m3 kg-1", "r"),
    ("ww", "WW", "Z", "mu-coupled eta-dot", "Pa s-1", "r"),
    ("t_init", "T_INIT", "", "initial potential temperature", "K", "r"),
    ("h_diabatic", "H_DIABATIC", "", "MICROPHYSICS LATENT HEATING", "K s-1", "r"),
    ("moist_qv", "QVAPOR", "", "Water vapor mixing ratio", "kg kg-1", "hr"),
    ("moist_qc", "QCLOUD", "", "Cloud water mixing ratio", "kg kg-1", "hr"),
    ("moist_qr", "QRAIN", "", "Rain water mixing ratio", "kg kg-1", "hr"),
    ("rainnc", "RAINNC", "", "ACCUMULATED TOTAL GRID SCALE PRECIPITATION", "mm", "hr"),
    ("rainncv", "RAINNCV", "", "TIME-STEP NONCONVECTIVE PRECIPITATION", "mm", "r"),
    ("ht", "HGT", "", "Terrain Height", "m", "hr"),
    ("msftx", "MAPFAC_MX", "", "Map scale factor on mass grid, x direction", "", "hr"),
    ("msfty", "MAPFAC_MY", "", "Map scale factor on mass grid, y direction", "", "hr"),
    ("msfux", "MAPFAC_UX", "X", "Map scale factor on u-grid, x direction", "", "hr"),
    ("msfuy", "MAPFAC_UY", "X", "Map scale factor on u-grid, y direction", "", "hr"),
    ("msfvx", "MAPFAC_VX", "Y", "Map scale factor on v-grid, x direction", "", "hr"),
    ("msfvx_inv", "MF_VX_INV", "Y", "Inverse map scale factor on v-grid, x direction", "", "hr"),
    ("msfvy", "MAPFAC_VY", "Y", "Map scale factor on v-grid, y direction", "", "hr"),
    ("f", "F", "", "Coriolis sine latitude term", "s-1", "hr"),
    ("e", "E", "", "Coriolis cosine latitude term", "s-1", "hr"),
    ("sina", "SINALPHA", "", "Local sine of map rotation", "", "hr"),
    ("cosa", "COSALPHA", "", "Local cosine of map rotation", "", "hr"),
    ("znu", "ZNU", "", "eta values on half (mass) levels", "", "hr"),
    ("znw", "ZNW", "Z", "eta values on full (w) levels", "", "hr"),
    ("fnm", "FNM", "", "upper weight for vertical stretching", "", "hr"),
    ("fnp", "FNP", "", "lower weight for vertical stretching", "", "hr"),
    ("rdnw", "RDNW", "", "inverse d(eta) values between full (w) levels", "", "hr"),
    ("rdn", "RDN", "", "inverse d(eta) values between half (mass) levels", "", "hr"),
    ("dnw", "DNW", "", "d(eta) values between full (w) levels", "", "hr"),
    ("dn", "DN", "", "d(eta) values between half (mass) levels", "", "hr"),
    ("c1h", "C1H", "", "half levels, c1h = d bf / d eta, using znw", "Dimensionless", "hr"),
    ("c2h", "C2H", "", "half levels, c2h = (1-c1h)*(p0-pt)", "Pa", "hr"),
    ("c1f", "C1F", "Z", "full levels, c1f = d bf / d eta, using znu", "Dimensionless", "hr"),
    ("c2f", "C2F", "Z", "full levels, c2f = (1-c1f)*(p0-pt)", "Pa", "hr"),
    ("c3h", "C3H", "", "half levels, c3h = bh", "Dimensionless", "hr"),
    ("c4h", "C4H", "", "half levels, c4h = (eta-bh)*(p0-pt)", "Pa", "hr"),
    ("c3f", "C3F", "Z", "full levels, c3f = bf", "Dimensionless", "hr"),
    ("c4f", "C4F", "Z", "full levels, c4f = (eta-bf)*(p0-pt)", "Pa", "hr"),
    ("u_base", "U_BASE", "", "BASE STATE X WIND IN IDEALIZED CASES", "", "r"),
    ("v_base", "V_BASE", "", "BASE STATE Y WIND IN IDEALIZED CASES", "", "r"),
    ("t_base", "T_BASE", "", "BASE STATE T IN IDEALIZED CASES", "K", "r"),
    ("qv_base", "QV_BASE", "", "BASE STATE QV IN IDEALIZED CASES", "", "r"),
    ("z_base", "Z_BASE", "", "BASE STATE HEIGHT IN IDEALIZED CASES", "", "r"),
]
# restart also needs the second time level of the prognostic fields
_RESTART_EXTRA = [("u_1", "U_1", "X"), ("v_1", "V_1", "Y"), ("w_1", "W_1", "Z"), ("ph_1", "PH_1", "Z"), ("t_1", "THM_1", ""), ("mu_1", "MU_1", "")]


def stream_fields(available, stream="h"):
    """The device fields of output stream `stream` ('h' history, 'r' restart) that this solver holds, in Registry order."""
    have = set(available)
    out = [f for f, _, _, _, _, s in _STATE if stream in s and f in have]
    if stream == "r":
        out += [f for f, _, _ in _RESTART_EXTRA if f in have]
        out += [f for f in available if f.startswith(("scalar_", "moist_old_", "scalar_old_")) and "tend" not in f]
    else:
        out += [f for f in available if f.startswith("scalar_") and "old" not in f and "tend" not in f]
    return out


def _field_of(varname):
    for f, name, _, _, _, _ in _STATE:
        if name == varname:
            return f
    for f, name, _ in _RESTART_EXTRA:
        if name == varname:
            return f
    return varname.lower()


def read_fields(path, dims, rec=-1):
    """Read one time record of a wrfout / wrfrst file written by WrfOut back into arrays of the solver's memory shape
    ((nj, nk, ni) / (nj, ni) / (nk,), zero outside the patch).  With a restart file this is the reference's restart read
    (share/mediation_wrfmain.F med_initialdata_input with config_flags%restart, input_restart): feed the arrays to a new
    solver and call start_domain, which refills halos and boundary zones as start_domain_em does on a restart."""
    from scipy.io import netcdf_file
    d = dims
    nj, nk, ni = d.jme - d.jms + 1, d.kme - d.kms + 1, d.ime - d.ims + 1
    out = {}
    nc = netcdf_file(path, "r", mmap=False)
    try:
        i0 = int(getattr(nc, "WEST-EAST_PATCH_START_UNSTAG"))
        j0 = int(getattr(nc, "SOUTH-NORTH_PATCH_START_UNSTAG"))
        for name, v in nc.variables.items():
            if name in ("Times", "T"):
                continue
            a = np.asarray(v[rec], dtype=np.float32)
            if a.ndim == 3:
                full = np.zeros((nj, nk, ni), np.float32)
                k, j, i = a.shape
                full[j0 - d.jms:j0 - d.jms + j, :k, i0 - d.ims:i0 - d.ims + i] = a.transpose(1, 0, 2)
            elif a.ndim == 2:
                full = np.zeros((nj, ni), np.float32)
                j, i = a.shape
                full[j0 - d.jms:j0 - d.jms + j, i0 - d.ims:i0 - d.ims + i] = a
            else:
                full = np.zeros(nk, np.float32)
                full[:a.shape[0]] = a
            out[_field_of(name)] = full
    finally:
        nc.close()
    return out


def time_string(start: str, seconds: float) -> str:
    """WRF date strings are YYYY-MM-DD_hh:mm:ss (19 characters, share/module_date_time.F)."""
    t0 = _dt.datetime.strptime(start, "%Y-%m-%d_%H:%M:%S")
    t = t0 + _dt.timedelta(seconds=float(seconds))
    return f"{t.year:04d}-{t.month:02d}-{t.day:02d}_{t.hour:02d}:{t.minute:02d}:{t.second:02d}"


def _meta(field):
    for f, name, stag, desc, units, _ in _STATE:
        if f == field:
            return name, stag, desc, units
    for f, name, stag in _RESTART_EXTRA:
        if f == field:
            return name, stag, "", ""
    return field.upper(), "", "", ""


class WrfOut:
    """One wrfout / wrfrst file of one rank's patch.  dims: the rank's wrfgpu_dims (patch = ips..jpe);
    split=True names the file `<path>_<rank:04d>` like io_form 102."""

    def __init__(self, path, dims, config, title="OUTPUT FROM WRF_B200 (solve_em on sm_100a)", start="0001-01-01_00:00:00",
                 split=False):
        from scipy.io import netcdf_file
        d = dims
        self.d, self.c = d, config
        if split:
            path = f"{path}_{int(config.mytask):04d}"
        self.path = path
        self.nc = nc = netcdf_file(path, "w", version=2)
        # patch extents: mass points stop at ide-1 / jde-1, staggered dims own the closing face on the last patch
        self.i0, self.i1m, self.i1s = d.ips, min(d.ipe, d.ide - 1), min(d.ipe, d.ide)
        self.j0, self.j1m, self.j1s = d.jps, min(d.jpe, d.jde - 1), min(d.jpe, d.jde)
        nk = d.kde - 1
        nc.createDimension("Time", None)
        nc.createDimension("DateStrLen", 19)
        nc.createDimension("west_east", self.i1m - self.i0 + 1)
        nc.createDimension("south_north", self.j1m - self.j0 + 1)
        nc.createDimension("bottom_top", nk)
        nc.createDimension("bottom_top_stag", nk + 1)
        nc.createDimension("west_east_stag", self.i1s - self.i0 + 1)
        nc.createDimension("south_north_stag", self.j1s - self.j0 + 1)
        nc.TITLE = title
        nc.START_DATE = start
        nc.SIMULATION_START_DATE = start
        setattr(nc, "WEST-EAST_GRID_DIMENSION", np.int32(d.ide))
        setattr(nc, "SOUTH-NORTH_GRID_DIMENSION", np.int32(d.jde))
        setattr(nc, "BOTTOM-TOP_GRID_DIMENSION", np.int32(d.kde))
        nc.DX, nc.DY, nc.DT = np.float32(config.dx), np.float32(config.dy), np.float32(config.dt)
        for a in ("diff_opt", "km_opt", "damp_opt", "mp_physics", "moist_adv_opt", "scalar_adv_opt", "momentum_adv_opt",
                  "zadvect_implicit", "use_theta_m", "hypsometric_opt", "time_step_sound", "rk_ord"):
            setattr(nc, a.upper(), np.int32(getattr(config, a)))
        for a in ("dampcoef", "zdamp", "khdif", "kvdif", "epssm", "smdiv", "emdiv", "p_top"):
            setattr(nc, a.upper(), np.float32(getattr(config, a)))
        # io_form 102 patch attributes (external/io_netcdf/wrf_io.F90)
        for nm, a, b in (("WEST-EAST_PATCH_START_UNSTAG", self.i0, None), ("WEST-EAST_PATCH_END_UNSTAG", self.i1m, None),
                         ("WEST-EAST_PATCH_START_STAG", self.i0, None), ("WEST-EAST_PATCH_END_STAG", self.i1s, None),
                         ("SOUTH-NORTH_PATCH_START_UNSTAG", self.j0, None), ("SOUTH-NORTH_PATCH_END_UNSTAG", self.j1m, None),
                         ("SOUTH-NORTH_PATCH_START_STAG", self.j0, None), ("SOUTH-NORTH_PATCH_END_STAG", self.j1s, None),
                         ("BOTTOM-TOP_PATCH_START_UNSTAG", 1, None), ("BOTTOM-TOP_PATCH_END_UNSTAG", nk, None),
                         ("BOTTOM-TOP_PATCH_START_STAG", 1, None), ("BOTTOM-TOP_PATCH_END_STAG", nk + 1, None)):
            setattr(nc, nm, np.int32(a))
        self.times = nc.createVariable("Times", "c", ("Time", "DateStrLen"))
        self.vars = {}
        self.nrec = 0

    def _slices(self, stag, ndim):
        d = self.d
        js = slice(self.j0 - d.jms, (self.j1s if stag == "Y" else self.j1m) - d.jms + 1)
        is_ = slice(self.i0 - d.ims, (self.i1s if stag == "X" else self.i1m) - d.ims + 1)
        ks = slice(0, d.kde if stag == "Z" else d.kde - 1)
        if ndim == 3:
            return (js, ks, is_)
        if ndim == 2:
            return (js, is_)
        return (ks,)

    def _define(self, name, stag, desc, units, ndim):
        we = "west_east_stag" if stag == "X" else "west_east"
        sn = "south_north_stag" if stag == "Y" else "south_north"
        bt = "bottom_top_stag" if stag == "Z" else "bottom_top"
        if ndim == 3:
            dims, order = ("Time", bt, sn, we), "XYZ"
        elif ndim == 2:
            dims, order = ("Time", sn, we), "XY "
        else:
            dims, order = ("Time", bt), "Z  "
        v = self.nc.createVariable(name, "f", dims)
        v.FieldType = np.int32(104)
        v.MemoryOrder = order
        v.description = desc
        v.units = units
        v.stagger = stag
        self.vars[name] = v
        return v

    def write(self, time_str, fields):
        """Append one time record.  fields: {device field name: array in the solver's memory shape
        (nj, nk, ni) / (nj, ni) / (nk,)} -- what Solver.snapshot_wait() or Solver.get() return."""
        r = self.nrec
        self.times[r] = np.frombuffer(time_str.encode().ljust(19)[:19], dtype="S1")
        for f, a in fields.items():
            name, stag, desc, units = _meta(f)
            a = np.asarray(a)
            v = self.vars.get(name) or self._define(name, stag, desc, units, a.ndim)
            sub = a[self._slices(stag, a.ndim)]
            v[r] = sub.transpose(1, 0, 2) if a.ndim == 3 else sub      # memory (j, k, i) -> file (k, j, i)
        if "t_2" in fields and "T" not in fields:      # T = theta - t0 (th_phy_m_t0), module_big_step_utilities_em.F:4919-4933
            thm = np.asarray(fields["t_2"])
            if int(self.c.use_theta_m) == 1 and "moist_qv" in fields:
                rvovrd = np.float32(461.6 / 287.0)
                t = ((thm + np.float32(300.0)) / (np.float32(1.0) + rvovrd * np.asarray(fields["moist_qv"])) - np.float32(300.0)).astype(np.float32)
            else:
                t = thm
            v = self.vars.get("T") or self._define("T", "", "perturbation potential temperature theta-t0", "K", 3)
            v[r] = t[self._slices("", 3)].transpose(1, 0, 2)
        self.nrec += 1
        self.nc.flush()

    def close(self):
        self.nc.close()


class HistoryStream:
    """The history (or restart) alarm of med_before_solve_io for a device-resident run: every `interval` steps a
    snapshot of the stream's fields is started; it is written out as soon as it has landed (checked after every
    step without blocking) or, at the latest, when the next alarm rings or the stream is closed."""

    def __init__(self, solver, path, interval, stream="h", fields=None, start="0001-01-01_00:00:00", split=False):
        self.s, self.interval, self.start = solver, int(interval), start
        self.fields = list(fields) if fields is not None else stream_fields(solver.names(), stream)
        self.out = WrfOut(path, solver.case.dims, solver.case.config, start=start, split=split)
        self.pending = None      # time string of the snapshot in flight
        self.blocking_waits = 0  # alarms that had to wait for the previous snapshot

    def _flush(self):
        if self.pending is not None:
            self.out.write(self.pending, self.s.snapshot_wait())
            self.pending = None

    def after_step(self, istep):
        """Call after time step `istep` (0 = the initial state, like the reference's first history write)."""
        if self.pending is not None and not self.s.snapshot_busy():
            self._flush()
        if istep % self.interval == 0:
            if self.pending is not None:
                self.blocking_waits += 1
                self._flush()
            self.s.snapshot_begin(self.fields)
            self.pending = time_string(self.start, istep * float(self.s.case.config.dt))

    def close(self):
        self._flush()
        self.out.close()
