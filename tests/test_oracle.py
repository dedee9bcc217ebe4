"""CPU tests of the oracle (test infrastructure): since the reference ships no golden outputs
for this path, the oracle is held to the invariants the reference's own
regression philosophy and documentation state, to its embedded ADVECT_KERNEL harness, and to
frozen regression checksums (tests/golden/, produced by tests/golden/make_golden.py)."""
import json
import os
import zlib

import numpy as np
import pytest

from tests import parity
from tests.oracle_lib import Oracle
from wrf_b200 import make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def interior(o, name):
    d = o.case.dims
    a = o.get(name)
    if a.ndim == 3:
        return a[d.jds - d.jms:d.jde - d.jms, :d.kde - 1, d.ids - d.ims:d.ide - d.ims]
    return a[d.jds - d.jms:d.jde - d.jms, d.ids - d.ims:d.ide - d.ims]


# ---------------------------------------------------------------- ADVECT_KERNEL harness
@pytest.mark.parametrize("adv_opt", [0, 1])
def test_feeder_tophat(adv_opt):
    """PROGRAM feeder (module_advect_em.F:10567): the top-hat moves at ru/mu = 90 m/s, i.e. 0.45
    cells per loop; mass is conserved; the positive-definite variant stays >= 0 while the plain
    5th-order scheme undershoots; nothing varies in y or z."""
    o = Oracle(parity.tophat_case(adv_opt))
    q0 = interior(o, "scalar_1").astype(np.float64)
    x = np.arange(q0.shape[2], dtype=np.float64)
    m0, c0 = q0[1, 0].sum(), (q0[1, 0] * x).sum() / q0[1, 0].sum()
    nloops = 40
    parity.feeder_loop(o, nloops)
    q = interior(o, "scalar_1").astype(np.float64)
    assert np.isfinite(q).all()
    assert np.ptp(q, axis=0).max() == 0.0 and np.ptp(q, axis=1).max() == 0.0      # 1-D problem stays 1-D, bit for bit
    row = q[1, 0]
    assert abs(row.sum() - m0) <= 1e-4 * m0                                          # flux form conserves
    assert abs((row * x).sum() / row.sum() - (c0 + 0.45 * nloops)) < 0.05            # 18 cells downstream
    assert 0.9 < row.max() < 1.2
    if adv_opt == 1:
        assert row.min() >= -1e-12      # the limiter guard (1e-20) leaves rounding-level negatives
    else:
        assert row.min() < -1e-3                                                     # Gibbs undershoot of the unlimited scheme
    # literal feeder (halo never refreshed): the same away from the edges, up to the dispersive
    # ripples (~1e-8) that the periodic fill lets re-enter from the far side
    o2 = Oracle(parity.tophat_case(adv_opt))
    parity.feeder_loop(o2, nloops, fill_halo=False)
    assert np.abs(interior(o2, "scalar_1")[:, :, 10:80] - interior(o, "scalar_1")[:, :, 10:80]).max() < 1e-6


def test_feeder_column_dump_format():
    """the feeder's output is NINT(field*100) per i (column, :82-112); after 0 loops it is the hat"""
    o = Oracle(parity.tophat_case(0))
    col = np.rint(interior(o, "scalar_1")[1, 0] * 100).astype(int)
    assert col.sum() == 19 * 100 and col[35] == 100 and col[34] == 0 and col[53] == 100 and col[54] == 0


# ---------------------------------------------------------------- BIT4BIT: serial == threaded
@pytest.mark.parametrize("name", ["quarter_ss", "b_wave", "les"])
def test_threads_bit_for_bit(name):
    """tools/regtest.csh:14-18,281 -- serial and OpenMP runs must agree bit for bit; the oracle's
    j-tiles play the role of OpenMP tiles"""
    case = parity.small_cases()[name]
    a, b = Oracle(case), Oracle(case)
    b.set_threads(3)
    for o in (a, b):
        o.start_domain(); o.step(2)
    for n in a.names():
        x, y = a.get(n), b.get(n)
        assert np.array_equal(x, y, equal_nan=True), n


# ---------------------------------------------------------------- documented invariants
def test_2d_case_has_no_y_variation():
    """doc/README.test_cases:43-47 (2-D cases): v stays zero and nothing varies in y"""
    o = Oracle(parity.small_cases()["hill2d_x"])
    o.start_domain(); o.step(5)
    assert np.abs(interior(o, "v_2")).max() == 0.0
    for n in ("u_2", "w_2", "t_2", "ph_2", "mu_2", "p"):
        a = interior(o, n)
        assert np.ptp(a, axis=0).max() == 0.0, n
    assert np.abs(interior(o, "w_2")).max() > 1e-3      # and the hill does force a wave


def test_resting_atmosphere_stays_at_rest():
    """hydrostatically balanced, horizontally uniform, calm state: the discrete pressure-gradient and
    buoyancy terms cancel (module_initialize_ideal.F:1000-1085 balances with the model's own operators)"""
    case = make_case("synthetic", nx=20, ny=18, nz=21, noise=0.0, calm=True, bubble_lattice=1.0e9)
    case.fields["t_1"][:] = case.fields["t_1"][:1, :, :1]   # remove the (single, far away) bubble if any
    o = Oracle(make_case("b_wave", nx=16, ny=20, nz=17, calm=True))
    o.start_domain(); o.step(5)
    assert np.abs(interior(o, "w_2")).max() < 2e-3
    assert np.abs(interior(o, "u_2")).max() < 2e-3


def test_dry_mass_conserved_in_periodic_box():
    """sum of mu over a doubly periodic domain is constant (flux-form continuity, advance_mu_t :1092-1108)"""
    o = Oracle(parity.small_cases()["periodic"])
    o.start_domain()
    m0 = (interior(o, "mu_2").astype(np.float64) + interior(o, "mub")).sum()
    o.step(5)
    m1 = (interior(o, "mu_2").astype(np.float64) + interior(o, "mub")).sum()
    assert abs(m1 - m0) / m0 < 5e-7


def test_positive_definite_moisture():
    """moist_adv_opt = 1 (POSITIVEDEF, Registry.EM_COMMON:3275-3279): qc, qr never go negative"""
    o = Oracle(parity.small_cases()["quarter_ss"])
    o.start_domain(); o.step(10)
    for n in ("moist_qv", "moist_qc", "moist_qr"):
        assert interior(o, n).min() >= 0.0, n


def test_scalar_mass_conserved_by_pd_advection():
    o = Oracle(parity.small_cases()["periodic"])
    o.start_domain()

    def mass(o):
        mu = (interior(o, "mu_2").astype(np.float64) + interior(o, "mub"))[:, None, :]
        dnw = -o.get("dnw")[:o.case.dims.kde - 1].astype(np.float64)[None, :, None]
        return (interior(o, "scalar_1").astype(np.float64) * mu * dnw).sum()
    m0 = mass(o)
    o.step(4)
    assert abs(mass(o) - m0) / m0 < 1e-5
    assert interior(o, "scalar_1").min() >= -1e-12


def test_zonally_uniform_channel_stays_uniform():
    """the unperturbed b_wave channel (periodic x, symmetric y walls, f-plane) has no x variation
    and must not develop any -- the analogue of symm_check (module_solvedebug_em.F:292) for
    this set-up -- and nothing leaks through the symmetric walls (v = 0 on them)"""
    o = Oracle(parity.small_cases()["b_wave"])
    o.start_domain(); o.step(3)
    for n in ("u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2"):
        assert np.ptp(interior(o, n), axis=-1).max() == 0.0, n
    d = o.case.dims
    v = o.get("v_2")
    assert np.abs(v[d.jds - d.jms]).max() == 0.0 and np.abs(v[d.jde - d.jms]).max() == 0.0


# ---------------------------------------------------------------- frozen regression fixtures
def crc(a):
    return zlib.crc32(np.ascontiguousarray(a, dtype=np.float32).tobytes()) & 0xFFFFFFFF


def test_golden_checksums():
    """tests/golden/oracle_step.json freezes the oracle's output after start_domain + 2 steps
    (CRC32 of the bit patterns and an L2 norm per prognostic field) so that an accidental change
    of the restatement is caught.  CRCs are only comparable on the libm family they were made
    with; the norms are checked everywhere."""
    path = os.path.join(ROOT, "tests", "golden", "oracle_step.json")
    gold = json.load(open(path))
    same_libm = os.confstr("CS_GNU_LIBC_VERSION") == gold["_libc"]
    cases = parity.small_cases()
    for name, rec in gold["cases"].items():
        o = Oracle(cases[name])
        o.start_domain(); o.step(2)
        for n, (c, l2) in rec.items():
            a = o.get(n)
            assert abs(float(np.sqrt((a.astype(np.float64) ** 2).sum())) - l2) <= 1e-5 * max(l2, 1e-30), (name, n)
            if same_libm:
                assert crc(a) == c, (name, n)


def test_hill2d_x_mountain_wave_matches_documented_solution():
    """doc/README.test_cases:62-68: em_hill2d_x (10 km half-width, 100 m hill, 10 m/s, N = 0.01/s) "is in
    linear hydrostatic regime, so vertical tilted waves with ~6km vertical wavelength".  Linear theory for
    the bell-shaped hill gives w(x=0, z) = -(U hm / a) sin(N z / U) (times the density factor): amplitude
    U hm / a = 0.1 m/s, zeros every pi U / N = 3.14 km, downward motion in the lowest lobe.  The oracle, run
    on the shipped 202 x 3 x 41 grid for 3 h of model time, must show exactly that below the damping layer
    (zdamp = 20 km under the 30 km lid).  This pins the restatement to the reference's documented physics;
    it does not pin it bit for bit (see oracle.h)."""
    import os
    case = make_case("hill2d_x")
    o = Oracle(case)
    o.set_threads(os.cpu_count() or 1)
    o.start_domain()
    o.step(540)                                     # 3 h at dt = 20 s
    d = case.dims
    j, i = 1 - d.jms, d.ide // 2 - d.ims            # the column over the crest
    z = ((o.get("ph_2")[j, :, i] + o.get("phb")[j, :, i]) / 9.81).astype(np.float64)
    w = o.get("w_2")[j, :, i].astype(np.float64)
    low = z < 10000.0                               # below the Rayleigh layer
    zl, wl = z[low], w[low]
    # zero crossings by linear interpolation
    zc = [zl[k] - wl[k] * (zl[k + 1] - zl[k]) / (wl[k + 1] - wl[k]) for k in range(1, len(zl) - 1) if wl[k] * wl[k + 1] < 0]
    assert len(zc) >= 2, zc
    lam = 2.0 * (zc[1] - zc[0])
    assert 5500.0 <= lam <= 7500.0, lam             # "~6 km"; 2 pi U / N = 6283 m
    assert wl[1:4].min() < -0.05                    # lowest lobe: descent over the crest
    amp = np.abs(wl).max()
    assert 0.07 <= amp <= 0.14, amp                 # U hm / a = 0.1 m/s (grows slowly with height as rho^-1/2)
    # and the 2-D case stays 2-D: no v, no y variation
    assert np.abs(o.get("v_2")).max() == 0.0


def test_constant_k_diffusion_decays_a_sinusoid_at_the_analytic_rate():
    """diff_opt = 1 / km_opt = 1 (module_big_step_utilities_em.F:2715-2950 for a mass-point field): a passive scalar
    sin(2 pi x / L) in a resting, horizontally uniform atmosphere is only touched by horizontal_diffusion('m') with
    K = xkhh = khdif / prandtl = 3 khdif (module_diffusion_em.F:1718-1772, solve_em.F:2230).  The centred operator
    has the eigenvalue -K (2 - 2 cos(2 pi dx / L)) / dx^2, so after n forward-Euler steps the amplitude is
    (1 - dt K (2 - 2 cos(k dx)) / dx^2)^n exactly, up to rounding."""
    nx, L = 33, 32                                  # 32 mass points per period
    case = make_case("b_wave", nx=nx, ny=12, nz=11, calm=True, num_scalar=1, scalar_adv_opt=1,      # periodic x, no bubble, at rest
                     diff_opt=1, km_opt=1, khdif=1.0e6, kvdif=0.0)
    d = case.dims
    ii = np.arange(d.ims, d.ime + 1, dtype=np.float64)
    prof = (1.0 + 0.5 * np.sin(2 * np.pi * (ii - 1) / L)).astype(np.float32)
    s = case.fields["scalar_1"]
    s[:] = prof[None, None, :]
    s[:, -1, :] = 0
    o = Oracle(case)
    o.start_domain()
    n = 20
    o.step(n)
    c = case.config
    K = np.float64(c.khdif) * 3.0
    lam = 1.0 - c.dt * K * (2.0 - 2.0 * np.cos(2 * np.pi / L)) / c.dx ** 2
    got = o.get("scalar_1")[4 - d.jms, 3, 1 - d.ims:1 - d.ims + L].astype(np.float64)
    amp = 2.0 * np.abs(np.fft.rfft(got - got.mean())[1]) / L
    assert abs(amp / 0.5 - lam ** n) < 2e-4 * lam ** n, (amp / 0.5, lam ** n)
    assert abs(got.mean() - 1.0) < 1e-5                               # diffusion conserves the mean
    assert np.abs(o.get("w_2")).max() < 1e-3                          # and the atmosphere stayed at rest


def test_smag2d_k_of_a_sheared_jet_is_the_documented_formula():
    """diff_opt = 1 / km_opt = 4 (smag2d_km, module_diffusion_em.F:1934-2043): K_h = c_s^2 l^2 |D| with l^2 = dx dy and
    D^2 = 0.25 (D11 - D22)^2 + D12^2 (4-point average of D12 to the mass point).  For v = V sin(2 pi x / L) on a flat,
    Cartesian, periodic-x grid D11 = D22 = 0 and D12 = dv/dx (cal_deform_and_div :611-870 without its terrain terms), so with
    the centred differences the code uses K(i) = c_s^2 dx dy |V (sin(k x_{i+1}) - sin(k x_{i-1})) / (2 dx)| -- the x-average
    of the two D12 that flank a mass point; xkhh = K / prandtl = 3 K; and the tendency of v is the divergence of K dv/dx."""
    nx, L, V = 33, 32, 8.0
    case = make_case("b_wave", nx=nx, ny=12, nz=11, calm=True, diff_opt=1, km_opt=4, c_s=0.25)
    d, c = case.dims, case.config
    ii = np.arange(d.ims, d.ime + 1, dtype=np.float64)
    vprof = (V * np.sin(2 * np.pi * (ii - 1) / L)).astype(np.float32)          # v sits at the mass-point x
    for n in ("v_1", "v_2"):
        case.fields[n][:] = vprof[None, None, :]
        case.fields[n][:, -1, :] = 0
    o = Oracle(case)
    o.start_domain()
    for ph, args in (("rk_step_prep", (1,)), ("rk_phys_bc_dry_1", ()), ("init_zero_tendency", ()), ("diffusion_prep", ())):
        o.call(ph, *args)
    j, k = 5 - d.jms, 3
    sl = slice(1 - d.ims, 1 - d.ims + L)
    x = np.arange(L, dtype=np.float64)
    kk = 2 * np.pi / L
    d12_u = V * (np.sin(kk * x) - np.sin(kk * (x - 1))) / c.dx                 # D12 at the u-point west of mass point i
    want = c.c_s ** 2 * c.dx * c.dy * np.abs(0.5 * (d12_u + np.roll(d12_u, -1)))
    got = o.get("xkmh")[j, k, sl].astype(np.float64)
    assert np.allclose(got, want, rtol=2e-5, atol=1e-3 * want.max()), np.abs(got - want).max()
    assert np.allclose(o.get("xkhh")[j, k, sl], 3.0 * got, rtol=1e-6)
    assert np.abs(o.get("defor11")[j, k, sl]).max() < 1e-9 and np.abs(o.get("defor13")).max() == 0.0
    assert want.max() > 5.0e3                                                   # a K that matters on this grid (dx = 100 km)
    # the diffusion tendency it feeds: rv_tendf = d/dx ( mu K_u dv/dx ) with K_u the 4-point average to the u-point
    o.call("rk_tendency", 1)
    mu = float(o.get("mut")[j, 1 - d.ims])
    c1, c2 = float(o.get("c1h")[k]), float(o.get("c2h")[k])
    ku = 0.5 * (got + np.roll(got, 1))
    flux = (c1 * mu + c2) * ku * d12_u
    want_t = (np.roll(flux, -1) - flux) / c.dx
    got_t = o.get("rv_tendf")[j, k, sl].astype(np.float64)
    assert np.allclose(got_t, want_t, rtol=1e-4, atol=2e-4 * np.abs(want_t).max()), np.abs(got_t - want_t).max()


def test_physical_space_diffusion_decays_a_sinusoid_and_a_vertical_profile_at_the_analytic_rates():
    """diff_opt = 2 / km_opt = 1 (horizontal_diffusion_s, vertical_diffusion_s, module_diffusion_em.F:3711, :4789): over flat
    terrain the physical-space operator is d/dx(rho K dq/dx) / rho_hyd with K = xkhh = khdif / prandtl = 3 khdif, so a passive
    sin(2 pi x / L) decays by 1 - dt K (2 - 2 cos(k dx)) / dx^2 per step up to the ratio of the moist density rho = (1 + qv) / alt
    the stresses are built with to the dry hydrostatic mass the tendency is divided by (a few 1e-4 here); the mean is conserved.
    In the vertical (kvdif, xkhv = 3 kvdif) the flux -rho K dq/dz vanishes at the surface and the top, so the column integral of
    the tendency of any profile (here one linear in z) is zero."""
    nx, L = 33, 32
    case = make_case("b_wave", nx=nx, ny=12, nz=11, calm=True, num_scalar=2, scalar_adv_opt=1, diff_opt=2, km_opt=1, khdif=1.0e6, kvdif=30.0)
    d, c = case.dims, case.config
    ii = np.arange(d.ims, d.ime + 1, dtype=np.float64)
    prof = (1.0 + 0.5 * np.sin(2 * np.pi * (ii - 1) / L)).astype(np.float32)
    s = case.fields["scalar_1"]
    s[:] = prof[None, None, :]
    s[:, -1, :] = 0
    o = Oracle(case)
    o.start_domain()
    # scalar_2 = height of the mass level: linear in z
    for ph, args in (("rk_step_prep", (1,)), ("rk_phys_bc_dry_1", ()), ("init_zero_tendency", ()), ("diffusion_prep", ())):
        o.call(ph, *args)
    z = o.get("z").copy()
    z[:, -1, :] = 0
    o = Oracle(case)
    o.set("scalar_2", z * 1.0e-3)
    o.start_domain()
    for ph, args in (("rk_step_prep", (1,)), ("rk_phys_bc_dry_1", ()), ("init_zero_tendency", ()), ("diffusion_prep", ())):
        o.call(ph, *args)
    t2 = o.get("scalar_tend_2")[4 - d.jms, :-1, 1 - d.ims:1 - d.ims + L].astype(np.float64)
    dnw = o.get("dnw")[:-1].astype(np.float64)
    col = (t2 * dnw[:, None]).sum(axis=0)                                 # sum_k tend dnw / g = H3(top) - H3(surface) = 0
    assert np.abs(col).max() < 1e-5 * np.abs(t2 * dnw[:, None]).sum(axis=0).max()
    assert np.abs(t2).max() > 0.1                                         # rho K dq/dz is not uniform: rho falls with height
    o = Oracle(case)
    o.start_domain()
    n = 20
    o.step(n)
    c = case.config
    K = np.float64(c.khdif) * 3.0
    lam = 1.0 - c.dt * K * (2.0 - 2.0 * np.cos(2 * np.pi / L)) / c.dx ** 2
    got = o.get("scalar_1")[4 - d.jms, 3, 1 - d.ims:1 - d.ims + L].astype(np.float64)
    amp = 2.0 * np.abs(np.fft.rfft(got - got.mean())[1]) / L
    assert abs(amp / 0.5 - lam ** n) < 1.5e-3 * lam ** n, (amp / 0.5, lam ** n)
    assert abs(got.mean() - 1.0) < 1e-5                               # diffusion conserves the mean
    assert np.abs(o.get("w_2")).max() < 1e-3                          # and the atmosphere stayed at rest


def test_tke_closure_decays_at_rest_and_grows_in_the_supercell_updraft():
    """km_opt = 2 / diff_opt = 2 (tke_km, tke_rhs, bound_tke -- the closure of the shipped em_quarter_ss and em_les namelists).
    (i) In a resting, stably stratified, horizontally uniform atmosphere there is no shear production and the buoyancy term is a
    sink, so an initial tke can only decay: the domain total falls monotonically, tke stays within [0, tke_upper_bound] and the
    air stays at rest.  (ii) With the shipped em_quarter_ss closure settings the 3 K bubble's updraft spins tke up from exactly
    zero (tke_seed lets K start from the floor), and only where there is shear."""
    case = make_case("b_wave", nx=17, ny=12, nz=11, calm=True, diff_opt=2, km_opt=2, mix_full_fields=1)
    d = case.dims
    kk = np.arange(d.kms, d.kme + 1, dtype=np.float32)
    t0 = (0.5 * np.exp(-kk / 4.0))[None, :, None] * np.ones_like(case.fields["t_2"])
    t0[:, -1, :] = 0
    case.fields["scalar_1"] = t0.copy(); case.fields["scalar_old_1"] = t0.copy()
    o = Oracle(case)
    o.start_domain()
    own = (slice(1 - d.jms, d.jde - d.jms), slice(0, d.kde - 1), slice(1 - d.ims, d.ide - d.ims))
    tot = [float(o.get("scalar_1")[own].astype(np.float64).sum())]
    for _ in range(5):
        o.step(2)
        tke = o.get("scalar_1")[own]
        assert tke.min() >= 0.0 and tke.max() <= case.config.tke_upper_bound
        tot.append(float(tke.astype(np.float64).sum()))
    assert all(b < a for a, b in zip(tot, tot[1:])) and tot[-1] < 0.9 * tot[0], tot
    assert np.abs(o.get("w_2")).max() < 1e-3 and np.abs(o.get("u_2") - o.get("u_2")[own[0], :, :][:1, :, own[2]].mean()).max() < 50.0
    case = make_case("quarter_ss", nx=26, ny=24, nz=21, diff_opt=2, km_opt=2, mix_full_fields=1)
    o = Oracle(case)
    o.set_threads(8)
    o.start_domain()
    assert float(np.abs(o.get("scalar_1")).max()) == 0.0
    o.step(50)
    tke, w = o.get("scalar_1"), o.get("w_2")
    assert tke.min() >= 0.0 and 0.01 < tke.max() < 10.0
    j, k, i = np.unravel_index(np.argmax(tke), tke.shape)
    assert np.abs(w[max(j - 3, 0):j + 4, :, max(i - 3, 0):i + 4]).max() > 0.5 * np.abs(w).max()      # tke sits at the updraft
    assert float(o.get("xkmh").max()) > 10.0 * float(o.get("xkmh")[o.get("xkmh") > 0].min())


def test_sixth_order_filter_damps_the_two_delta_x_wave_at_the_documented_rate():
    """sixth_order_diffusion (module_big_step_utilities_em.F:6220): for q = (-1)^i the flux differences are +-32 q, so the
    tendency is -64 * (factor * 0.015625 / (2 dt')) * mu * q = -(factor / (2 dt')) mu q: the filter removes factor / 2 of a
    2 dx wave per direction and time step dt' ("diff_6th_factor = 0.12: 12 % of the 2 dx checkerboard per step").  For the
    scalars rk_scalar_tend passes dt' = dt_rk = dt / 3 of the first substep (module_em.F:1423-1426) while the tendency acts
    over the whole step, so a passive scalar in a resting atmosphere loses 3 * factor / 2 per step; a 4 dx wave,
    by the same algebra (flux differences -+ 8 q ... ), loses an eighth of that; diff_6th_opt = 2 changes neither
    (all fluxes run down the gradient), and scalar_mix6_off switches the filter off."""
    nx, factor = 33, 0.12
    ii = None
    got = {}
    for label, kw, wave in (("2dx", dict(diff_6th_opt=1), 2), ("2dx_opt2", dict(diff_6th_opt=2), 2), ("4dx", dict(diff_6th_opt=1), 4),
                            ("off", dict(diff_6th_opt=1, scalar_mix6_off=1), 2)):
        case = make_case("b_wave", nx=nx, ny=12, nz=11, calm=True, num_scalar=1, scalar_adv_opt=1, diff_6th_factor=factor, **kw)
        d = case.dims
        ii = np.arange(d.ims, d.ime + 1, dtype=np.float64)
        prof = (1.0 + 0.5 * np.cos(2 * np.pi * (ii - 1) / wave)).astype(np.float32)
        s = case.fields["scalar_1"]
        s[:] = prof[None, None, :]
        s[:, -1, :] = 0
        o = Oracle(case)
        o.start_domain()
        o.step(3)
        row = o.get("scalar_1")[4 - d.jms, 3, 1 - d.ims:1 - d.ims + 32].astype(np.float64)
        amp = 2.0 * np.abs(np.fft.rfft(row - row.mean())[32 // wave]) / 32 / (2.0 if wave == 2 else 1.0)
        got[label] = amp / 0.5
        assert abs(row.mean() - 1.0) < 1e-5
    assert abs(got["2dx"] - (1.0 - 1.5 * factor) ** 3) < 1e-4, got
    assert abs(got["2dx_opt2"] - got["2dx"]) < 1e-6, got
    assert abs(got["4dx"] - (1.0 - 1.5 * factor / 8.0) ** 3) < 1e-4, got
    assert abs(got["off"] - 1.0) < 1e-6, got


def test_even_order_advection_is_the_odd_order_without_its_dissipation():
    """h_sca_adv_order = 6 uses flux6 where 5 uses flux5 = flux6 - sign(u)/60 * (6th difference) (module_advect_em.F:202-213,
    3394-3400): the extra term is purely dissipative.  A 10 dx sine wave carried by the ADVECT_KERNEL harness's uniform
    flow (Courant number C = 0.45 per loop) keeps the same phase under both orders, and the 5th-order amplitude is the
    6th-order one times exp(-N C (2 sin(k dx / 2))^6 / 60), the damping of that term alone."""
    res = {}
    nloops, wl = 40, 10.0
    for h in (5, 6):
        case = parity.tophat_case(0)
        case.config.h_sca_adv_order = h
        d = case.dims
        ii = np.arange(d.ims, d.ime + 1, dtype=np.float64)
        q = (1.0 + 0.5 * np.sin(2 * np.pi * (ii - 1) / wl)).astype(np.float32)
        for n in ("scalar_1", "scalar_old_1"):
            case.fields[n][:] = q[None, None, :]
        o = Oracle(case)
        parity.feeder_loop(o, nloops)
        row = o.get("scalar_1")[1 - d.jms, 3, 1 - d.ims:1 - d.ims + 90].astype(np.float64)
        f = np.fft.rfft(row - row.mean())[9]
        res[h] = (2 * np.abs(f) / 90 / 0.5, np.angle(f), row.mean())
    courant = 90.0 * 5.0 / 1000.0
    want = np.exp(-nloops * courant * (2 * np.sin(np.pi / wl)) ** 6 / 60.0)
    assert abs(res[5][0] / res[6][0] - want) < 2e-4, (res, want)
    assert abs(res[5][1] - res[6][1]) < 1e-3                    # same phase speed
    assert res[6][0] > 0.985                                    # what is left is RK3's own damping
    assert abs(res[5][2] - 1.0) < 1e-6 and abs(res[6][2] - 1.0) < 1e-6


def test_monotonic_advection_creates_no_new_extrema():
    """advect_scalar_mono (module_advect_em.F:9495-10560, "monotonic advection option ... Smolarkiewicz MWR 1989"):
    a top-hat tracer advected through the sheared supercell environment stays inside its initial range [0, 1]
    up to rounding, where the positive-definite option only keeps it non-negative, and both conserve the tracer mass."""
    got = {}
    for opt in (1, 2):
        case = make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=1, scalar_adv_opt=opt)
        o = Oracle(case)
        o.start_domain()
        m0 = (interior(o, "scalar_1").astype(np.float64) * _mass(o)).sum()
        o.step(10)
        s = interior(o, "scalar_1")
        m1 = (s.astype(np.float64) * _mass(o)).sum()
        got[opt] = (float(s.min()), float(s.max()), abs(m1 - m0) / m0)
    assert got[2][0] > -1e-6 and got[2][1] < 1.0 + 1e-5, got       # monotonic: no under- or overshoot
    assert got[1][0] > -1e-6 and got[1][1] > 1.05, got             # positive definite: overshoots remain
    assert got[1][2] < 1e-4 and got[2][2] < 1e-4, got              # both conserve mass (real*4 sums over 10 steps)


def test_weno_advection_is_essentially_non_oscillatory():
    """advect_scalar_weno / advect_scalar_wenopd (module_advect_em.F:7889-9490, "5th-order WENO ... Jiang and Shu 1996"):
    the same top-hat tracer in the same flow.  The unlimited 5th-order scheme rings by 19 % below and 33 % above the
    initial range; WENO keeps the excursions to a few per cent on both sides (that is what "essentially
    non-oscillatory" buys, without a limiter), WENO-PD additionally keeps the tracer non-negative; all conserve mass."""
    got = {}
    for opt in (0, 3, 4):
        case = make_case("synthetic", nx=34, ny=30, nz=17, bubble_lattice=20000.0, noise=0.05, num_scalar=1, scalar_adv_opt=opt)
        o = Oracle(case)
        o.start_domain()
        m0 = (interior(o, "scalar_1").astype(np.float64) * _mass(o)).sum()
        o.step(10)
        s = interior(o, "scalar_1")
        m1 = (s.astype(np.float64) * _mass(o)).sum()
        got[opt] = (float(s.min()), float(s.max()), abs(m1 - m0) / m0)
    assert got[0][0] < -0.15 and got[0][1] > 1.25, got                      # Gibbs ringing of the linear scheme
    assert -0.04 < got[3][0] < 0.0 and 1.0 < got[3][1] < 1.02, got          # WENO: an order of magnitude less
    assert got[4][0] > -1e-6 and got[4][1] < 1.02, got                      # WENO-PD: and non-negative
    assert max(g[2] for g in got.values()) < 2e-4, got                      # flux form: mass conserved (real*4 sums)


def test_weno_switches_theta_advection_on_the_last_substep_only():
    """module_em.F:603-621: with a WENO option on either scalar family theta is advected by advect_scalar_weno on the
    last RK substep.  Where the data are smooth WENO's non-linear weights approach the linear ones, so one step of the
    supercell case (a 3 K bubble six cells across) moves theta by what the linear scheme does to within a tenth."""
    a = Oracle(make_case("quarter_ss", nx=26, ny=24, nz=21, moist_adv_opt=1))
    b = Oracle(make_case("quarter_ss", nx=26, ny=24, nz=21, moist_adv_opt=4))
    t0 = None
    for o in (a, b):
        o.start_domain()
        if t0 is None:
            t0 = interior(o, "t_2").astype(np.float64).copy()
        o.step(1)
    ta, tb = interior(a, "t_2").astype(np.float64), interior(b, "t_2").astype(np.float64)
    assert not np.array_equal(ta, tb)                                       # a different operator ran
    assert np.abs(ta - tb).max() < 0.2 * np.abs(ta - t0).max()              # 3.4e-3 K against a 3.5e-2 K step
    assert np.sqrt(((ta - tb) ** 2).mean()) < 0.1 * np.sqrt(((ta - t0) ** 2).mean())


def test_ieva_splits_the_mass_flux_and_conserves_tracer_mass():
    """module_ieva_em.F: WW_SPLIT hands the part of ww above the adjusted Courant limit to the implicit solver:
    wwE + wwI = ww, both carry the sign of ww, the explicit part's Courant number stays below alpha_max = 1.1
    (:7-8, :262-281), and wwI is non-zero in the updraft core.  The implicit upwind scheme is in flux form: a tracer
    advected through the core has the mass budget of the explicit run (and, with the positive-definite limiter on the
    explicit part, keeps its sign),
    and the solution stays close to the explicit one where both are stable."""
    got = {}
    for opt in (0, 1):
        case = parity.with_updraft(make_case("synthetic", nx=34, ny=30, nz=33, bubble_lattice=20000.0, noise=0.05, num_scalar=1,
                                             scalar_adv_opt=1, zadvect_implicit=opt, dt=20.0, time_step=20), 60.0)
        o = Oracle(case)
        o.start_domain()
        m0 = (interior(o, "scalar_1").astype(np.float64) * _mass(o)).sum()
        o.step(4)
        s = interior(o, "scalar_1")
        m1 = (s.astype(np.float64) * _mass(o)).sum()
        got[opt] = (abs(m1 - m0) / m0, float(s.min()), interior(o, "w_2").astype(np.float64), interior(o, "t_2").astype(np.float64))
        if opt == 1:
            ww, wwE, wwI = (interior(o, n).astype(np.float64) for n in ("ww_m", "wwE", "wwI"))
            assert np.abs(wwI).max() > 1.0                                   # the implicit part is active ...
            assert np.abs(wwE + wwI - ww).max() <= 1e-6 * np.abs(ww).max()   # ... a split, not a change, of ww
            assert (wwE * ww >= 0).all() and (wwI * ww >= 0).all()
            d, c = case.dims, case.config
            mass = o.get("c1f")[None, :d.kde - 1, None] * interior(o, "muts").astype(np.float64)[:, None, :] + o.get("c2f")[None, :d.kde - 1, None]
            cr = np.abs(wwE) * c.dt * np.abs(o.get("rdnw").astype(np.float64))[None, :d.kde - 1, None] / mass
            assert cr.max() <= 1.1 + 1e-5, cr.max()
    assert abs(got[1][0] - got[0][0]) < 2e-4, (got[0][0], got[1][0])         # the implicit part moves no tracer mass of its own
    assert got[1][1] > -1e-6                                                 # positive definite with the implicit part
    assert np.isfinite(got[1][2]).all()
    assert np.abs(got[1][2] - got[0][2]).max() < 0.25 * np.abs(got[0][2]).max()   # same flow as the explicit scheme
    assert np.abs(got[1][3] - got[0][3]).max() < 1.0


def _mass(o):
    """(c1h*mu + c2h) per mass point of the interior, shape (j, k, i)"""
    d = o.case.dims
    mu = (o.get("mu_2") + o.get("mub"))[1 - d.jms:d.jde - d.jms, 1 - d.ims:d.ide - d.ims].astype(np.float64)
    c1h, c2h = o.get("c1h").astype(np.float64), o.get("c2h").astype(np.float64)
    nk = d.kde - 1
    return c1h[None, :nk, None] * mu[:, None, :] + c2h[None, :nk, None]


def test_quarter_ss_supercell_with_documented_default_mixing():
    """test/em_quarter_ss/README.quarter_ss: "This test case produces a simulation of a supercell thunderstorm ...
    The default version of this test case uses a constant eddy viscosity for turbulent mixing -> diff_opt=1 and
    km_opt = 1" (Weisman-Klemp 1982 sounding, quarter-circle shear, 3 K bubble, Kessler, khdif = kvdif = 500 as
    shipped).  On the shipped 42 x 42 x 41 grid the oracle must grow the storm the literature shows for this
    sounding: a deep updraft of 15-40 m/s by 40 min, rain water of several g/kg aloft and rain at the ground,
    with everything finite.  (With the gather and constant-K kernels of the device path held bit for bit to the
    oracle, the same holds for the GPU.)"""
    import os
    case = make_case("quarter_ss", diff_opt=1, km_opt=1, khdif=500.0, kvdif=500.0)
    o = Oracle(case)
    o.set_threads(os.cpu_count() or 1)
    o.start_domain()
    o.step(200)                                     # 40 min at dt = 12 s
    w, qr, rain = interior(o, "w_2"), interior(o, "moist_qr"), interior(o, "rainnc")
    for n in ["u_2", "v_2", "w_2", "t_2", "ph_2", "mu_2", "moist_qv", "moist_qc", "moist_qr"]:
        assert np.isfinite(o.get(n)).all(), n
    assert 15.0 < w.max() < 40.0, w.max()
    assert -15.0 < w.min() < -2.0, w.min()          # compensating downdraughts
    assert 3.0e-3 < qr.max() < 2.0e-2, qr.max()     # several g/kg of rain water
    assert rain.max() > 1.0                         # mm at the ground under the core
    assert interior(o, "moist_qv").min() >= 0.0 and interior(o, "moist_qc").min() >= 0.0 and qr.min() >= 0.0   # PD advection + Kessler


def test_w_damp_starts_at_w_crit_cfl_with_ieva():
    """module_big_step_utilities_em.F:2602-2608: with zadvect_implicit > 0 the damping of w starts at w_crit_cfl, without it
    at w_beta = 1; between the two thresholds an IEVA run must leave rw_tend alone (the term -sign(w)*0.3*(cfl - w_crit_cfl)*mass
    would be positive there: it would accelerate w).  The case has Courant numbers in (1, 2) and above 2."""
    from tests.parity import small_cases
    base = small_cases()["ieva_wdamp"]
    tend = {}
    for damping in (0, 1):
        case = small_cases()["ieva_wdamp"]
        case.config.w_damping = damping
        o = Oracle(case)
        o.start_domain()
        for ph, args in (("rk_step_prep", (1,)), ("rk_phys_bc_dry_1", ()), ("init_zero_tendency", ()), ("rk_tendency", (1,))):
            o.call(ph, *args)
        tend[damping] = o.get("rw_tend").copy()
        if damping:
            ww, mut = o.get("ww"), o.get("mut")
            c1f, c2f, rdnw = o.get("c1f"), o.get("c2f"), o.get("rdnw")
            d = case.dims
            with np.errstate(invalid="ignore", divide="ignore"):     # the memory halo holds zeros
                cfl = np.abs(ww / (c1f[None, :, None] * mut[:, None, :] + c2f[None, :, None]) * rdnw[None, :, None] * np.float32(case.config.dt))
            sl = (slice(d.jps - d.jms, d.jpe - d.jms), slice(1, d.kde - 1), slice(d.ips - d.ims, d.ipe - d.ims))
            cfl = cfl[sl]
            changed = (tend[1] != tend[0])[sl]
            between = (cfl > 1.0) & (cfl <= 2.0)
            assert between.sum() > 20 and (cfl > 2.0).sum() > 20
            assert not changed[between].any()           # untouched below w_crit_cfl
            assert changed[cfl > 2.0].all()             # damped above it
            mv, mh = o.max_cfl()
            assert abs(mv - cfl.max()) <= 1e-6 * mv and 0 < mh < 1
    assert base.config.zadvect_implicit == 1 and base.config.w_crit_cfl == 2.0
