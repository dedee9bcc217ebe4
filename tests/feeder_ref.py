"""PROGRAM feeder of dyn_em/module_advect_em.F:10567-10870 (the reference's own kernel harness, built with -DADVECT_KERNEL), driven here
with the reference's advection routines EXECUTED FROM SOURCE (tests/fortran_subset.py).

The program body is ALLOCATEs, whole-array assignments and PRINTs around the kernel calls; those few statements are restated below in
numpy float32 (each line cites the statement it stands for), while every `CALL advect_scalar*` is taken from the program text itself
(tests/refrun.call_statement reads the CALL and its actual arguments) and run through the translated routine.  The harness's output is
what its `column` routine writes to NNNNNN.txt: NINT(field(:,1,2,1)*100) for i = its..ite.  Test infrastructure; needs /root/reference."""
import numpy as np

from tests import refrun

f32 = np.float32
SRC = "dyn_em/module_advect_em.F"


class Feeder:
    def __init__(self, adv_opt):
        self.L = refrun.library(5)
        # :10612-10615
        self.dims = dict(ids=1, ide=91, jds=1, jde=3, kds=1, kde=10, ims=-5, ime=96, jms=-5, jme=8, kms=1, kme=10,
                         its=1, ite=91, jts=1, jte=3, kts=1, kte=10)
        d = self.dims
        ni, nk, nj = d["ime"] - d["ims"] + 1, d["kme"] - d["kms"] + 1, d["jme"] - d["jms"] + 1
        z3, z2, z1 = (lambda: np.zeros((ni, nk, nj), f32)), (lambda: np.zeros((ni, nj), f32)), (lambda: np.zeros(nk, f32))
        cfg = refrun._Flags(h_sca_adv_order=5, v_sca_adv_order=3, periodic_x=True, periodic_y=True, scalar_adv_opt=adv_opt)      # init, :38-54
        for k in ("symmetric_xs", "symmetric_xe", "symmetric_ys", "symmetric_ye", "open_xs", "open_xe", "open_ys", "open_ye", "specified", "nested"):
            cfg[k] = False
        v = self.v = dict(d)
        v.update(config_flags=cfg, tenddec=False, im=1)
        field = z3()                                          # tophat, :56-80
        for i in range(d["its"], d["ite"] + 1):
            if 35 < i < 55:
                field[i - d["ims"], :, d["jts"] - d["jms"]:d["jte"] - d["jms"] + 1] = 1.0
        v.update(field=field, field_old=field.copy(), tendency=z3(), h_tendency=z3(), z_tendency=z3())
        v.update(mub=z2() + f32(1), mut=z2() + f32(1), mu_old=z2(), ru=z3() + f32(90), rv=z3(), rom=z3(), romi=z3())            # :10665-10672
        for n in ("msfux", "msfuy", "msfvx", "msfvy", "msftx", "msfty"):
            v[n] = z2() + f32(1)
        v.update(rdx=f32(f32(1) / f32(1000.)), rdy=f32(f32(1) / f32(1000.)))                                                     # :10679-10680
        kts, kte, kde = d["kts"], d["kte"], d["kde"]
        znw, rdzw, dnw, rdnw, dn, rdn, fzp, fzm = (z1() for _ in range(8))
        for k in range(kts, kte + 1):                         # :10681-10683
            znw[k - 1] = f32(1) - f32(f32(k) - f32(kts)) / f32(f32(kte) - f32(kts))
        for k in range(kts, kte):                             # :10684-10686
            rdzw[k - 1] = f32(1.) / f32(znw[k - 1] - znw[k])
        for k in range(1, kde):                               # :10687-10690
            dnw[k - 1] = znw[k] - znw[k - 1]
            rdnw[k - 1] = f32(1.) / dnw[k - 1]
        for k in range(2, kde):                               # :10691-10696
            dn[k - 1] = f32(0.5) * f32(dnw[k - 1] + dnw[k - 2])
            rdn[k - 1] = f32(1.) / dn[k - 1]
            fzp[k - 1] = f32(f32(.5) * dnw[k - 1]) / dn[k - 1]
            fzm[k - 1] = f32(f32(.5) * dnw[k - 2]) / dn[k - 1]
        v.update(znw=znw, rdzw=rdzw, dnw=dnw, rdnw=rdnw, dn=dn, rdn=rdn, fzp=fzp, fzm=fzm, c1=z1() + f32(1), c2=z1())           # :10697-10700
        v.update(time_step=5, dt=f32(5))                      # :10702-10703
        self.adv_opt = adv_opt
        third = {0: ("advect_scalar", 2), 1: ("advect_scalar_pd", 0), 2: ("advect_scalar_mono", 0), 3: ("advect_scalar_weno", 0),
                 4: ("advect_scalar_wenopd", 0)}[adv_opt]
        self.calls = [self._bind("advect_scalar", 0), self._bind("advect_scalar", 1), self._bind(*third)]

    def _bind(self, routine, occurrence):
        """the occurrence-th CALL of the routine in PROGRAM feeder -> (Routine, kwargs)"""
        import os
        from tests.fortran_subset import _cpp, _strip_comment
        R = self.L.routine(routine)
        pos, kw = refrun.call_statement(routine, occurrence, SRC, defines={"ADVECT_KERNEL": 1})
        assert not kw and len(pos) == len(R.dummies)
        args = {}
        for dummy, a in zip(R.dummies, pos):
            a = a.strip().lower().replace(" ", "")
            if a.endswith("(ims,kms,jms,im)") or a.endswith("(ims,kms,jms)"):      # the slab actuals
                a = a[:a.index("(")]
            if a in ("time_step/3", "time_step/2"):                                # integer division
                args[dummy] = self.v["time_step"] // int(a[-1])
            else:
                args[dummy] = self.v[a]
        return R, args

    def loop(self):
        """one trip of the DO loop = 1, 2000 body (:10711-10838)"""
        v = self.v
        for n, (R, args) in enumerate(self.calls):
            v["tendency"][...] = 0
            R(**args)
            t = (v["dt"] * v["tendency"]).astype(f32)
            if n == 0:
                t = (t / f32(3.)).astype(f32)                # field = field_old + dt * tendency / 3.
            elif n == 1:
                t = (t / f32(2.)).astype(f32)
            v["field"][...] = v["field_old"] + t
        d, f = self.dims, v["field"]
        I = lambda i: i - d["ims"]
        for k in range(1, d["kde"] + 1):                      # field(:,k,:,:) = field(:,2,:,:)
            f[:, k - 1, :] = f[:, 1, :]
        J = lambda j: j - d["jms"]
        f[:, :, J(2)] = f[:, :, J(1)]
        f[:, :, J(3)] = f[:, :, J(1)]
        for n in range(5):                                    # field(ite+n,...) = field(ids+n,...)
            f[I(d["ite"] + n)] = f[I(d["ids"] + n)]
        for n in range(5):                                    # field(ids-n,...) = field(ite-n,...)
            f[I(d["ids"] - n)] = f[I(d["ite"] - n)]
        v["field_old"][...] = f

    def column(self):
        """what `column` writes: NINT(field(its:ite,1,2,1) * 100)"""
        d = self.dims
        row = self.v["field"][d["its"] - d["ims"]:d["ite"] - d["ims"] + 1, 0, 2 - d["jms"]]
        return [int(x) for x in np.rint((row * f32(100)).astype(f32))]

    def row(self):
        d = self.dims
        return self.v["field"][d["its"] - d["ims"]:d["ite"] - d["ims"] + 1, 0, 2 - d["jms"]].copy()
