"""Plays RSL_LITE for the multi-rank ORACLE runs (test infrastructure).

The oracle calls back at every HALO_EM_* site of solve_em.F with the exchange's Registry name.
Which fields, which width and which index boxes move is asked from the PRODUCT's host logic
(wrfgpu_halo_spec / wrfgpu_halo_box behind the C-ABI) -- that is the code under test here: if it
is right, N ranks of the oracle reproduce the single-rank oracle bit for bit, the reference's own
BIT4BIT criterion (tools/regtest.csh:14-18).  Two transports: threads of one process, and
torch.distributed (gloo) between processes."""
from __future__ import annotations

import ctypes
import threading

import numpy as np

import wrf_b200

CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_char_p)


def halo_spec(config, name):
    lib = wrf_b200.load()
    w = ctypes.c_int(0)
    buf = ctypes.create_string_buffer(4096)
    if lib.wrfgpu_halo_spec(ctypes.byref(config), name.encode(), ctypes.byref(w), buf, ctypes.c_size_t(4096)) != 0:
        raise KeyError(name)
    return w.value, [f for f in buf.value.decode().split(",") if f]


def regular_width(name, width):
    """the Registry's HALO width towards a neighbour inside the domain (the PERIOD width applies across a periodic edge)"""
    return int(wrf_b200.load().wrfgpu_halo_regular_width(name.encode(), ctypes.c_int(width)))


def halo_box(dims, config, width, direction, send, hwidth=0):
    lib = wrf_b200.load()
    v = [ctypes.c_int(0) for _ in range(5)]
    ok = lib.wrfgpu_halo_box2(ctypes.byref(dims), ctypes.byref(config), width, hwidth, direction, 1 if send else 0, *[ctypes.byref(x) for x in v])
    if not ok:
        return None
    peer, i0, i1, j0, j1 = [x.value for x in v]
    return peer, i0, i1, j0, j1


class ThreadTransport:
    """ranks are threads of this process"""

    def __init__(self, n):
        self.n = n
        self.box = {}
        self.bar = threading.Barrier(n)

    def exchange(self, me, sends, recvs):
        """sends: {dir: (peer, ndarray)}, recvs: {dir: (peer, count)} -> {dir: ndarray}"""
        for d, (peer, buf) in sends.items():
            self.box[(me, peer, d)] = buf
        self.bar.wait()
        out = {d: self.box[(peer, me, d ^ 1)] for d, (peer, _) in recvs.items()}
        self.bar.wait()
        for d, (peer, _) in sends.items():
            self.box.pop((me, peer, d), None)
        return out


class GlooTransport:
    """ranks are processes of a torch.distributed (gloo) group"""

    def exchange(self, me, sends, recvs):
        import torch
        import torch.distributed as dist
        reqs, out, keep = [], {}, []
        for d, (peer, n) in recvs.items():
            t = torch.empty(n, dtype=torch.float32)
            out[d] = t
            reqs.append(dist.irecv(t, src=peer, tag=d ^ 1))
        for d, (peer, buf) in sends.items():
            t = torch.from_numpy(np.ascontiguousarray(buf))
            keep.append(t)
            reqs.append(dist.isend(t, dst=peer, tag=d))
        for r in reqs:
            r.wait()
        return {d: t.numpy() for d, t in out.items()}


class HaloEmulator:
    def __init__(self, oracle, transport):
        self.o, self.t = oracle, transport
        self.case = oracle.case
        self.me = self.case.config.mytask
        self.calls = []
        self._cb = CB(self._on_halo)
        oracle.lib.oracle_field_ptr.restype = ctypes.POINTER(ctypes.c_float)
        oracle.lib.oracle_set_halo_callback(oracle.h, self._cb, None)

    def view(self, name):
        d = self.case.dims
        ni, nk, nj = d.ime - d.ims + 1, d.kme - d.kms + 1, d.jme - d.jms + 1
        n = int(self.o.lib.oracle_field_size(self.o.h, name.encode()))
        if n == 0:
            return None
        p = self.o.lib.oracle_field_ptr(self.o.h, name.encode())
        a = np.ctypeslib.as_array(p, shape=(n,))
        return a.reshape(nj, nk, ni) if n == ni * nk * nj else a.reshape(nj, ni)

    def _on_halo(self, _ctx, cname):
        name = cname.decode()
        self.calls.append(name)
        d, c = self.case.dims, self.case.config
        if c.ntasks_x * c.ntasks_y == 1:
            return
        width, fields = halo_spec(c, name)
        hw = regular_width(name, width)
        views = [v for v in (self.view(f) for f in fields) if v is not None]
        for d0 in (0, 2):                              # Y phase, then X phase
            sends, recvs, rboxes = {}, {}, {}
            for q in (d0, d0 + 1):
                sb = halo_box(d, c, width, q, True, hw)
                if sb is None:
                    continue
                rb = halo_box(d, c, width, q, False, hw)
                peer, i0, i1, j0, j1 = sb
                sj, si = slice(j0 - d.jms, j1 - d.jms + 1), slice(i0 - d.ims, i1 - d.ims + 1)
                sends[q] = (peer, np.concatenate([(v[sj, :, si] if v.ndim == 3 else v[sj, si]).ravel() for v in views]))
                peer, i0, i1, j0, j1 = rb
                rj, ri = slice(j0 - d.jms, j1 - d.jms + 1), slice(i0 - d.ims, i1 - d.ims + 1)
                cnt = sum((v[rj, :, ri] if v.ndim == 3 else v[rj, ri]).size for v in views)
                recvs[q] = (peer, cnt)
                rboxes[q] = (rj, ri)
            got = self.t.exchange(self.me, sends, recvs)
            for q, buf in got.items():
                rj, ri = rboxes[q]
                pos = 0
                for v in views:
                    tgt = v[rj, :, ri] if v.ndim == 3 else v[rj, ri]
                    assert buf.size >= pos + tgt.size, (name, q, buf.size, pos, tgt.size)
                    tgt[...] = buf[pos:pos + tgt.size].reshape(tgt.shape)
                    pos += tgt.size
                assert pos == buf.size, (name, q, pos, buf.size)
            # a periodic dimension living on this rank is closed on the lines just received
            # (k_halo_close in csrc/halo.cu; cf. the unclipped extents of period.c)
            isy = d0 == 0
            if (isy and c.periodic_x and c.ntasks_x == 1) or ((not isy) and c.periodic_y and c.ntasks_y == 1):
                for q in got:
                    rj, ri = rboxes[q]
                    for v in views:
                        if isy:
                            rows = v[rj]
                            o0, de, ds = -d.ims, d.ide - d.ims, d.ids - d.ims
                            for n in range(1, 5):
                                rows[..., ds - n] = rows[..., de - n]
                            for n in range(0, 5):
                                rows[..., de + n] = rows[..., ds + n]
                        else:
                            cols = v[..., ri]
                            de, ds = d.jde - d.jms, d.jds - d.jms
                            for n in range(1, 5):
                                cols[ds - n] = cols[de - n]
                            for n in range(0, 5):
                                cols[de + n] = cols[ds + n]


def run_ranks_in_threads(cases, nsteps, oracle_cls):
    """one oracle per patch Case, each in its own thread; returns the oracles after nsteps"""
    n = len(cases)
    tr = ThreadTransport(n)
    oracles = [oracle_cls(c) for c in cases]
    emus = [HaloEmulator(o, tr) for o in oracles]
    errs = []

    def work(o):
        try:
            o.start_domain()
            o.step(nsteps)
        except BaseException as e:   # noqa: BLE001
            errs.append(e)
            tr.bar.abort()

    th = [threading.Thread(target=work, args=(o,)) for o in oracles]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    return oracles, emus
