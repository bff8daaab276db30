"""Run R ranks of the CUDA product in one process (one Python thread and one
GPU per rank, NCCL between them) and assemble global arrays -- the harness of
the multi-GPU parity tests.  ctypes releases the GIL during library calls, so
the blocking NCCL rendezvous inside moloch_b200_comm_init works."""
from __future__ import annotations

import threading

import numpy as np

from regcm_b200.moloch import MolochB200

import util


class MultiRank:
    def __init__(self, wl, px, py, fields, profiles, devices=None, transport="p2p", bdy=None, boundary=None, xbctime=None):
        self.wl, self.n = wl, px * py
        devices = devices or list(range(self.n))
        # transport: "p2p" (peer stores), "nccl", or "p2p+nccl" (peer-store halos, NCCL for the row/column
        # reductions of the spectral nudging)
        uid = MolochB200.comm_id(util.LIB) if (self.n > 1 and "nccl" in transport) else None
        self.ranks = [None] * self.n
        blobs = [None] * self.n
        bar = threading.Barrier(self.n)
        errs = []

        def boot(r):
            try:
                m = MolochB200(wl, rank=r, nranks=self.n, px=px, py=py, device=devices[r], bdy=bdy,
                               lib=util.LIB).allocate_moloch()
                self.ranks[r] = m
                if self.n > 1 and "p2p" in transport:
                    blobs[r] = m.p2p_export()
                bar.wait(timeout=120)
                if self.n > 1:
                    if "nccl" in transport:
                        m.comm_init(uid)
                    if "p2p" in transport:
                        m.p2p_connect(blobs)
                bar.wait(timeout=120)
                m.init_moloch(fields, profiles)
                if boundary is not None:
                    m.load_boundary(boundary)
                    if xbctime is not None:
                        m.set_xbctime(xbctime)
                if wl.do_slice:
                    m.set_calday(wl.calday, wl.dayspy)
            except Exception as e:  # noqa: BLE001
                errs.append((r, e))
                bar.abort()
        self._par(boot)
        if errs:
            raise RuntimeError(f"rank boot failed: {errs}")

    def _par(self, fn):
        ts = [threading.Thread(target=fn, args=(r,)) for r in range(self.n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    def call(self, method, *args):
        errs = []

        def run(r):
            try:
                getattr(self.ranks[r], method)(*args)
                self.ranks[r].sync()
            except Exception as e:  # noqa: BLE001
                errs.append((r, e))
        self._par(run)
        if errs:
            raise RuntimeError(f"{method} failed: {errs}")

    def get_global(self, name):
        out = np.zeros(self.ranks[0].global_shape(name))
        for m in self.ranks:
            m.get_into_global(name, out)
        return out

    def close(self):
        for m in self.ranks:
            if m is not None:
                m.close()
