#!/usr/bin/env python
"""Per-kernel device timing without torch (ctypes + the library's own CUDA events):
a quick A/B and roofline probe that starts in seconds on a fresh GPU box.

    python scripts/kbench.py [--workload cordex25] [--steps 5] [--boundary] [--slice] [--spectral] [--tke]

Prints one JSON line: ms/step (host clock around a synchronised batch of steps)
and, per kernel class, launches/step, average ms and GB/s of algorithmic bytes
where a figure is defined (bench.py's table for the dycore kernels; for the
boundary kernels the bytes are counted below from the arrays each one reads
and writes).  Not the judged benchmark: that is bench.py.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from regcm_b200 import synthetic as S  # noqa: E402
from regcm_b200.moloch import MolochB200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cordex25", choices=sorted(S.WORKLOADS))
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--boundary", action="store_true")
    ap.add_argument("--slice", action="store_true")
    ap.add_argument("--spectral", action="store_true")
    ap.add_argument("--tke", action="store_true")
    ap.add_argument("--massck", action="store_true", help="also time massck + the ps guard once per step")
    ap.add_argument("--diag", action="store_true", help="tendency diagnostics (idiag, ichdiag)")
    ap.add_argument("--crop", type=int, default=0, help="crop the horizontal domain to N x N")
    ap.add_argument("--jx", type=int, default=0, help="crop to jx x iy columns (e.g. the per-rank grid of a decomposed run)")
    ap.add_argument("--iy", type=int, default=0)
    ap.add_argument("--digest", action="store_true",
                    help="add the wrap-around sums of the bit patterns of the prognostic fields after the last step "
                         "(two library builds that print the same digests computed the same bits)")
    args = ap.parse_args()
    wl = S.WORKLOADS[args.workload]
    if args.crop:
        wl = S.small(wl, min(wl.jx, args.crop), min(wl.iy, args.crop), wl.kz)
    if args.jx or args.iy:
        wl = S.small(wl, args.jx or wl.jx, args.iy or wl.iy, wl.kz)
    kw = {}
    if args.boundary:
        kw.update(do_bdy=1, present_qc=1, present_qi=1, mo_top_nudge=1, ichebdy=1)
    if args.spectral:
        kw.update(do_bdy=1, mo_spectral_nudge=1, dtrad=wl.dt)       # nudging active every step
    if args.slice:
        kw.update(do_slice=1, icldmstrat=1)
    if args.tke:
        kw.update(ibltyp=2, tkemin=1.0e-4)
    if args.massck:
        kw.update(do_massck=1)
    if args.diag:
        kw.update(idiag=1, ichdiag=1)
    wl = replace(wl, **kw)
    t0 = time.perf_counter()
    m = MolochB200(wl).allocate_moloch()
    fields, profiles, boxes = S.model_inputs_local(wl, m.g)
    if wl.ibltyp == 2:
        zf = S.md_zeta(S.model_zitaf(wl.kz, wl.mo_ztop)[:, None, None], 0.0, wl.mo_ztop, wl.mo_h, wl.mo_a0)
        g = m.g
        fields["tke"] = np.ascontiguousarray(np.broadcast_to(
            wl.tkemin + 0.4 * np.exp(-np.maximum(zf, 0.0) / 800.0),
            (wl.kz + 1, g.ice2 - g.ice1 + 1, g.jce2 - g.jce1 + 1)))
        boxes["tke"] = (g.jce1, g.jce2, g.ice1, g.ice2)
    if wl.do_slice or wl.do_massck:
        g = m.g
        Jg, Ig = np.meshgrid(np.arange(g.jce1, g.jce2 + 1, dtype=np.float64),
                             np.arange(g.ice1, g.ice2 + 1, dtype=np.float64))
        ht = S._height(wl, Jg, Ig) * S.egrav
        fields["zetaf"] = S.md_zeta(S.model_zitaf(wl.kz, wl.mo_ztop)[:, None, None], ht[None], wl.mo_ztop, wl.mo_h,
                                    wl.mo_a0)
        boxes["zetaf"] = (g.jce1, g.jce2, g.ice1, g.ice2)
        if wl.do_slice:
            dl = S.raddeg * wl.dx / S.earthrad
            fields["xlat"] = np.ascontiguousarray(wl.clat - dl * (float(wl.iy) * 0.5 - Ig + 0.5))
            boxes["xlat"] = boxes["zetaf"]
    m.init_moloch(fields, profiles, boxes)
    if wl.do_bdy:
        base = {n: np.zeros(m.global_shape(n)) for n in ("u", "v", "t", "pai", "qx", "ps")}
        for n in base:
            m.get_into_global(n, base[n])
        m.load_boundary(S.make_boundary(wl, base))
    del fields
    t_init = time.perf_counter() - t0
    m.moloch(args.warmup)
    m.sync()
    t0 = time.perf_counter()
    m.moloch(args.steps)
    m.sync()
    ms_step = (time.perf_counter() - t0) / args.steps * 1e3
    m.profile_enable(True)
    for _ in range(args.steps):
        m.moloch(1)
        if args.massck:
            m.massck()
            m.ps_check()
    prof = m.profile_read()
    m.profile_enable(False)
    g = m.g
    cells = (g.jce2 - g.jce1 + 1) * (g.ice2 - g.ice1 + 1) * wl.kz
    nsp = wl.nqx + wl.ntr
    # 3-D FP64 arrays read + written per cell, whole-grid kernels only
    # bdy_relax: only the sponge ring and the top layers do work; the figure is the upper bound
    alg = {"bdy_finish": 2 + 2 + 1 + (5 if wl.ipptls > 1 else 2) + 1 + 2,            # u,v -> ux,vx; t,qx,pai -> tvirt,tetav
           "mkslice": 2 + 1 + 2 + 2 * nsp + 2 + 2 + 2,                                # pai,p,t | qx,trac rw | qsat,rho,w | 4 outputs
           "status_update": 9 + 3 * nsp + 7 + 7}
    out = {"workload": wl.name, "grid": [wl.jx, wl.iy, wl.kz], "F": wl.nfields, "ms_per_step": ms_step,
           "cell_updates_per_s": wl.cells / (ms_step * 1e-3), "init_s": t_init, "steps": args.steps,
           "finite": bool(np.isfinite(m.get_local("pai")).all()), "kernels": {}}
    for name, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        avg = r["ms"] / max(r["launches"], 1)
        e = {"launches_per_step": r["launches"] / args.steps, "avg_ms": round(avg, 5),
             "ms_per_step": round(r["ms"] / args.steps, 5)}
        if name in alg and avg > 0:
            e["alg_gbs"] = round(alg[name] * 8 * cells / (avg * 1e-3) / 1e9, 1)
        out["kernels"][name] = e
    if args.digest:
        out["digest"] = {}
        for n in ("u", "v", "w", "pai", "tetav", "t", "qx", "trac", "ux", "vx"):
            if n == "trac" and wl.ntr == 0:
                continue
            a = np.ascontiguousarray(m.get_local(n), dtype=np.float64)
            out["digest"][n] = int(a.view(np.uint64).sum(dtype=np.uint64))
    print(json.dumps(out), flush=True)
    m.close()


if __name__ == "__main__":
    main()
