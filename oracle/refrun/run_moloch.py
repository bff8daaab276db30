"""Execute the reference's own MOLOCH time-step source (TEST INFRASTRUCTURE).

`ReferenceRun` reads the routines of the hot path out of the reference tree
where they lie --

    Main/mod_moloch.F90     moloch, reset_tendencies, dynamical_core, sound, divergence_damping,
                            divergence_diffusion, advection, wafone, local_flow_param, zstagtoh, htozstag,
                            uvstagtouvx, uvxtouvstag, tvirt_to_temp, temp_to_tvirt,
                            extrapolate_surface_pressure, status_update, boundary
    Main/mod_bdycod.F90     bdyval, morelax_external, morelax_fraction, motopnudge, mospectral_nudge,
                            lowpass_filter
    Main/chemlib/mod_che_bdyco.F90   chem_bdyval_uncoupled, morelax_chiten
    Main/mod_slice.F90      mkslice
    Share/pfwsat.inc        pfwsat
    Share/mod_constants.F90, Main/mpplib/mod_runparams.F90   parameters

-- translates them mechanically (fortran_subset.py) and runs them on the state
of one MPI rank that owns the whole domain (nproc = 1).  What is NOT reference
code here: the allocation of the arrays (bounds restated from
Main/mod_atm_interface.F90:579-624 and Main/mod_moloch.F90:159-199), the
single-rank halo exchange (a periodic copy or nothing, which is what
exchange_lr/_bt/_lrbt reduce to for one rank), the namelist scalars, and stubs
for the driver's bookkeeping (timer, zenith angle, reports).  The static
fields and the initial state are INPUTS, taken from the same arrays the oracle
and the CUDA library are initialised with.

    python -m oracle.refrun.run_moloch            # regenerate tests/golden/reference_*.json

Only tests/ may import this module; it needs /root/reference at run time.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import numpy as np

from . import fortran_subset as F
from .runtime import INTRINSICS, FArr

REF = os.environ.get("REGCM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

SOURCES = {
    "Main/mod_moloch.F90": ["moloch", "reset_tendencies", "dynamical_core", "sound", "divergence_damping",
                            "divergence_diffusion", "advection", "wafone", "local_flow_param", "zstagtoh", "htozstag",
                            "uvstagtouvx", "uvxtouvstag", "tvirt_to_temp", "temp_to_tvirt",
                            "extrapolate_surface_pressure", "status_update", "boundary"],
    "Main/mod_bdycod.F90": ["bdyval", "morelax_external", "morelax_fraction", "motopnudge", "mospectral_nudge",
                            "lowpass_filter"],
    "Main/chemlib/mod_che_bdyco.F90": ["chem_bdyval_uncoupled", "morelax_chiten"],
    "Main/mod_slice.F90": ["mkslice"],
}


def available() -> bool:
    return os.path.exists(os.path.join(REF, "Main", "mod_moloch.F90"))


class _Obj(types.SimpleNamespace):
    pass


class ReferenceRun:
    """One rank owning the whole domain, state held in Fortran-bounded arrays."""

    def __init__(self, wl, oracle, boundary: dict | None = None, dump_dir: str | None = None, px: int = 1,
                 py: int = 1, rank: int = 0, comm=None):
        """px, py, rank, comm: one rank of a px x py run (MultiRankReference); the halo exchange is then the
        reference's own exchange routines on top of an emulated mpi_neighbor_alltoallv."""
        from regcm_b200 import hostmodel as H
        from regcm_b200.decomp import make_geom
        self.wl, self.H = wl, H
        g = self.g = make_geom(wl.jx, wl.iy, wl.kz, wl.i_band, wl.i_crm, px, py, rank)
        ns = self.ns = dict(INTRINSICS)
        # ---- parameters straight from the reference's modules --------------------------------
        ex = F.Expr(set())
        for rel in ("Share/mod_constants.F90", "Main/mpplib/mod_runparams.F90", "Main/mod_moloch.F90",
                    "Main/mod_slice.F90"):
            st = F.preprocess(open(os.path.join(REF, rel)).read())
            for line in F.module_parameters(st, ex):
                try:
                    exec(F.compile_source(line, rel), ns)
                except Exception:       # a parameter that needs something outside the subset: not used by the path
                    pass
        # ---- index ranges (setup_model_indexes) and namelist scalars -------------------------------
        kz = wl.kz
        for n in ("jde1", "jde2", "ide1", "ide2", "jdi1", "jdi2", "idi1", "idi2",
                  "jce1", "jce2", "ice1", "ice2", "jci1", "jci2", "ici1", "ici2"):
            ns[n] = getattr(g, n)
        # Main/mod_atm_interface.F90:300-330
        ns.update(jdii1=g.jde1 + (2 if g.bl else 0), jdii2=g.jde2 - (2 if g.br else 0),
                  idii1=g.ide1 + (2 if g.bb else 0), idii2=g.ide2 - (2 if g.bt else 0))
        gl, gr, gb, gt = g.gl, g.gr, g.gb, g.gt
        # the ga/gb/gc ranges: widened by 1/2/3 points where a neighbour exists (Main/mod_atm_interface.F90:332-382)
        for base in ("jce", "jci", "jde", "jdi"):
            for w_, suf in ((1, "ga"), (2, "gb"), (3, "gc")):
                ns[f"{base}1{suf}"] = ns[base + "1"] - w_ * gl
                ns[f"{base}2{suf}"] = ns[base + "2"] + w_ * gr
        for base in ("ice", "ici", "ide", "idi"):
            for w_, suf in ((1, "ga"), (2, "gb"), (3, "gc")):
                ns[f"{base}1{suf}"] = ns[base + "1"] - w_ * gb
                ns[f"{base}2{suf}"] = ns[base + "2"] + w_ * gt
        crm = wl.i_crm == 1
        band = wl.i_band == 1 or crm
        jcross2, icross2 = (wl.jx if band else wl.jx - 1), (wl.iy if crm else wl.iy - 1)
        ns.update(jmin=1, jmax=jcross2, imin=1, imax=icross2)                 # Main/mod_moloch.F90:280-293
        if band:
            ns.update(jmin=-1, jmax=jcross2 + 2)
        if crm:
            ns.update(imin=-1, imax=icross2 + 2)
        ns.update(kz=kz, kzp1=kz + 1, kzm1=kz - 1, jx=wl.jx, iy=wl.iy, nqx=wl.nqx, ntr=wl.ntr, iqfrst=2,
                  ipptls=wl.ipptls, ibltyp=wl.ibltyp, ichem=int(wl.ntr > 0), ichebdy=wl.ichebdy, idiag=wl.idiag,
                  ichdiag=wl.ichdiag,
                  idynamic=3, mo_nadv=wl.mo_nadv, mo_nsound=wl.mo_nsound, dtsec=wl.dt, dt=wl.dt, rdt=1.0 / wl.dt,
                  dx=wl.dx, rdx=1.0 / wl.dx, lrotllr=bool(wl.lrotllr), do_divdamp=bool(wl.mo_divdamp),
                  do_divfilter=bool(wl.mo_divfilter), do_apply_bdy=bool(wl.do_bdy), moloch_realcase=False,
                  mo_top_nudge=bool(wl.mo_top_nudge), mo_spectral_nudge=bool(wl.mo_spectral_nudge),
                  present_qc=bool(wl.present_qc), present_qi=bool(wl.present_qi), dtbdys=wl.dtbdys, dtrad=wl.dtrad,
                  rtb=1.0 / wl.dtbdys, xbctime=0.0, tkemin=wl.tkemin, rhmin=wl.rhmin, rhmax=wl.rhmax,
                  icldmstrat=wl.icldmstrat, debug_level=0, islab_ocean=0, myid=0, italk=0, nspgx=wl.nspgx,
                  mo_h=wl.mo_h, iconvec=0, total_precip_points=0, xslabtime=0.0, icup=FArr.alloc([(1, 2)], "int"))
        mo_dzita = wl.mo_ztop / float(kz)            # zita(kz), Main/mod_params.F90:2461-2463
        ns.update(mo_dzita=mo_dzita, rdzita=1.0 / mo_dzita)                    # Main/mod_moloch.F90:275
        ns["dtstepa"] = wl.dt / float(wl.mo_nadv)                              # :306
        ns["dtsound"] = ns["dtstepa"] / float(wl.mo_nsound)                    # :307
        ns["ma"] = _Obj(has_bdyleft=g.bl, has_bdyright=g.br, has_bdybottom=g.bb, has_bdytop=g.bt,
                        bandflag=band, crmflag=crm, left=g.left, right=g.right, bottom=g.bottom, top=g.top)
        ns["mpi_proc_null"] = -1
        ns["syncro_rep"] = _Obj(act=lambda: False)
        ns["rcmtimer"] = _Obj(advance=lambda: None, integrating=lambda: True, str=lambda: "")
        ns["zenitm"] = lambda *a: None
        ns["massck"] = lambda: None
        ns["physical_parametrizations"] = lambda: None
        ns["is_present_qc"] = lambda: bool(wl.present_qc)
        ns["is_present_qi"] = lambda: bool(wl.present_qi)
        for n in ("xlat", "xlon", "coszrs"):
            ns[n] = None
        dbx = g.ext("dot", 0, 0)
        ns["mddom"] = _Obj(ldmsk=FArr(np.ones((dbx[3] - dbx[2] + 1, dbx[1] - dbx[0] + 1), dtype=np.int64),
                                      [dbx[0], dbx[2]]))       # land everywhere: the SST update of bdyval is a no-op
        ns.update(lakemod=0, iocncpl=0, iwavcpl=0)
        # ---- arrays: allocate_atmosphere + allocate_moloch bounds, filled from the oracle -----------
        self.arr = {}

        def from_oracle(name, oname=None, spec=None):
            stag, gj, gi, lv = spec or H.ALLOC[name]
            box = g.ext(stag, gj, gi)
            a = np.array(H.cut(oracle.get(oname or name), g, box), dtype=np.float64)
            lb = [box[0], box[2]] + ([1] if a.ndim >= 3 else []) + ([1] if a.ndim == 4 else [])
            if a.ndim == 4:          # species-major (n,k,i,j) is already the reversed (j,i,k,n)
                pass
            self.arr[name] = ns[name] = FArr(a, lb)
            return ns[name]

        for n in ("u", "v", "ux", "vx", "w", "pai", "tetav", "t", "qx", "tvirt", "p", "rho", "fmz", "fmzf", "rfmzu",
                  "rfmzv", "hx", "hy", "coru", "corv", "ps"):
            from_oracle(n)
        for n, bx in (("bdywtu", (g.jdi1, g.jdi2, g.ici1, g.ici2)), ("bdywtv", (g.jci1, g.jci2, g.idi1, g.idi2)),
                      ("bdywtw", (g.jci1, g.jci2, g.ici1, g.ici2))):       # Main/mod_moloch.F90:164-166
            a = np.array(H.cut(oracle.get(n), g, bx), dtype=np.float64)
            self.arr[n] = ns[n] = FArr(a, [bx[0], bx[2], 1])
        from_oracle("z", "zeta", H.ALLOC["zeta"])
        from_oracle("qsat")
        ns["mo_atm"] = _Obj()
        from_oracle("zetaf", "zetaf", ("cross", 0, 0, "kzp1"))
        for n, on in (("mx", "msfx"), ("mu", "msfu"), ("mv", "msfv"), ("mx2", "mx2"), ("rmx", "rmx"), ("rmu", "rmu"),
                      ("rmv", "rmv")):
            from_oracle(n, on, ("dot", 1, 1, 1))
        if wl.ntr > 0:
            from_oracle("trac")
        for n in ("gzitak", "gzitakh", "xkdamp", "xknu", "ffilt"):
            self.arr[n] = ns[n] = FArr(np.array(oracle.get(n), dtype=np.float64), [1])
        rl = np.zeros(wl.iy + 1)
        if wl.lrotllr:
            from regcm_b200 import synthetic as S
            rl = np.asarray(S.make_primary(wl)["rlat"], dtype=np.float64)
        ns["rlat"] = FArr(rl, [1])

        def zeros(name, stag, gj, gi, k1, k2, nspec=0):
            box = g.ext(stag, gj, gi)
            bnds = [(box[0], box[1]), (box[2], box[3])] + ([(k1, k2)] if k2 >= k1 else []) + \
                   ([(1, nspec)] if nspec else [])
            self.arr[name] = ns[name] = FArr.alloc(bnds)
            return ns[name]

        # allocate_moloch (Main/mod_moloch.F90:159-199) and the tendencies (mod_atm_interface.F90:605-618)
        zeros("s", "cross", 0, 0, 1, kz + 1); zeros("wwkw", "cross", 0, 0, 2, kz + 1)
        zeros("tetavf", "cross", 0, 0, 2, kz); zeros("zdiv2", "cross", 1, 1, 1, kz)
        zeros("wz", "cross", 2, 2, 1, kz); zeros("p0", "cross", 2, 2, 1, kz); zeros("wfw", "cross", 0, 0, 1, kz + 1)
        zeros("wx", "cross", 1, 1, 1, kz)
        inner = [(g.jci1, g.jci2), (g.ici1, g.ici2)]

        def interior(name, k2, nspec=0):
            self.arr[name] = ns[name] = FArr.alloc(inner + [(1, k2)] + ([(1, nspec)] if nspec else []))
        interior("laplacian", kz)
        zeros("ud", "u", 0, 0, 1, kz); zeros("vd", "v", 0, 0, 1, kz)
        ns["zpby"] = FArr.alloc([(g.jce1, g.jce2), (g.ici1, g.ice2 + gt), (1, kz)])
        ns["zpbw"] = FArr.alloc([(g.jci1, g.jce2 + gr), (g.ice1, g.ice2), (1, kz)])
        for n in ("tten", "uten", "vten", "cldfra", "cldlwc"):    # cldfra/cldlwc: physics arrays reset_tendencies clears too
            interior(n, kz)
        interior("qxten", kz, wl.nqx)
        if wl.ntr > 0:
            interior("chiten", kz, wl.ntr)
        ns["qv"] = ns["qx"].view_last(1); ns["qvten"] = ns["qxten"].view_last(1)
        if wl.ipptls > 0:
            ns["qc"] = ns["qx"].view_last(2)
            if wl.ipptls > 1:
                ns["qi"], ns["qr"], ns["qs"] = (ns["qx"].view_last(n) for n in (3, 4, 5))
        if wl.ibltyp == 2:
            from_oracle("tke"); interior("tketen", kz + 1); zeros("tkex", "cross", 0, 0, 1, kz)
        ib = [(g.jci1, g.jci2), (g.ici1, g.ici2), (1, kz)]
        if wl.idiag > 0:       # Main/mod_moloch.F90:187-188 and the tdiag/qdiag members used by the path
            ns["ten0"], ns["qen0"] = FArr.alloc(ib), FArr.alloc(ib)
            ns["tdiag"] = _Obj(adh=FArr.alloc(ib), bdy=FArr.alloc(ib))
            ns["qdiag"] = _Obj(adh=FArr.alloc(ib), bdy=FArr.alloc(ib))
        if wl.ichdiag > 0 and wl.ntr > 0:
            for n in ("chiten0", "cadvhdiag", "cbdydiag"):
                ns[n] = FArr.alloc(ib + [(1, wl.ntr)])
        # mo_atm%*, sfs%*, atms%* aliases used by bdyval / mkslice / morelax
        mo = ns["mo_atm"]
        for n in ("u", "v", "w", "t", "pai", "qx", "zetaf"):
            setattr(mo, n, ns[n])
        mo.zeta = ns["z"]
        mo.tke = ns.get("tke")
        ns["sfs"] = _Obj(psa=ns["ps"], psb=ns["ps"])
        ns["chemt"] = ns.get("trac")
        # ---- lateral boundary ---------------------------------------------------------------------------
        if wl.do_bdy:
            B = boundary or {}
            pairs = {"dub": ("dub0", "dub1", "u"), "dvb": ("dvb0", "dvb1", "v"), "xtb": ("xtb0", "xtb1", "t"),
                     "xpaib": ("xpaib0", "xpaib1", "pai"), "xqb": ("xqb0", "xqb1", "qx1"),
                     "xlb": ("xlb0", "xlb1", "qx1"), "xib": ("xib0", "xib1", "qx1"), "xpsb": ("xpsb0", "xpsb1", "ps")}
            for name, (k0, k1, _) in pairs.items():
                stag = "u" if name == "dub" else "v" if name == "dvb" else "cross"
                box = g.ext(stag, 0, 0)
                o = _Obj()
                for key, attr in ((k0, "b0"), (k1, "b1")):
                    a = np.array(H.cut(np.asarray(B[key]), g, box), dtype=np.float64)
                    setattr(o, attr, FArr(a, [box[0], box[2]] + ([1] if a.ndim == 3 else [])))
                ns[name] = o
            if wl.ntr > 0 and wl.ichebdy != 0:
                box = g.ext("cross", 0, 0)
                for key in ("chib0", "chib1"):
                    ns[key] = FArr(np.array(H.cut(np.asarray(B[key]), g, box), dtype=np.float64), [box[0], box[2], 1, 1])
            from regcm_b200 import synthetic as S
            T = S.bdycon_setup(wl, oracle.get("zeta"))
            dbox = g.ext("dot", 0, 0)
            for which, nm in (("cr", "ba_cr"), ("ud", "ba_ud"), ("vd", "ba_vd")):
                ib = np.array(H.cut(T["ibnd"][which], g, dbox), dtype=np.int64)
                ns[nm] = _Obj(ibnd=FArr(ib, [dbox[0], dbox[2]]), havebound=bool((ib > 0).any()))
            cb = ns["ba_cr"].ibnd.a > 0
            ns["cba"] = _Obj(ibnd=ns["ba_cr"].ibnd, havebound=bool(cb.any()), ns=1, nn=0, nw=0, ne=0,
                             bsouth=_Mask(cb, dbox), bnorth=_Mask(cb & False, dbox), bwest=_Mask(cb & False, dbox),
                             beast=_Mask(cb & False, dbox))
            if wl.nspgx > 0:
                ns["hefc"] = FArr(np.array(oracle.get("hefc"), dtype=np.float64).reshape(kz, wl.nspgx), [1, 1])
                ns["fcx"] = FArr(np.asarray(S.chem_fcx(wl), dtype=np.float64), [1])
            if wl.mo_spectral_nudge:
                # lowpass_init's tables (Main/mod_bdycod.F90:3844-3896) are inputs, same bits as the oracle's;
                # the work arrays are allocated as there (:422, :3868-3873)
                km, lm = oracle.get_int("km"), oracle.get_int("lm")
                ns.update(km=km, lm=lm)
                ns["bvx"] = FArr(np.array(oracle.get("bvx")).reshape(2 * km, wl.jx)[:, g.jde1 - 1:g.jde2].copy(), [g.jde1, 1])
                ns["bvy"] = FArr(np.array(oracle.get("bvy")).reshape(2 * lm, wl.iy)[:, g.ide1 - 1:g.ide2].copy(), [g.ide1, 1])
                ns["cnudge"] = FArr(np.array(oracle.get("cnudge"), dtype=np.float64), [1])
                ns["zn1"] = FArr.alloc([(g.jde1, g.jde2), (g.ide1, g.ide2)])
                for n in ("sx", "sxg"):
                    ns[n] = FArr.alloc([(g.ide1, g.ide2), (1, 2 * km)])
                for n in ("sy", "syg"):
                    ns[n] = FArr.alloc([(g.jde1, g.jde2), (1, 2 * lm)])

                def reduce1(m, gl_, a1, a2):
                    # MPI_Allreduce(m, g, nk*(a2-a1+1), SUM) on a communicator of one rank
                    # (Main/mpplib/mod_mppparam.F90:20620-20664): copies the first `count`
                    # elements of the contiguous array
                    count = m.a.shape[0] * (a2 - a1 + 1)
                    gl_.a.reshape(-1)[:count] = m.a.reshape(-1)[:count]
                ns["row_reduce"] = reduce1
                ns["column_reduce"] = reduce1
            ns["tnudge"] = FArr(np.array(oracle.get("tnudge"), dtype=np.float64), [1])
            ns["nztop"] = oracle.get_int("nztop")
            ns["xbctime"] = oracle.get_xbctime()
        # mkslice targets
        ns["atms"] = _Obj(ps2d=ns["ps"], rhob3d=ns["rho"], zq=ns["zetaf"], tb3d=ns["t"], pb3d=ns["p"],
                          qxb3d=ns["qx"], qsb3d=ns["qsat"], chib3d=ns.get("trac"))
        at = ns["atms"]
        at.pf3d = FArr.alloc([(g.jce1, g.jce2), (g.ice1, g.ice2), (1, kz + 1)])
        at.th3d = FArr.alloc([(g.jce1, g.jce2), (g.ice1, g.ice2), (1, kz)])
        for n in ("rhb3d", "wpx3d"):
            setattr(at, n, FArr.alloc([(g.jci1, g.jci2), (g.ici1, g.ici2), (1, kz)]))
        for n in ("rhox2d", "tp2d", "th700"):
            setattr(at, n, FArr.alloc([(g.jci1, g.jci2), (g.ici1, g.ici2)]))
        # the common tail of mkslice (tropopause and PBL-top indices for the physics, Main/mod_slice.F90:342-384)
        at.za = ns["z"]
        from regcm_b200 import synthetic as S_
        xl = np.array(H.cut(S_.make_primary(wl)["xlat"], g, dbx), dtype=np.float64)
        ns["mddom"].xlat = FArr(xl, [dbx[0], dbx[2]])
        ns.update(irceideal=wl.irceideal, calday=wl.calday, dayspy=wl.dayspy,
                  ptrop=FArr.alloc([(g.jci1, g.jci2), (g.ici1, g.ici2)]),
                  ktrop=FArr.alloc([(g.jci1, g.jci2), (g.ici1, g.ici2)], "int"),
                  kmxpbl=FArr.alloc([(g.jci1, g.jci2), (g.ici1, g.ici2)], "int"))
        # ---- the single-rank halo exchange --------------------------------------------------------------------
        ns["exchange_lr"] = lambda a, nex, j1, j2, i1, i2, *k: self._exchange(a, nex, j1, j2, i1, i2, True, False)
        ns["exchange_bt"] = lambda a, nex, j1, j2, i1, i2, *k: self._exchange(a, nex, j1, j2, i1, i2, False, True)
        ns["exchange_lrbt"] = lambda a, nex, j1, j2, i1, i2, *k: self._exchange(a, nex, j1, j2, i1, i2, True, True)
        if comm is not None:
            # the reference's own exchange routines (MPI-3 variants, Main/mpplib/mod_mppparam.F90:3809-3878,
            # 4257-4309, 4661-4712) over an emulated neighbourhood collective
            mst = F.preprocess(open(os.path.join(REF, "Main/mpplib/mod_mppparam.F90")).read(), defines=("USE_MPI3",))
            mr = F.find_routines(mst)
            tr0 = F.Translator({"ml", "sdata", "rdata", "counts", "displs"})
            for n in ("real8_3d_exchange_left_right_bottom_top", "real8_3d_exchange_left_right",
                      "real8_3d_exchange_bottom_top"):
                exec(F.compile_source(tr0.routine(mr[n]), f"<mod_mppparam.F90:{n}>"), ns)
            nbrs = [g.left, g.right, g.bottom, g.top]
            ns["mpi_neighbor_alltoallv"] = lambda sd, sc, sdis, st, rd, rc, rdis, rt, cm, err: \
                comm.neighbor_alltoallv(rank, nbrs, sd, sc, sdis, rd)
            ns.update(mpi_real8=0, cartesian_communicator=0, mpierr=0, mpi_success=0)
            if wl.do_bdy and wl.mo_spectral_nudge:
                # the reference's own row_reduce / column_reduce (Main/mpplib/mod_mppparam.F90:20618-20664) on an
                # emulated mpi_allreduce over the row / column communicators (:1459-1469: row colour = loci,
                # column colour = locj); MPI leaves the order of the sum open -- rank order here
                for n in ("real8_row_reduce", "real8_column_reduce"):
                    exec(F.compile_source(tr0.routine(mr[n]), f"<mod_mppparam.F90:{n}>"), ns)
                ns["size"] = lambda m, d: int(m.a.shape[m.a.ndim - d])
                ns.update(mpi_sum=0, cartesian_row_communicator="row", cartesian_column_communicator="col")
                groups = {"row": [q for q in range(px * py) if q % py == rank % py],
                          "col": [q for q in range(px * py) if q // py == rank // py]}
                ns["mpi_allreduce"] = lambda m, gl_, count, ty, op, cm, err: comm.allreduce(rank, groups[cm], m, gl_, count)
                ns["row_reduce"] = ns["real8_row_reduce"]
                ns["column_reduce"] = ns["real8_column_reduce"]
            ns["exchange_lr"] = ns["real8_3d_exchange_left_right"]
            ns["exchange_bt"] = ns["real8_3d_exchange_bottom_top"]
            ns["exchange_lrbt"] = ns["real8_3d_exchange_left_right_bottom_top"]
        ns["morelax"] = lambda j1, j2, i1, i2, ba, f, x: (
            ns["morelax_fraction"](j1, j2, i1, i2, ba, f, x) if isinstance(x, float)
            else ns["morelax_external"](j1, j2, i1, i2, ba, f, x))      # interface morelax, Main/mod_bdycod.F90:124-127
        ns["chem_bdyval"] = lambda u, v: ns["chem_bdyval_uncoupled"](u, v)   # interface, mod_che_bdyco.F90:75-78
        # ---- translate and load the reference routines ---------------------------------------------------------
        arrays = {k for k, v in ns.items() if isinstance(v, FArr)}
        arrays |= {"qc", "qi", "qr", "qs", "tke", "tkex", "tketen", "trac", "chiten", "chemt", "chib0", "chib1",
                   "hefc", "fcx", "tnudge", "cnudge", "qxzeroval", "qxcheckval", "ten0", "qen0", "chiten0"}
        tr = F.Translator(arrays)
        self.sources = {}
        pf = F.find_routines(F.preprocess(open(os.path.join(REF, "Share/pfwsat.inc")).read()))
        todo = [("Share/pfwsat.inc", pf, ["pfwsat"])]
        for rel, names in SOURCES.items():
            todo.append((rel, F.find_routines(F.preprocess(open(os.path.join(REF, rel)).read())), names))
        for rel, routines, names in todo:
            for n in names:
                src = tr.routine(routines[n])
                self.sources[n] = src
                exec(F.compile_source(src, f"<{rel}:{n}>"), ns)
        assert isinstance(ns["qxcheckval"], FArr) and isinstance(ns["anorth"], FArr)   # array parameters read from the source
        if dump_dir:
            os.makedirs(dump_dir, exist_ok=True)
            for n, src in self.sources.items():
                open(os.path.join(dump_dir, n + ".py"), "w").write(src)

    # exchange_lr/_bt/_lrbt for one rank: the rank is its own neighbour in a periodic
    # direction (mod_mppparam.F90:3809-3878 with ma%left == myid), otherwise nothing happens
    def _exchange(self, a, nex, j1, j2, i1, i2, lr, bt):
        wl = self.wl
        band = wl.i_band == 1 or wl.i_crm == 1
        crm = wl.i_crm == 1
        if a.nd == 2:
            views = [a.a[None]]
        elif a.nd == 3:
            views = [a.a]
        else:
            views = [a.a[n] for n in range(a.a.shape[0])]
        jl, il = a.lb[0], a.lb[1]
        for v in views:
            if lr and band:
                for x in range(1, nex + 1):
                    v[:, i1 - il:i2 - il + 1, j1 - x - jl] = v[:, i1 - il:i2 - il + 1, j2 - (x - 1) - jl]
                    v[:, i1 - il:i2 - il + 1, j2 + x - jl] = v[:, i1 - il:i2 - il + 1, j1 + (x - 1) - jl]
            if bt and crm:
                for x in range(1, nex + 1):
                    v[:, i1 - x - il, j1 - jl:j2 - jl + 1] = v[:, i2 - (x - 1) - il, j1 - jl:j2 - jl + 1]
                    v[:, i2 + x - il, j1 - jl:j2 - jl + 1] = v[:, i1 + (x - 1) - il, j1 - jl:j2 - jl + 1]

    # ---- driving it -------------------------------------------------------------------------------------------
    def call(self, name, *args):
        return self.ns[name](*args)

    def step(self, n=1):
        for _ in range(n):
            self.ns["moloch"]()

    def get(self, name) -> np.ndarray:
        """Owned cells on the global grid, like Oracle.get."""
        wl, g, H = self.wl, self.g, self.H
        ns = self.ns
        a = {"zeta": ns["z"], "pf3d": ns["atms"].pf3d, "th3d": ns["atms"].th3d, "rhb3d": ns["atms"].rhb3d,
             "wpx3d": ns["atms"].wpx3d, "rhox2d": ns["atms"].rhox2d, "tp2d": ns["atms"].tp2d,
             "th700": ns["atms"].th700}.get(name)
        if a is None and name[1:6] == "diag_":
            a = getattr(ns[name[:5]], name[6:])
        if a is None:
            a = ns[name]
        stag = H.ALLOC[name][0] if name in H.ALLOC else "cross"
        own = g.ext(stag, 0, 0)
        b = a.bounds()
        j1, j2, i1, i2 = max(own[0], b[0][0]), min(own[1], b[0][1]), max(own[2], b[1][0]), min(own[3], b[1][1])
        lead = a.a.shape[:-2]
        out = np.zeros(lead + (wl.iy, wl.jx))
        out[..., i1 - 1:i2, j1 - 1:j2] = a.a[..., i1 - b[1][0]:i2 - b[1][0] + 1, j1 - b[0][0]:j2 - b[0][0] + 1]
        return out


class SetupRun:
    """The reference's set-up routines on one rank: model_zitaf/model_zitah and the metric functions
    (Share/mod_zita.F90), compute_moloch_static (internal to `param`, Main/mod_params.F90:3316-3395),
    init_moloch (Main/mod_moloch.F90:201-308), setup_bdywt and paicompute (Main/mod_bdycod.F90:4033-4047,
    3762-3796), from the workload's file-like inputs (terrain, map factors, latitudes, ps, t, qv).  Limited-area
    domains only (every halo exchange is then a no-op on one rank)."""

    def __init__(self, wl, P: dict):
        from regcm_b200 import hostmodel as H
        from regcm_b200 import synthetic as S
        from regcm_b200.decomp import make_geom
        assert wl.i_band == 0 and wl.i_crm == 0
        self.wl = wl
        g = self.g = make_geom(wl.jx, wl.iy, wl.kz, 0, 0, 1, 1, 0)
        ns = self.ns = dict(INTRINSICS)
        ex = F.Expr(set())
        for rel in ("Share/mod_constants.F90", "Main/mpplib/mod_runparams.F90", "Main/mod_moloch.F90"):
            for line in F.module_parameters(F.preprocess(open(os.path.join(REF, rel)).read()), ex):
                try:
                    exec(F.compile_source(line, rel), ns)
                except Exception:
                    pass
        kz = wl.kz
        for n in ("jde1", "jde2", "ide1", "ide2", "jdi1", "jdi2", "idi1", "idi2", "jce1", "jce2", "ice1", "ice2",
                  "jci1", "jci2", "ici1", "ici2"):
            ns[n] = getattr(g, n)
        ns.update(kz=kz, kzp1=kz + 1, kzm1=kz - 1, nqx=wl.nqx, ntr=wl.ntr, ipptls=wl.ipptls, ibltyp=wl.ibltyp,
                  ichem=int(wl.ntr > 0), dtsec=wl.dt, dx=wl.dx, rdx=1.0 / wl.dx, mo_ztop=wl.mo_ztop, mo_h=wl.mo_h,
                  mo_a0=wl.mo_a0, mo_nadv=wl.mo_nadv, mo_nsound=wl.mo_nsound, mo_divdamp=bool(wl.mo_divdamp),
                  mo_divfilter=bool(wl.mo_divfilter), iproj="ROTLLR" if wl.lrotllr else "LAMCON",
                  jcross1=1, jcross2=wl.jx - 1, icross1=1, icross2=wl.iy - 1, irceideal=0, moloch_realcase=True,
                  myid=0, italk=0, rayzd=0.0, nspgx=wl.nspgx,
                  ma=_Obj(has_bdyleft=True, has_bdyright=True, has_bdybottom=True, has_bdytop=True, bandflag=False,
                          crmflag=False))
        for n in ("exchange", "exchange_lr", "exchange_bt", "exchange_lrbt"):
            ns[n] = lambda *a: None          # one rank, no periodic direction: nothing to exchange

        def box(stag):
            return g.ext(stag, 0, 0)

        def arr(stag, nk=0, fill=None, nspec=0):
            b = box(stag)
            bnds = [(b[0], b[1]), (b[2], b[3])] + ([(1, nk)] if nk else []) + ([(1, nspec)] if nspec else [])
            a = FArr.alloc(bnds)
            if fill is not None:
                a.a[...] = np.asarray(fill)[..., b[2] - 1:b[3], b[0] - 1:b[1]]
            return a

        md = ns["mddom"] = _Obj()
        for n in ("ht", "htu", "htv", "msfx", "msfu", "msfv", "ulat", "vlat", "xlat"):
            setattr(md, n, arr("dot", fill=P[n]))
        md.xlon = arr("dot")
        md.hx, md.hy = arr("u"), arr("v")
        md.rlat = FArr(np.asarray(P["rlat"], dtype=np.float64).copy(), [1])
        mo = ns["mo_atm"] = _Obj(zeta=arr("cross", kz), fmz=arr("cross", kz), rfmzu=arr("u", kz), rfmzv=arr("v", kz),
                                 fmzf=arr("cross", kz + 1), zetaf=arr("cross", kz + 1), dz=arr("cross", kz),
                                 pai=arr("cross", kz), tetav=arr("cross", kz), u=arr("u", kz), ux=arr("cross", kz),
                                 v=arr("v", kz), vx=arr("cross", kz), w=arr("cross", kz + 1), tvirt=arr("cross", kz),
                                 p=arr("cross", kz), t=arr("cross", kz, fill=P["t"]), rho=arr("cross", kz),
                                 qx=arr("cross", kz, fill=P["qx"], nspec=wl.nqx), qs=arr("cross", kz),
                                 uten=arr("cross", kz), vten=arr("cross", kz), tten=arr("cross", kz),
                                 qxten=arr("cross", kz, nspec=wl.nqx), tke=None, tketen=None,
                                 trac=arr("cross", kz, nspec=max(wl.ntr, 1)), chiten=arr("cross", kz, nspec=max(wl.ntr, 1)))
        ns["sfs"] = _Obj(psa=arr("cross", fill=P["ps"]), tg=arr("cross"), t2m=arr("cross"))
        ns["zita"], ns["zitah"] = FArr.alloc([(1, kz + 1)]), FArr.alloc([(1, kz)])
        # allocate_moloch (Main/mod_moloch.F90:159-199)
        ns.update(coru=arr("u"), corv=arr("v"), gzitak=FArr.alloc([(1, kz + 1)]), gzitakh=FArr.alloc([(1, kz)]),
                  xkdamp=FArr.alloc([(1, kz)]), xknu=FArr.alloc([(1, kz)]))
        for n in ("mx2", "rmx", "rmu", "rmv"):
            ns[n] = arr("dot")
        ns["bdywtu"] = FArr.alloc([(g.jdi1, g.jdi2), (g.ici1, g.ici2), (1, kz)])
        ns["bdywtv"] = FArr.alloc([(g.jci1, g.jci2), (g.idi1, g.idi2), (1, kz)])
        ns["bdywtw"] = FArr.alloc([(g.jci1, g.jci2), (g.ici1, g.ici2), (1, kz)])
        for n in ("mu", "mv", "mx", "rlat", "hx", "hy", "xlat", "xlon", "ht", "ps", "ts", "t2m", "fmz", "rfmzu", "rfmzv",
                  "fmzf", "pai", "tetav", "u", "ux", "v", "vx", "w", "tvirt", "z", "p", "t", "rho", "qx", "qsat", "uten",
                  "vten", "tten", "qxten", "qv", "qvten", "qc", "qi", "qr", "qs", "trac", "chiten", "tke", "tketen"):
            ns[n] = None       # module pointers, associated by init_moloch's assignpnt calls
        T = S.bdycon_setup(wl)
        dbox = g.ext("dot", 0, 0)
        for which, nm in (("cr", "ba_cr"), ("ud", "ba_ud"), ("vd", "ba_vd")):
            ib = np.array(H.cut(T["ibnd"][which], g, dbox), dtype=np.int64)
            ns[nm] = _Obj(ibnd=FArr(ib, [dbox[0], dbox[2]]), havebound=bool((ib > 0).any()))
        if wl.nspgx > 0:
            ns["hefc"] = FArr(np.asarray(P["hefc"], dtype=np.float64).reshape(kz, wl.nspgx).copy(), [1, 1])
        arrays = {k for k, v in ns.items() if isinstance(v, FArr)} | {
            "mu", "mv", "mx", "w", "zita", "zitah", "hefc", "mask", "zitaf"}
        tr = F.Translator(arrays)
        self.sources = {}
        zst = F.preprocess(open(os.path.join(REF, "Share/mod_zita.F90")).read())
        zr = F.find_routines(zst)
        todo = [(zr[n], "Share/mod_zita.F90") for n in ("zfz", "bzita", "bzitap", "gzita", "gzitap", "md_fmz_h",
                                                         "md_zeta_h", "md_fmz", "md_zeta", "model_zitaf", "model_zitah")]
        pst = F.preprocess(open(os.path.join(REF, "Main/mod_params.F90")).read())
        todo.append((F.find_internal(pst, "compute_moloch_static"), "Main/mod_params.F90"))
        mr = F.find_routines(F.preprocess(open(os.path.join(REF, "Main/mod_moloch.F90")).read()))
        todo.append((mr["init_moloch"], "Main/mod_moloch.F90"))
        br = F.find_routines(F.preprocess(open(os.path.join(REF, "Main/mod_bdycod.F90")).read()))
        todo += [(br["setup_bdywt"], "Main/mod_bdycod.F90"), (br["paicompute"], "Main/mod_bdycod.F90")]
        # two blocks of `init` (Main/mod_init.F90): the thermodynamic state after paicompute (:941-953 ...)
        # and the top sponge ffilt of the implicit solver (:1008-1026)
        ist = F.preprocess(open(os.path.join(REF, "Main/mod_init.F90")).read())
        todo.append((F.fragment(ist, "mo_atm%p(j,i,k) = (mo_atm%pai(j,i,k)**cpovr) * p00", "init_state_block", occurrence=1, back=1),
                     "Main/mod_init.F90"))
        todo.append((F.fragment(ist, "np = real(njcross*nicross,rk8)", "init_ffilt_block"), "Main/mod_init.F90"))
        pf = F.find_routines(F.preprocess(open(os.path.join(REF, "Share/pfwsat.inc")).read()))
        todo.append((pf["pfwsat"], "Share/pfwsat.inc"))
        zeros_cross = lambda nk: arr("cross", nk)
        mo.pf = zeros_cross(kz + 1)
        ns.update(njcross=wl.jx - 1, nicross=wl.iy - 1, sumall=lambda x: x, gmeanz=FArr.alloc([(1, kz)]),
                  ffilt=FArr.alloc([(1, kz)]), mo_zfilt_fac=0.8)       # Main/mod_init.F90:62
        for r, rel in todo:
            src = tr.routine(r)
            self.sources[r.name] = src
            exec(F.compile_source(src, f"<{rel}:{r.name}>"), ns)

    def run(self):
        ns, kz = self.ns, self.wl.kz
        ns["model_zitaf"](ns["zita"], ns["mo_ztop"])      # Main/mod_params.F90:2461-2463
        ns["model_zitah"](ns["zitah"], ns["mo_ztop"])
        ns["mo_dzita"] = ns["zita"][kz]
        ns["compute_moloch_static"]()
        ns["init_moloch"]()
        # Main/mod_init.F90:941: the hydrostatic Exner function of the initial state
        ns["paicompute"](ns["sfs"].psa, ns["mo_atm"].zeta, ns["mo_atm"].t, ns["qv"], ns["mo_atm"].pai)
        ns["init_state_block"]()
        ns["init_ffilt_block"]()
        return self

    def get(self, name) -> np.ndarray:
        ns, wl = self.ns, self.wl
        mo, md = ns["mo_atm"], ns["mddom"]
        a = {"zeta": mo.zeta, "fmz": mo.fmz, "rfmzu": mo.rfmzu, "rfmzv": mo.rfmzv, "fmzf": mo.fmzf, "zetaf": mo.zetaf,
             "hx": md.hx, "hy": md.hy, "pai": mo.pai, "p": mo.p, "qsat": mo.qs, "rho": mo.rho, "tvirt": mo.tvirt,
             "tetav": mo.tetav}.get(name)
        if a is None:
            a = ns[name]
        if a.nd == 1:
            return a.a.copy()
        b = a.bounds()
        out = np.zeros(a.a.shape[:-2] + (wl.iy, wl.jx))
        out[..., b[1][0] - 1:b[1][1], b[0][0] - 1:b[0][1]] = a.a
        return out


class BdySetupRun:
    """The reference's lateral-boundary set-up on one rank, executed from source: setup_boundaries
    (Main/mod_atm_interface.F90:384-532) for ba_cr/ba_ud/ba_vd, setup_bdycon's MOLOCH branch with the Lehmann
    coefficients (Main/mod_bdycod.F90:478-568; relax_coefficients, Main/mpplib/mod_runparams.F90:645-697) and
    lowpass_init (Main/mod_bdycod.F90:3844-3896).  Input: the level heights zeta."""

    def __init__(self, wl, zeta: np.ndarray, lehmann: bool = True):
        """lehmann=False: the exponential branch -- exponential_nudging (Main/mpplib/mod_runparams.F90:611-636)
        and spline1d (Share/mod_spline.F90:387-459) are executed from source too, on MOLOCH's sigma/hsigma
        (Main/mod_params.F90:2460-2465)."""
        from regcm_b200.decomp import make_geom
        self.wl = wl
        g = self.g = make_geom(wl.jx, wl.iy, wl.kz, wl.i_band, 0, 1, 1, 0)
        ns = self.ns = dict(INTRINSICS)
        ex = F.Expr(set())
        for rel in ("Share/mod_constants.F90", "Main/mpplib/mod_runparams.F90", "Main/mod_bdycod.F90"):
            for line in F.module_parameters(F.preprocess(open(os.path.join(REF, rel)).read()), ex):
                try:
                    exec(F.compile_source(line, rel), ns)
                except Exception:
                    pass
        kz = wl.kz
        for n in ("jde1", "jde2", "ide1", "ide2", "jce1", "jce2", "ice1", "ice2", "jci1", "jci2", "ici1", "ici2"):
            ns[n] = getattr(g, n)
        band = wl.i_band == 1
        njcross, nicross = (wl.jx if band else wl.jx - 1), wl.iy - 1
        ns.update(kz=kz, jx=wl.jx, iy=wl.iy, jxm1=wl.jx - 1, iym1=wl.iy - 1, nspgx=wl.nspgx, nspgd=wl.nspgx,
                  njcross=njcross, nicross=nicross, ds=wl.ds_km, dx=wl.dx, dtsec=wl.dt, mo_nadv=wl.mo_nadv,
                  mo_h=wl.mo_h, idynamic=3, dtbdys=wl.dtbdys, dtrad=wl.dtrad, myid=0, italk=0,
                  mo_top_nudge=bool(wl.mo_top_nudge), mo_spectral_nudge=bool(wl.mo_spectral_nudge),
                  bdy_use_lehmann=bool(lehmann), iboudy=5, rtb=0.0, nztop=0, km=0, lm=0, cn0=0.0,
                  ma=_Obj(bandflag=band, crmflag=False), sumall=lambda x: x, vprntv=lambda *a: None,
                  ba_cr=_Obj(), ba_ud=_Obj(), ba_vd=_Obj())
        b = g.ext("cross", 0, 0)
        zt = FArr.alloc([(b[0], b[1]), (b[2], b[3]), (1, kz)])
        zt.a[...] = np.asarray(zeta)[:, b[2] - 1:b[3], b[0] - 1:b[1]]
        ns["mo_atm"] = _Obj(zeta=zt)
        # allocate_mod_bdycon (Main/mod_bdycod.F90:417-476)
        ns.update(hefc=FArr.alloc([(1, wl.nspgx), (1, kz)]), gmeanz=FArr.alloc([(1, kz)]), tnudge=FArr.alloc([(1, kz)]),
                  cnudge=FArr.alloc([(1, kz)]))
        for n in ("bvx", "bvy", "sx", "sy", "sxg", "syg", "px", "py"):
            ns[n] = None
        arrays = {"hefc", "gmeanz", "tnudge", "cnudge", "bvx", "bvy", "sx", "sy", "sxg", "syg", "px", "py", "anudge",
                  "coeff", "p", "q", "pp", "qq"}
        if not lehmann:
            from regcm_b200 import synthetic as S_
            # Share/mod_dynparam.F90:150-152 (namelist defaults); sigma, hsigma as Main/mod_params.F90:2460-2465
            ns.update(high_nudge=3.0, medium_nudge=2.0, low_nudge=1.0, kzp1=kz + 1, myid=1,
                      anudge=FArr.alloc([(1, kz)]), d_one=1.0, d_half=0.5)
            sg, hs = FArr.alloc([(1, kz + 1)]), FArr.alloc([(1, kz)])
            sg.a[...] = 1.0 - S_.model_zitaf(kz, wl.mo_ztop) / wl.mo_ztop
            hs.a[...] = 1.0 - S_.model_zitah(kz, wl.mo_ztop) / wl.mo_ztop
            ns.update(sigma=sg, hsigma=hs)
            arrays |= {"sigma", "hsigma", "nudge", "ncin", "zcin", "ycin", "xold", "yold", "y2", "xnew", "ynew"}
        tr = F.Translator(arrays)
        ar = F.find_routines(F.preprocess(open(os.path.join(REF, "Main/mod_atm_interface.F90")).read()))
        br = F.find_routines(F.preprocess(open(os.path.join(REF, "Main/mod_bdycod.F90")).read()))
        rr = F.find_routines(F.preprocess(open(os.path.join(REF, "Main/mpplib/mod_runparams.F90")).read()))
        self.sources = {}
        extra = ()
        if not lehmann:
            sr = F.find_routines(F.preprocess(open(os.path.join(REF, "Share/mod_spline.F90")).read()))
            # its internal function findwhere (find_routines ends the host routine at the first `end function`,
            # so the internal procedure is the tail of the host's body after `contains`)
            eb = rr["exponential_nudging"].body
            inner = F.find_routines(eb[eb.index("contains") + 1:] + ["end function findwhere"])
            extra = ((sr["spline1d"], "Share/mod_spline.F90"), (inner["findwhere"], "Main/mpplib/mod_runparams.F90"),
                     (rr["exponential_nudging"], "Main/mpplib/mod_runparams.F90"))
        for r, rel in extra + ((ar["setup_boundaries"], "Main/mod_atm_interface.F90"),
                       (rr["relax_coefficients"], "Main/mpplib/mod_runparams.F90"),
                       (br["lowpass_init"], "Main/mod_bdycod.F90"), (br["setup_bdycon"], "Main/mod_bdycod.F90")):
            src = tr.routine(r)
            self.sources[r.name] = src
            exec(F.compile_source(src, f"<{rel}:{r.name}>"), ns)

    def run(self):
        ns = self.ns
        ns["setup_boundaries"](False, False, ns["ba_cr"])      # Main/mod_params.F90:2233-2238 (cross = .false.)
        ns["setup_boundaries"](True, False, ns["ba_ud"])
        ns["setup_boundaries"](False, True, ns["ba_vd"])
        ns["setup_bdycon"]()
        return self


def reference_massck(wl, o) -> dict:
    """massck (Main/mod_massck.F90:58-372) executed from source on the oracle's current state (one rank).
    Its sums are local variables; they are observed through the routine's own `call sumall(x, y)` reductions,
    which a one-rank run turns into y = x: tdrym, tqmass, tqadv, tdadv, tcrai, tncrai, tqeva in call order."""
    from regcm_b200.decomp import make_geom
    g = make_geom(wl.jx, wl.iy, wl.kz, wl.i_band, wl.i_crm, 1, 1, 0)
    ns = dict(INTRINSICS)
    ex = F.Expr(set())
    for rel in ("Share/mod_constants.F90", "Main/mpplib/mod_runparams.F90", "Main/mod_massck.F90"):
        for line in F.module_parameters(F.preprocess(open(os.path.join(REF, rel)).read()), ex):
            try:
                exec(F.compile_source(line, rel), ns)
            except Exception:
                pass
    for n in ("jde1", "jde2", "ide1", "ide2", "jce1", "jce2", "ice1", "ice2", "jci1", "jci2", "ici1", "ici2"):
        ns[n] = getattr(g, n)
    seen = []

    def sumall(x):
        seen.append(x)
        return x
    kz = wl.kz

    def fa(name, stag, nk=0, nspec=0):
        b = g.ext(stag, 0, 0)
        a = np.asarray(o.get(name))[..., b[2] - 1:b[3], b[0] - 1:b[1]]
        return FArr(np.array(a, dtype=np.float64), [b[0], b[2]] + ([1] if nk else []) + ([1] if nspec else []))
    zf = fa("zetaf", "cross", kz + 1)
    dz = FArr(zf.a[:-1] - zf.a[1:], zf.lb)                        # Main/mod_params.F90:3389
    cb = g.ext("cross", 0, 0)
    zero2 = FArr.alloc([(cb[0], cb[1]), (cb[2], cb[3])])
    ns.update(kz=kz, nqx=wl.nqx, idynamic=3, dt=wl.dt, dx=wl.dx, dxsq=wl.dx * wl.dx, myid=0, italk=0, sumall=sumall,
              ma=_Obj(has_bdyleft=g.bl, has_bdyright=g.br, has_bdybottom=g.bb, has_bdytop=g.bt),
              mo_atm=_Obj(dz=dz, rho=fa("rho", "cross", kz), u=fa("u", "u", kz), v=fa("v", "v", kz),
                          qx=fa("qx", "cross", kz, wl.nqx)),
              crrate=zero2, ncrrate=zero2, sfs=_Obj(qfx=zero2), rcmtimer=_Obj(start=lambda: False),
              alarm_day=_Obj(act=lambda: False), syncro_dbg=_Obj(act=lambda: False),
              dryini=0.0, watini=0.0, dryerror=0.0, waterror=0.0, mcrai=0.0, mncrai=0.0, mevap=0.0, mdryadv=0.0,
              mqadv=0.0)
    r = F.find_routines(F.preprocess(open(os.path.join(REF, "Main/mod_massck.F90")).read()))["massck"]
    exec(F.compile_source(F.Translator({"crrate", "ncrrate", "dsigma"}).routine(r), "<Main/mod_massck.F90:massck>"), ns)
    ns["massck"]()
    return dict(zip(("tdrym", "tqmass", "tqadv", "tdadv", "tcrai", "tncrai", "tqeva"), seen))


SETUP_FIELDS = ["hx", "hy", "zeta", "fmz", "fmzf", "zetaf", "rfmzu", "rfmzv", "coru", "corv", "mx2", "rmx", "rmu", "rmv",
                "gzitak", "gzitakh", "xkdamp", "xknu", "bdywtu", "bdywtv", "bdywtw", "pai", "p", "qsat", "rho", "tvirt",
                "tetav", "ffilt"]


def reference_allocation_bounds(run: "ReferenceRun") -> dict:
    """allocate_atmosphere (Main/mod_atm_interface.F90:579-624) and allocate_moloch (Main/mod_moloch.F90:159-199)
    executed from source with the index ranges of `run`'s rank: name -> Fortran bounds of every array."""
    ns = dict(INTRINSICS)
    keys = [k for k, v in run.ns.items() if isinstance(v, (int, float, bool)) and not k.startswith("_")]
    ns.update({k: run.ns[k] for k in keys})
    mods = ("gzitak", "gzitakh", "laplacian", "bdywtu", "bdywtv", "bdywtw", "wwkw", "tetavf", "s", "zdiv2", "wz", "p0",
            "wfw", "wx", "zpby", "zpbw", "mx2", "rmx", "rmu", "rmv", "coru", "corv", "tkex", "ten0", "qen0", "chiten0",
            "ud", "vd", "xkdamp", "xknu")
    for n in mods:
        ns[n] = None
    tr = F.Translator(set())
    ast_ = F.preprocess(open(os.path.join(REF, "Main/mod_atm_interface.F90")).read())
    mst = F.preprocess(open(os.path.join(REF, "Main/mod_moloch.F90")).read())
    exec(F.compile_source(tr.routine(F.find_internal(ast_, "allocate_atmosphere")), "<allocate_atmosphere>"), ns)
    exec(F.compile_source(tr.routine(F.find_routines(mst)["allocate_moloch"]), "<allocate_moloch>"), ns)
    atm = _Obj()
    ns["allocate_atmosphere"](atm)
    ns["allocate_moloch"]()
    out = {"mo_atm%" + k: v.bounds() for k, v in vars(atm).items() if isinstance(v, FArr)}
    out.update({n: ns[n].bounds() for n in mods if isinstance(ns[n], FArr)})
    return out


def reference_set_nproc(jx, iy, kz, i_band, i_crm, nproc, myid, njxcpus=-1, niycpus=-1) -> dict:
    """set_nproc (Main/mpplib/mod_mppparam.F90:1250-1641) executed from source for one rank.  The MPI Cartesian
    topology calls are stubs with the standard's semantics (row-major coordinates, mpi_cart_shift with
    periodic wrap, mpi_proc_null outside a non-periodic dimension).  Returns the decomposition it computed."""
    ns = dict(INTRINSICS)
    seen = {}

    class Cart:
        def __init__(self, dims, periods):
            self.dims, self.periods = [int(x) for x in dims.a], [bool(x) for x in periods.a]

        def rank_of(self, c):
            c = list(c)
            for d in (0, 1):
                if c[d] < 0 or c[d] >= self.dims[d]:
                    if not self.periods[d]:
                        return -1
                    c[d] %= self.dims[d]
            return c[0] * self.dims[1] + c[1]

        def coords(self, r):
            return [r // self.dims[1], r % self.dims[1]]

    def mpi_cart_create(comm, nd, dims, periods, reorder, new, err):
        seen["dims"] = [int(x) for x in dims.a]
        return Cart(dims, periods)

    def mpi_cart_coords(comm, rank, nd, coords, err):
        coords.a[:] = comm.coords(rank)

    def mpi_cart_shift(comm, direction, disp, src, dst, err):
        c = comm.coords(myid)
        lo, hi = list(c), list(c)
        lo[direction] -= disp
        hi[direction] += disp
        return comm.rank_of(lo), comm.rank_of(hi)

    def fatal(*a):
        raise RuntimeError("fatal: " + str(a[-1]))
    ma = _Obj(location=FArr(np.zeros(2, dtype=np.int64), [1]))
    ns.update(ma=ma, mpi_proc_null=-1, nproc=nproc, myid=myid, jx=jx, iy=iy, kz=kz, kzp1=kz + 1, i_band=i_band,
              i_crm=i_crm, njxcpus=njxcpus, niycpus=niycpus, nsg=1, jxsg=jx, iysg=iy, mycomm=0, lreorder=False,
              iocpu=0, ccio=0, italk=0, ccid=0, mpierr=0, mpi_success=0, ifake=None, windows=None,
              mpi_cart_create=mpi_cart_create, mpi_comm_rank=lambda comm, r, err: myid,
              mpi_cart_coords=mpi_cart_coords, mpi_comm_split=lambda *a: 0, mpi_cart_shift=mpi_cart_shift,
              mpi_cart_rank=lambda comm, coords, r, err: comm.rank_of([int(x) for x in coords.a]),
              bcast=lambda *a: None, fatal=fatal, cartesian_communicator=None, cartesian_row_communicator=None,
              cartesian_column_communicator=None)
    st = F.preprocess(open(os.path.join(REF, "Main/mpplib/mod_mppparam.F90")).read(), defines=("USE_MPI3",))
    r = F.find_routines(st)["set_nproc"]
    exec(F.compile_source(F.Translator({"location"}).routine(r), "<mod_mppparam.F90:set_nproc>"), ns)
    ns["set_nproc"]()
    # setup_model_indexes (Main/mod_atm_interface.F90:182-382): the loop and ghost ranges of this rank
    ar = F.find_routines(F.preprocess(open(os.path.join(REF, "Main/mod_atm_interface.F90")).read()))
    exec(F.compile_source(F.Translator(set()).routine(ar["setup_model_indexes"]),
                          "<mod_atm_interface.F90:setup_model_indexes>"), ns)
    ns.update(idynamic=3, idiffu=1)
    ns["setup_model_indexes"]()
    idx = {k: ns[k] for k in ("jde1", "jde2", "ide1", "ide2", "jdi1", "jdi2", "idi1", "idi2", "jdii1", "jdii2", "idii1",
                              "idii2", "jce1", "jce2", "ice1", "ice2", "jci1", "jci2", "ici1", "ici2", "jce1ga",
                              "jce2ga", "ice1ga", "ice2ga", "jce1gb", "jce2gb", "ice1gb", "ice2gb", "jde1ga", "jde2ga",
                              "ide1ga", "ide2ga", "jde1gb", "jde2gb", "ide1gb", "ide2gb", "jci1ga", "jci2ga", "ici1ga",
                              "ici2ga") if k in ns}
    out = {k: ns[k] for k in ("jxp", "iyp", "global_dot_jstart", "global_dot_jend", "global_dot_istart",
                              "global_dot_iend", "global_cross_jstart", "global_cross_jend", "global_cross_istart",
                              "global_cross_iend")}
    out.update(left=ma.left, right=ma.right, bottom=ma.bottom, top=ma.top, has_bdyleft=bool(ma.has_bdyleft),
               has_bdyright=bool(ma.has_bdyright), has_bdybottom=bool(ma.has_bdybottom),
               has_bdytop=bool(ma.has_bdytop), dims=seen.get("dims", [1, 1]), indexes=idx)
    return out


class _Comm:
    """mpi_neighbor_alltoallv on a 2-D Cartesian communicator, for ranks that run as threads of this
    process: neighbour order (dim 0 -, dim 0 +, dim 1 -, dim 1 +) = (left, right, bottom, top); what a rank
    sends towards + is what its + neighbour receives from -."""

    def __init__(self, n):
        import threading
        self.box, self.bar = {}, threading.Barrier(n, timeout=600)

    def neighbor_alltoallv(self, rank, nbrs, sdata, counts, displs, rdata):
        opp = (1, 0, 3, 2)
        for d in range(4):
            if nbrs[d] >= 0:
                c, o = int(counts.a[d]), int(displs.a[d])
                self.box[(rank, d)] = sdata.a[o:o + c].copy()
        self.bar.wait()
        for d in range(4):
            if nbrs[d] >= 0:
                c, o = int(counts.a[d]), int(displs.a[d])
                rdata.a[o:o + c] = self.box[(nbrs[d], opp[d])]
        self.bar.wait()


def _comm_allreduce(self, rank, group, m, gl_, count):
    """mpi_allreduce(SUM) of the first `count` elements of the contiguous arrays, added in rank order."""
    self.box[("red", rank)] = m.a.reshape(-1)[:count].copy()
    self.bar.wait()
    acc = None
    for q in group:
        acc = self.box[("red", q)].copy() if acc is None else acc + self.box[("red", q)]
    gl_.a.reshape(-1)[:count] = acc
    self.bar.wait()


_Comm.allreduce = _comm_allreduce


class MultiRankReference:
    """The reference's `moloch` on px x py ranks, one thread per rank, halos through the reference's own
    exchange routines: pins the decomposition and halo semantics (and the oracle's emulation of them)."""

    def __init__(self, wl, oracle, boundary, px, py):
        self.wl, self.n = wl, px * py
        comm = self.comm = _Comm(self.n)
        self.ranks = [ReferenceRun(wl, oracle, boundary, px=px, py=py, rank=r, comm=comm) for r in range(self.n)]

    def step(self, nsteps=1):
        import threading
        errs = []

        def run(r):
            try:
                r.step(nsteps)
            except BaseException as e:  # noqa: BLE001
                errs.append(e)
                self.comm.bar.abort()      # release the ranks waiting in a collective
        ts = [threading.Thread(target=run, args=(r,)) for r in self.ranks]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]

    def get(self, name):
        out = None
        for r in self.ranks:
            a = r.get(name)
            out = a if out is None else out + a       # owned cells are disjoint, the rest is zero
        return out


class _Mask:
    """logical(j,i) array of cbound_area (bsouth ...): Fortran-bounded boolean lookup."""

    def __init__(self, m, box):
        self.m, self.j0, self.i0 = m, box[0], box[2]

    def __getitem__(self, idx):
        return bool(self.m[idx[1] - self.i0, idx[0] - self.j0])

    def __call__(self, j, i):
        return self[j, i]


# ---- golden fixtures ------------------------------------------------------------------------------------------------
GOLDEN_FIELDS = ["u", "v", "w", "pai", "tetav", "t", "qx", "ux", "vx", "tvirt", "p", "rho", "qsat", "ps"]


def digest(a: np.ndarray) -> dict:
    a = np.ascontiguousarray(a, dtype=np.float64) + 0.0      # -0.0 -> +0.0: equal values, equal bytes
    return {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "sum": float(a.sum()), "absmax": float(np.abs(a).max())}


def golden_cases():
    from regcm_b200 import synthetic as S
    lam = S.small(S.WORKLOADS["cordex25"], 16, 14, 8, ntr=1, nspgx=4, mo_nsound=3)
    return {
        "periodic_hills": (S.small(S.WORKLOADS["isc24_small"], 14, 12, 8, oro="sine", oro_h=700.0, msf_amp=0.03,
                                   clat=30.0, mo_nsound=3), 2),
        "limited_area": (lam, 2),
        "limited_area_rotllr": (S.small(lam, 16, 14, 8, lrotllr=1), 1),
        "limited_area_boundary": (S.small(lam, 16, 14, 8, do_bdy=1, present_qc=1, present_qi=1, mo_top_nudge=1,
                                          ichebdy=1, do_slice=1, icldmstrat=1), 2),
        # spectral nudging active every step (dtrad == dt), no ICBC condensate
        "limited_area_spectral": (S.small(lam, 18, 16, 6, do_bdy=1, mo_top_nudge=1, mo_spectral_nudge=1, ichebdy=1,
                                          ds_km=150.0, dtrad=150.0, dt=150.0), 3),
        # periodic in j (band) with south/north boundaries and sponge; no divergence damping / filter; vapour only
        "band_boundary": (S.small(lam, 16, 14, 6, i_band=1, oro="sine", do_bdy=1, present_qc=1, mo_top_nudge=1,
                                  mo_ztop=30000.0), 2),
        "no_damp_no_filter": (S.small(lam, 14, 12, 6, mo_divdamp=0, mo_divfilter=0, ntr=0), 1),
        "vapour_only": (S.small(lam, 14, 12, 6, ipptls=0, nqx=1, ntr=0, do_bdy=1), 2),
        # tendency diagnostics of dynamical_core and boundary (idiag, ichdiag)
        "limited_area_diag": (S.small(lam, 16, 14, 6, do_bdy=1, present_qc=1, mo_top_nudge=1, mo_ztop=30000.0, idiag=1,
                                      ichdiag=1), 2),
        # UW-PBL TKE advected by the dycore, with its boundary values
        "limited_area_tke": (S.small(lam, 16, 14, 8, do_bdy=1, present_qc=1, ibltyp=2, tkemin=1.0e-4, ipptls=1,
                                     nqx=2), 2),
    }


def case_fields(wl):
    return GOLDEN_FIELDS + (["trac"] if wl.ntr else []) + (["tke"] if wl.ibltyp == 2 else []) + \
        (["pf3d", "th3d", "rhb3d", "wpx3d", "rhox2d", "tp2d", "th700", "ptrop", "ktrop", "kmxpbl"] if wl.do_slice else []) + \
        ((["tdiag_adh", "qdiag_adh"] + (["tdiag_bdy", "qdiag_bdy"] if wl.do_bdy else [])) if wl.idiag > 0 else []) + \
        ((["cadvhdiag"] + (["cbdydiag"] if wl.do_bdy else [])) if wl.ichdiag > 0 and wl.ntr > 0 else [])


def setup_cases():
    from regcm_b200 import synthetic as S
    lam = S.small(S.WORKLOADS["cordex25"], 16, 14, 8, ntr=1, nspgx=4)
    return {"setup_limited_area": lam, "setup_limited_area_rotllr": S.small(lam, 18, 12, 7, lrotllr=1, nspgx=5)}


def run_setup_case(wl):
    """(reference set-up run, oracle) from the same file-like inputs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from regcm_b200 import synthetic as S
    from util import make_oracle_bdy
    o, _ = make_oracle_bdy(wl)
    return SetupRun(wl, S.make_primary(wl)).run(), o


def run_case(wl, nsteps, dump_dir=None):
    """(reference run, oracle) after nsteps of `moloch` from the same initial state."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import make_oracle_bdy
    o, B = make_oracle_bdy(wl)
    r = ReferenceRun(wl, o, B, dump_dir=dump_dir)
    r.step(nsteps)
    o.step(nsteps)
    return r, o


def main():
    if not available():
        raise SystemExit(f"{REF} not found: the reference sources are needed to run them")
    out = {}
    for name, (wl, nsteps) in golden_cases().items():
        r, o = run_case(wl, nsteps, dump_dir=os.path.join(ROOT, "oracle", "_ref", "translated"))
        fields = case_fields(wl)
        out[name] = {"steps": nsteps, "grid": [wl.jx, wl.iy, wl.kz], "fields": {f: digest(r.get(f)) for f in fields}}
        agree = {f: bool(np.array_equal(r.get(f), o.get(f))) for f in fields}
        print(name, "reference == oracle:", agree)
    for name, wl in setup_cases().items():
        sr, o = run_setup_case(wl)
        out[name] = {"steps": 0, "grid": [wl.jx, wl.iy, wl.kz], "fields": {f: digest(sr.get(f)) for f in SETUP_FIELDS}}
        print(name, "reference == oracle:", {f: bool(np.array_equal(sr.get(f), o.get(f))) for f in SETUP_FIELDS})
    path = os.path.join(ROOT, "tests", "golden", "reference_moloch.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
