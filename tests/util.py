"""Shared helpers of the parity tests: one oracle world and one (or more) GPU
ranks fed with exactly the same static fields and initial state."""
from __future__ import annotations

import numpy as np

from oracle.oracle import Oracle
from regcm_b200 import synthetic as S
from regcm_b200.moloch import PROFILE_NAMES, STATE_FIELDS, STATIC_FIELDS, MolochB200

# Library object handed to MolochB200(lib=...).  None: the product's CUDA library
# (GPU tests).  tests/test_emu_full.py points it at the host build of the same
# sources (tests/emu_lib.py) and re-runs the GPU tests' bodies on the CPU.
LIB = None

PROGNOSTIC = ["u", "v", "w", "pai", "tetav", "t", "qx", "ux", "vx"]
DIAGNOSTIC = ["tvirt", "p", "rho", "qsat", "ps"]


def make_oracle(wl, px=1, py=1, checked=False, ffilt=None):
    P = S.make_primary(wl)
    o = Oracle(wl, px=px, py=py, checked=checked)
    o.load_primary(P)
    if ffilt is not None:
        o.set("ffilt", ffilt)
    return o, P


def make_oracle_bdy(wl, px=1, py=1, checked=False, same=False):
    """Oracle with the boundary / slice / TKE extension (Workload.needs_ext):
    the b0/b1 buffers are generated from the oracle's own initial state."""
    P = S.make_primary(wl)
    if wl.nspgx > 0:
        P["fcx"] = S.chem_fcx(wl)
    o = Oracle(wl, px=px, py=py, checked=checked)
    o.load_primary(P)
    if wl.ibltyp == 2:
        o.set("tke", S.make_tke(wl, o.get("zetaf")))
    B = {}
    if wl.do_bdy:
        base = {n: o.get(n) for n in ("u", "v", "t", "pai", "qx", "ps")}
        B = S.make_boundary(wl, base)
        if same:
            for n in list(B):
                if n.endswith("1"):
                    B[n] = B[n[:-1] + "0"].copy()
        o.load_boundary(B)
    return o, B


def bdy_tables_from_oracle(wl, o):
    """setup_bdycon results for MolochB200(bdy=...): the ibnd planes and hefc from
    the host model, the sumall-dependent profiles (nztop, tnudge, cnudge) from the
    oracle so that both sides use identical bits."""
    if not wl.do_bdy:
        return None
    T = S.bdycon_setup(wl, o.get("zeta"))
    T["nztop"] = o.get_int("nztop")
    if wl.mo_top_nudge:
        T["tnudge"] = o.get("tnudge")
    if wl.mo_spectral_nudge:
        T["cnudge"] = o.get("cnudge")
        assert (T["km"], T["lm"]) == (o.get_int("km"), o.get_int("lm"))
    if wl.nspgx > 0:
        T["hefc"] = o.get("hefc").reshape(wl.kz, wl.nspgx)
    return T


def oracle_inputs(o, wl):
    names = [n for n in STATIC_FIELDS + STATE_FIELDS if not (n == "trac" and wl.ntr == 0)]
    fields = {n: o.get(n) for n in names}
    profiles = {n: o.get(n) for n in PROFILE_NAMES}
    return fields, profiles


def make_gpu(wl, fields, profiles, rank=0, nranks=1, px=None, py=None, device=-1):
    m = MolochB200(wl, rank=rank, nranks=nranks, px=px, py=py, device=device, lib=LIB).allocate_moloch()
    if wl.lrotllr:
        profiles = dict(profiles)
    m.init_moloch(fields, profiles)
    return m


def make_gpu_bdy(wl, o, B, device=-1):
    """One GPU rank with the boundary / slice / TKE extension, fed from the oracle `o`."""
    fields, profiles = oracle_inputs(o, wl)
    if wl.lrotllr:
        profiles["rlat"] = S.make_primary(wl)["rlat"]
    m = MolochB200(wl, device=device, bdy=bdy_tables_from_oracle(wl, o), lib=LIB).allocate_moloch()
    if wl.ibltyp == 2:
        fields["tke"] = o.get("tke")
    if wl.do_slice:
        fields["zetaf"] = o.get("zetaf")
        fields["xlat"] = o.get("xlat")
    m.init_moloch(fields, profiles)
    if wl.do_slice:
        m.set_calday(wl.calday, wl.dayspy)
    if wl.do_bdy:
        m.load_boundary(B)
        m.set_xbctime(o.get_xbctime())
    return m


def compare(o, m, names, exact=True, rtol=0.0, label=""):
    bad = []
    for n in names:
        if n == "trac" and m.wl.ntr == 0:
            continue
        a, b = o.get(n), m.get_global(n)
        if exact:
            if not np.array_equal(a, b):
                d = np.abs(a - b)
                bad.append(f"{label}{n}: max abs diff {d.max():.3e} at {np.unravel_index(d.argmax(), d.shape)} "
                           f"(ref {a.flat[d.argmax()]:.17g})")
        else:
            den = np.maximum(np.abs(a), 1e-300)
            r = np.abs(a - b) / den
            if not (r.max() <= rtol):
                bad.append(f"{label}{n}: max rel diff {r.max():.3e} > {rtol}")
    assert not bad, "\n".join(bad)
