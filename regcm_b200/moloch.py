"""Host-side mirror of RegCM's `mod_moloch` over the C ABI (include/moloch_b200.h).

The reference's interface for this path is three argument-less module
procedures acting on module state (Main/mod_moloch.F90:127):

    allocate_moloch  :159     init_moloch  :201     moloch  :312

`MolochB200` keeps those names and meanings; the module state (`mo_atm`,
`mddom`, the &molochparam knobs, the decomposition `ma`) is the object's
state.  In a RegCM build the same C entry points are called from the Fortran
shim in INTEGRATION.md; this Python class exists so that tests and benchmarks
can drive exactly that ABI without a Fortran toolchain.

There is no CPU fallback: construction fails if the CUDA library is missing
or no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import hostmodel as H
from .decomp import Geom, default_cpus_per_dim, make_geom

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOLOCH_B200_LIB") or os.path.join(_HERE, "libmoloch_b200.so")

FIELDS = ["u", "v", "w", "pai", "tetav", "t", "qx", "trac", "ux", "vx", "tvirt", "p", "rho", "qsat", "ps",
          "zeta", "fmz", "fmzf", "rfmzu", "rfmzv", "hx", "hy", "msfx", "msfu", "msfv", "coru", "corv",
          "bdywtu", "bdywtv", "bdywtw", "tten", "uten", "vten", "qxten", "chiten", "s", "zdiv2", "wx", "wz",
          "p0", "tetavf",
          # ABI v2: TKE, boundary buffers (b0/b1), mkslice outputs
          "tke", "tketen", "tkex",
          "dub0", "dub1", "dvb0", "dvb1", "xtb0", "xtb1", "xpaib0", "xpaib1", "xqb0", "xqb1",
          "xlb0", "xlb1", "xib0", "xib1", "xpsb0", "xpsb1", "chib0", "chib1",
          "pf3d", "th3d", "rhb3d", "wpx3d", "rhox2d", "tp2d", "th700", "zetaf",
          "xlat", "ptrop", "ktrop", "kmxpbl",
          "ten0", "qen0", "tdiag_adh", "qdiag_adh", "tdiag_bdy", "qdiag_bdy", "chiten0", "cadvhdiag", "cbdydiag"]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}
PROFILES = ["gzitak", "gzitakh", "ffilt", "xkdamp", "xknu", "rlat"]
PROFILE_ID = {n: i for i, n in enumerate(PROFILES)}
TABLES = ["hefc", "tnudge", "cnudge", "fcx", "bvx", "bvy"]
TABLE_ID = {n: i for i, n in enumerate(TABLES)}
IBND_ID = {"cr": 0, "ud": 1, "vd": 2}
SPECIES = {"qx", "trac", "qxten", "chiten", "chib0", "chib1", "chiten0", "cadvhdiag", "cbdydiag"}

# every symbol include/moloch_b200.h declares
ABI_SYMBOLS = [
    "moloch_b200_last_error", "moloch_b200_abi_version", "moloch_b200_device_count", "moloch_b200_create",
    "moloch_b200_destroy", "moloch_b200_comm_id", "moloch_b200_comm_init", "moloch_b200_set_stream",
    "moloch_b200_sync", "moloch_b200_set_field", "moloch_b200_get_field", "moloch_b200_set_profile",
    "moloch_b200_host_alloc", "moloch_b200_host_free", "moloch_b200_init", "moloch_b200_reset_tendencies",
    "moloch_b200_sound", "moloch_b200_advection", "moloch_b200_wafone", "moloch_b200_dynamical_core",
    "moloch_b200_diagnostics", "moloch_b200_status_update", "moloch_b200_step", "moloch_b200_profile_enable",
    "moloch_b200_profile_read", "moloch_b200_launch_count", "moloch_b200_device_bytes",
    "moloch_b200_halo_plan", "moloch_b200_p2p_blob_size", "moloch_b200_p2p_export", "moloch_b200_p2p_connect", "moloch_b200_set_async",
    "moloch_b200_set_table", "moloch_b200_set_ibnd", "moloch_b200_boundary", "moloch_b200_bdyval",
    "moloch_b200_set_xbctime", "moloch_b200_get_xbctime", "moloch_b200_bdy_shift", "moloch_b200_mkslice",
    "moloch_b200_massck", "moloch_b200_ps_check", "moloch_b200_set_calday", "moloch_b200_config_size",
    "moloch_b200_handoff", "moloch_b200_host_register", "moloch_b200_host_unregister", "moloch_b200_set_option",
]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "jx", "iy", "kz", "nqx", "ntr", "iqfrst", "jde1", "jde2", "ide1", "ide2", "jce1", "jce2", "ice1",
        "ice2", "has_bdy_left", "has_bdy_right", "has_bdy_bottom", "has_bdy_top", "bandflag", "crmflag",
        "nbr_left", "nbr_right", "nbr_bottom", "nbr_top", "rank", "nranks", "mo_nadv", "mo_nsound",
        "mo_divdamp", "mo_divfilter", "lrotllr", "ipptls", "device", "ibltyp")] + [
        (n, C.c_double) for n in ("dtsec", "dx", "mo_dzita")] + [(n, C.c_int32) for n in (
        "do_bdy", "nspgx", "present_qc", "present_qi", "mo_top_nudge", "mo_spectral_nudge", "nztop", "ichem",
        "ichebdy", "do_slice", "icldmstrat", "km", "lm", "do_massck")] + [
        (n, C.c_double) for n in ("dtbdys", "dtrad", "rhmin", "rhmax", "tkemin")] + [
        (n, C.c_int32) for n in ("irceideal", "idiag", "ichdiag", "niycpus")]


class Xfer(C.Structure):
    """moloch_b200_xfer (include/moloch_b200.h)."""
    _fields_ = [("field", C.c_int32), ("n", C.c_int32), ("host", C.c_void_p)] + [
        (n, C.c_int32) for n in ("jlo", "jhi", "ilo", "ihi", "klo", "khi")]


PHYSICS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32)


class MolochError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen the CUDA library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MolochError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the product has no CPU fallback)")
    _lib = bind_library(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL), LIB_PATH)
    return _lib


def load_fast_library():
    """The FMA-contracting build (libmoloch_b200_fast.so): same entry points, results within the fast-mode
    tolerances of SURVEY.md 8(c) instead of bit-identical."""
    path = os.path.join(_HERE, "libmoloch_b200_fast.so")
    if not os.path.exists(path):
        raise MolochError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    return bind_library(C.CDLL(path, mode=C.RTLD_LOCAL), path)


def bind_library(lib, path: str = "?"):
    """ctypes signatures of every entry of include/moloch_b200.h on a loaded library object."""
    ctx = C.c_void_p
    lib.moloch_b200_last_error.restype = C.c_char_p
    lib.moloch_b200_config_size.restype = C.c_uint64
    if int(lib.moloch_b200_config_size()) != C.sizeof(Config):
        raise MolochError(f"{path}: moloch_b200_config is {int(lib.moloch_b200_config_size())} bytes in the "
                          f"library, {C.sizeof(Config)} in this binding (stale build?)")
    lib.moloch_b200_create.argtypes = [C.POINTER(Config), C.POINTER(ctx)]
    lib.moloch_b200_destroy.argtypes = [ctx]
    lib.moloch_b200_comm_id.argtypes = [C.c_void_p]
    lib.moloch_b200_comm_init.argtypes = [ctx, C.c_void_p]
    lib.moloch_b200_set_stream.argtypes = [ctx, C.c_void_p]
    lib.moloch_b200_p2p_blob_size.restype = C.c_uint64
    lib.moloch_b200_p2p_export.argtypes = [ctx, C.c_void_p]
    lib.moloch_b200_p2p_connect.argtypes = [ctx, C.c_void_p, C.c_int]
    lib.moloch_b200_sync.argtypes = [ctx]
    lib.moloch_b200_set_option.argtypes = [ctx, C.c_char_p, C.c_int]
    lib.moloch_b200_set_async.argtypes = [ctx, C.c_int]
    xf = [ctx, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 6
    lib.moloch_b200_set_field.argtypes = xf
    lib.moloch_b200_get_field.argtypes = xf
    lib.moloch_b200_set_profile.argtypes = [ctx, C.c_int, C.c_void_p, C.c_int]
    lib.moloch_b200_handoff.argtypes = [ctx, C.POINTER(Xfer), C.c_int, C.POINTER(Xfer), C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]
    lib.moloch_b200_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
    lib.moloch_b200_host_free.argtypes = [C.c_void_p]
    lib.moloch_b200_host_register.argtypes = [C.c_void_p, C.c_uint64]
    lib.moloch_b200_host_unregister.argtypes = [C.c_void_p]
    lib.moloch_b200_set_table.argtypes = [ctx, C.c_int, C.c_void_p, C.c_int]
    lib.moloch_b200_set_ibnd.argtypes = [ctx, C.c_int, C.c_void_p] + [C.c_int] * 4
    lib.moloch_b200_set_calday.argtypes = [ctx, C.c_double, C.c_double]
    lib.moloch_b200_massck.argtypes = [ctx, C.c_void_p]
    lib.moloch_b200_ps_check.argtypes = [ctx, C.c_void_p, C.c_void_p]
    lib.moloch_b200_set_xbctime.argtypes = [ctx, C.c_double]
    lib.moloch_b200_get_xbctime.argtypes = [ctx]
    lib.moloch_b200_get_xbctime.restype = C.c_double
    for f in ("init", "reset_tendencies", "sound", "advection", "dynamical_core", "diagnostics", "status_update",
              "boundary", "bdyval", "bdy_shift", "mkslice"):
        getattr(lib, "moloch_b200_" + f).argtypes = [ctx]
    lib.moloch_b200_wafone.argtypes = [ctx, C.c_int, C.c_int]
    lib.moloch_b200_step.argtypes = [ctx, C.c_int]
    lib.moloch_b200_profile_enable.argtypes = [ctx, C.c_int]
    lib.moloch_b200_profile_read.argtypes = [ctx, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.moloch_b200_launch_count.argtypes = [ctx, C.c_int]
    lib.moloch_b200_launch_count.restype = C.c_int64
    lib.moloch_b200_device_bytes.argtypes = [ctx]
    lib.moloch_b200_device_bytes.restype = C.c_uint64
    lib.moloch_b200_halo_plan.argtypes = [C.POINTER(Config), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p]
    return lib


def make_config(wl, g: Geom, device: int = -1, mo_dzita: float | None = None, bdy: dict | None = None) -> Config:
    """moloch_b200_config from the workload (namelist values), the rank's
    geometry (set_nproc / setup_model_indexes results) and, for a run with
    lateral boundaries, the results of setup_bdycon (`bdy`: nztop, km, lm)."""
    if mo_dzita is None:
        mo_dzita = wl.mo_ztop / float(wl.kz)   # zita(kz), Main/mod_params.F90:2461-2463
    bdy = bdy or {}
    return Config(jx=wl.jx, iy=wl.iy, kz=wl.kz, nqx=wl.nqx, ntr=wl.ntr, iqfrst=2,
                  jde1=g.jde1, jde2=g.jde2, ide1=g.ide1, ide2=g.ide2, jce1=g.jce1, jce2=g.jce2, ice1=g.ice1,
                  ice2=g.ice2, has_bdy_left=int(g.bl), has_bdy_right=int(g.br), has_bdy_bottom=int(g.bb),
                  has_bdy_top=int(g.bt), bandflag=int(g.band), crmflag=int(g.crm), nbr_left=g.left,
                  nbr_right=g.right, nbr_bottom=g.bottom, nbr_top=g.top, rank=g.rank, nranks=g.px * g.py,
                  mo_nadv=wl.mo_nadv, mo_nsound=wl.mo_nsound, mo_divdamp=wl.mo_divdamp,
                  mo_divfilter=wl.mo_divfilter, lrotllr=wl.lrotllr, ipptls=wl.ipptls, device=device,
                  ibltyp=wl.ibltyp, dtsec=wl.dt, dx=wl.dx, mo_dzita=mo_dzita,
                  do_bdy=wl.do_bdy, nspgx=wl.nspgx if wl.do_bdy else 0, present_qc=wl.present_qc,
                  present_qi=wl.present_qi, mo_top_nudge=wl.mo_top_nudge if wl.do_bdy else 0,
                  mo_spectral_nudge=wl.mo_spectral_nudge if wl.do_bdy else 0, nztop=int(bdy.get("nztop", 0)),
                  ichem=int(wl.ntr > 0), ichebdy=wl.ichebdy, do_slice=wl.do_slice, icldmstrat=wl.icldmstrat,
                  km=int(bdy.get("km", 0)), lm=int(bdy.get("lm", 0)), do_massck=int(getattr(wl, "do_massck", 0)), dtbdys=wl.dtbdys,
                  dtrad=wl.dtrad, rhmin=wl.rhmin, rhmax=wl.rhmax, tkemin=wl.tkemin,
                  irceideal=int(getattr(wl, "irceideal", 0)), idiag=int(getattr(wl, "idiag", 0)),
                  ichdiag=int(getattr(wl, "ichdiag", 0)), niycpus=g.py)


def halo_plan(cfg: Config, stag: int, nex: int, lr: bool, bt: bool):
    """Host-only: send/recv boxes of one exchange (no GPU needed)."""
    lib = load_library()
    s = (C.c_int32 * 16)()
    r = (C.c_int32 * 16)()
    if lib.moloch_b200_halo_plan(C.byref(cfg), stag, nex, int(lr), int(bt), s, r):
        raise MolochError(lib.moloch_b200_last_error().decode())
    return np.array(s).reshape(4, 4), np.array(r).reshape(4, 4)


class MolochB200:
    """One rank's MOLOCH dycore on one B200 (mirror of `mod_moloch`)."""

    def __init__(self, wl, rank: int = 0, nranks: int = 1, px: int | None = None, py: int | None = None,
                 device: int = -1, bdy: dict | None = None, lib=None):
        # `lib`: a pre-loaded library object with the same entry points (the CPU
        # tests bind the host-compiled instantiation of the boundary cell functions
        # this way); the product always loads libmoloch_b200.so.
        self.lib = lib if lib is not None else load_library()
        self.wl = wl
        if px is None or py is None:
            px, py = default_cpus_per_dim(nranks, wl.jx, wl.iy)
        if px * py != nranks:
            raise ValueError("px*py != nranks")
        self.g = make_geom(wl.jx, wl.iy, wl.kz, wl.i_band, wl.i_crm, px, py, rank)
        # setup_bdycon results (host side, Main/mod_bdycod.F90:478-568): needed before
        # allocation because nztop/km/lm are part of the configuration
        if wl.do_bdy and bdy is None:
            from . import synthetic as S
            bdy = S.bdycon_setup(wl)
        self.bdy = bdy
        self.cfg = make_config(wl, self.g, device, bdy=bdy)
        self.ctx = C.c_void_p()
        self._pinned = []
        self._registered = []

    # ---- error convention: non-zero -> fatal(__FILE__,__LINE__,msg) ---------
    def _chk(self, rc):
        if rc != 0:
            raise MolochError(self.lib.moloch_b200_last_error().decode())

    # ---- the reference's three entry points -----------------------------------
    def allocate_moloch(self):
        """allocate_moloch (Main/mod_moloch.F90:159) + device copies of mo_atm."""
        self._chk(self.lib.moloch_b200_create(C.byref(self.cfg), C.byref(self.ctx)))
        return self

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._chk(self.lib.moloch_b200_comm_init(self.ctx, buf))

    def p2p_export(self) -> bytes:
        """This rank's blob for the direct peer transport (all-gather them)."""
        n = int(self.lib.moloch_b200_p2p_blob_size())
        buf = C.create_string_buffer(n)
        self._chk(self.lib.moloch_b200_p2p_export(self.ctx, buf))
        return buf.raw

    def p2p_connect(self, blobs: list):
        """blobs: every rank's p2p_export(), in rank order."""
        raw = b"".join(blobs)
        buf = C.create_string_buffer(raw, len(raw))
        self._chk(self.lib.moloch_b200_p2p_connect(self.ctx, buf, len(blobs)))

    @staticmethod
    def comm_id(lib=None) -> bytes:
        lib = lib if lib is not None else load_library()
        buf = C.create_string_buffer(128)
        if lib.moloch_b200_comm_id(buf):
            raise MolochError(lib.moloch_b200_last_error().decode())
        return buf.raw

    def init_moloch(self, fields: dict, profiles: dict, boxes: dict | None = None):
        """init_moloch (:201): hand the static fields, profiles and the initial
        state to the device.  `fields` maps names to GLOBAL arrays (cut here to
        the rank's bounds) or, when `boxes` gives their bounds, to rank-local
        arrays."""
        for name, arr in fields.items():
            if boxes is not None and name in boxes:
                self.set_local(name, arr, boxes[name])
            else:
                self.set_global(name, arr)
        for name, v in profiles.items():
            self.set_profile(name, v)
        self._chk(self.lib.moloch_b200_init(self.ctx))
        if self.wl.do_bdy:
            self.init_boundary()
        return self

    def init_boundary(self):
        """Hand the results of setup_bdycon / setup_boundaries / lowpass_init to
        the device (Main/mod_bdycod.F90:478-568, 3844-3896;
        Main/mod_atm_interface.F90:384-532)."""
        b, g, wl = self.bdy, self.g, self.wl
        if wl.nspgx > 0:
            self.set_table("hefc", b["hefc"])
            for which, ib in b["ibnd"].items():
                self.set_ibnd(which, ib)
            if wl.ntr > 0 and wl.ichebdy != 0:
                self.set_table("fcx", b["fcx"])
        if wl.mo_top_nudge:
            self.set_table("tnudge", b["tnudge"])
        if wl.mo_spectral_nudge:
            self.set_table("cnudge", b["cnudge"])
            self.set_table("bvx", np.ascontiguousarray(b["bvx"][:, g.jde1 - 1:g.jde2]))
            self.set_table("bvy", np.ascontiguousarray(b["bvy"][:, g.ide1 - 1:g.ide2]))

    def set_table(self, name: str, v):
        a = np.ascontiguousarray(v, dtype=np.float64)
        self._chk(self.lib.moloch_b200_set_table(self.ctx, TABLE_ID[name], a.ctypes.data, a.size))

    def set_ibnd(self, which: str, glob: np.ndarray):
        """ba_cr/ba_ud/ba_vd %ibnd: cut the rank's (jde1:jde2, ide1:ide2) box out of the global plane."""
        g = self.g
        a = np.ascontiguousarray(np.asarray(glob)[g.ide1 - 1:g.ide2, g.jde1 - 1:g.jde2], dtype=np.int32)
        self._chk(self.lib.moloch_b200_set_ibnd(self.ctx, IBND_ID[which], a.ctypes.data, g.jde1, g.jde2, g.ide1, g.ide2))

    def load_boundary(self, B: dict):
        """Upload the ICBC b0/b1 buffers (global arrays keyed dub0, dub1, ...)."""
        wl = self.wl
        for name, arr in B.items():
            if (name.startswith("xlb") and not wl.present_qc) or (name.startswith("xib") and not wl.present_qi) \
                    or (name.startswith("chib") and (wl.ntr == 0 or wl.ichebdy == 0)):
                continue        # buffers the configuration does not allocate
            self.set_global(name, arr)

    def boundary(self): self._chk(self.lib.moloch_b200_boundary(self.ctx))
    def bdyval(self): self._chk(self.lib.moloch_b200_bdyval(self.ctx))
    def bdy_shift(self): self._chk(self.lib.moloch_b200_bdy_shift(self.ctx))
    def mkslice(self): self._chk(self.lib.moloch_b200_mkslice(self.ctx))
    def set_calday(self, calday: float, dayspy: float = 365.2422):
        self._chk(self.lib.moloch_b200_set_calday(self.ctx, float(calday), float(dayspy)))
    def massck(self) -> np.ndarray:
        """This rank's tdrym, tdadv, tqmass, tqadv (Main/mod_massck.F90:77-185)."""
        out = np.zeros(4)
        self._chk(self.lib.moloch_b200_massck(self.ctx, out.ctypes.data))
        return out

    def ps_check(self):
        """(max ps, min ps, number of non-finite values) over the interior (Main/mod_moloch.F90:407-422)."""
        mm = np.zeros(2)
        bad = C.c_int32(0)
        self._chk(self.lib.moloch_b200_ps_check(self.ctx, mm.ctypes.data, C.byref(bad)))
        return float(mm[0]), float(mm[1]), int(bad.value)

    def set_xbctime(self, t: float): self._chk(self.lib.moloch_b200_set_xbctime(self.ctx, float(t)))
    def get_xbctime(self) -> float: return float(self.lib.moloch_b200_get_xbctime(self.ctx))

    def moloch(self, nsteps: int = 1):
        """moloch (:312) with the host physics returning zero tendencies."""
        self._chk(self.lib.moloch_b200_step(self.ctx, int(nsteps)))

    # ---- per-subroutine entries (used by the parity tests) --------------------
    def reset_tendencies(self): self._chk(self.lib.moloch_b200_reset_tendencies(self.ctx))
    def sound(self): self._chk(self.lib.moloch_b200_sound(self.ctx))
    def advection(self): self._chk(self.lib.moloch_b200_advection(self.ctx))
    def dynamical_core(self): self._chk(self.lib.moloch_b200_dynamical_core(self.ctx))
    def diagnostics(self): self._chk(self.lib.moloch_b200_diagnostics(self.ctx))
    def status_update(self): self._chk(self.lib.moloch_b200_status_update(self.ctx))

    def wafone(self, field: str, n: int = 1):
        self._chk(self.lib.moloch_b200_wafone(self.ctx, FIELD_ID[field], int(n)))

    def sync(self): self._chk(self.lib.moloch_b200_sync(self.ctx))

    def set_async(self, on: bool):
        """Batch mode for set_local/get_local: transfers are only enqueued; end with sync()."""
        self._chk(self.lib.moloch_b200_set_async(self.ctx, int(on)))

    def xfer_list(self, items):
        """items: (name, species_or_0, array, box) -> a ctypes array of moloch_b200_xfer.  `array` is the
        host array of one species/field, (nk, ni, nj) C-contiguous float64 with Fortran bounds `box`."""
        arr = (Xfer * max(len(items), 1))()
        for q, (name, n, a, box) in enumerate(items):
            if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"{name}: the hand-off needs C-contiguous float64 host arrays")
            nk = self._levels(name)
            jlo, jhi, ilo, ihi = box
            if a.size != nk * (ihi - ilo + 1) * (jhi - jlo + 1):
                raise ValueError(f"{name}: shape {a.shape} does not match bounds {box} x {nk} levels")
            arr[q] = Xfer(FIELD_ID[name], int(n), a.ctypes.data, jlo, jhi, ilo, ihi, 1, nk)
        return arr, len(items)

    def handoff(self, down, up, nslabs: int = 8, physics=None):
        """The pipelined physics hand-off (moloch_b200_handoff): `down`/`up` are xfer_list() results;
        `physics(i1, i2)` is called on the host for each slab of rows once its state has arrived."""
        cb = None
        raised = []
        if physics is not None:
            def _cb(user, i1, i2):
                # ctypes prints and swallows an exception raised inside a callback and returns 0 to C:
                # keep it, report failure to the library (which stops before uploading that slab's
                # tendencies) and re-raise once the C call has returned
                try:
                    return int(physics(int(i1), int(i2)) or 0)
                except BaseException as exc:  # noqa: BLE001
                    raised.append(exc)
                    return 1
            cb = PHYSICS_FN(_cb)
        rc = self.lib.moloch_b200_handoff(self.ctx, down[0], down[1], up[0], up[1], int(nslabs),
                                          C.cast(cb, C.c_void_p) if cb is not None else None, None)
        if raised:
            raise raised[0]
        self._chk(rc)

    def set_option(self, name: str, value: int):
        """Kernel-variant switch ("wsolve", "waf", "fuse_halo"); all variants are bit-identical."""
        self._chk(self.lib.moloch_b200_set_option(self.ctx, name.encode(), int(value)))

    def set_stream(self, cuda_stream: int):
        self._chk(self.lib.moloch_b200_set_stream(self.ctx, C.c_void_p(cuda_stream)))

    # ---- host <-> device -------------------------------------------------------
    def _levels(self, name):
        lv = H.ALLOC[name][3]
        return self.wl.kz if lv == "kz" else self.wl.kz + 1 if lv == "kzp1" else 1

    def set_local(self, name: str, arr: np.ndarray, box, n: int = 0):
        """arr: (nk, ni, nj) [or (nspec, nk, ni, nj)] with Fortran bounds box."""
        a = np.ascontiguousarray(arr, dtype=np.float64)
        jlo, jhi, ilo, ihi = box
        nk = self._levels(name)
        if name in SPECIES and n == 0:
            for s in range(a.shape[0]):
                self.set_local(name, a[s], box, s + 1)
            return
        if a.size != nk * (ihi - ilo + 1) * (jhi - jlo + 1):
            raise ValueError(f"{name}: shape {a.shape} does not match bounds {box} x {nk} levels")
        self._chk(self.lib.moloch_b200_set_field(self.ctx, FIELD_ID[name], n, a.ctypes.data, jlo, jhi, ilo, ihi,
                                                 1, nk))

    def get_local(self, name: str, box=None, n: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        if box is None:
            box = H.bounds(self.g, name)
        jlo, jhi, ilo, ihi = box
        nk = self._levels(name)
        if name in SPECIES and n == 0:
            nspec = self.wl.nqx if name in ("qx", "qxten") else self.wl.ntr
            return np.stack([self.get_local(name, box, s + 1) for s in range(nspec)])
        if out is None:
            out = np.zeros((nk, ihi - ilo + 1, jhi - jlo + 1))
        self._chk(self.lib.moloch_b200_get_field(self.ctx, FIELD_ID[name], n, out.ctypes.data, jlo, jhi, ilo,
                                                 ihi, 1, nk))
        return out if nk > 1 else out.reshape(ihi - ilo + 1, jhi - jlo + 1)

    def set_global(self, name: str, glob: np.ndarray):
        """Cut the rank's array (reference bounds + ghosts) out of a global one."""
        box = H.bounds(self.g, name)
        self.set_local(name, H.cut(np.asarray(glob), self.g, box), box)

    def get_into_global(self, name: str, glob: np.ndarray):
        """Write this rank's owned cells of `name` into a global array."""
        own = H.owned(self.g, name)
        loc = self.get_local(name, own)
        H.paste(glob, loc, own, own)
        return glob

    def global_shape(self, name: str):
        nk = self._levels(name)
        shp = (self.wl.iy, self.wl.jx) if nk == 1 else (nk, self.wl.iy, self.wl.jx)
        if name in ("qx", "qxten"):
            shp = (self.wl.nqx,) + shp
        if name in ("trac", "chiten", "chib0", "chib1", "chiten0", "cadvhdiag", "cbdydiag"):
            shp = (self.wl.ntr,) + shp
        return shp

    def get_global(self, name: str) -> np.ndarray:
        """Single-rank convenience: the owned cells on the global grid."""
        return self.get_into_global(name, np.zeros(self.global_shape(name)))

    def set_profile(self, name: str, v):
        a = np.ascontiguousarray(v, dtype=np.float64)
        if name == "rlat":   # rlat(ide1 : ide2+1)
            a = np.ascontiguousarray(a[self.g.ide1 - 1:self.g.ide2 + 1])
        self._chk(self.lib.moloch_b200_set_profile(self.ctx, PROFILE_ID[name], a.ctypes.data, a.size))

    # ---- pinned host buffers ----------------------------------------------------
    def pinned_empty(self, shape) -> np.ndarray:
        n = int(np.prod(shape))
        p = C.c_void_p()
        self._chk(self.lib.moloch_b200_host_alloc(C.byref(p), n * 8))
        self._pinned.append(p)
        buf = (C.c_double * n).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64).reshape(shape)

    def host_register(self, a: np.ndarray):
        """Page-lock a host array the caller owns (until host_unregister / close)."""
        self._chk(self.lib.moloch_b200_host_register(C.c_void_p(a.ctypes.data), a.nbytes))
        self._registered.append(a)

    # ---- instrumentation ---------------------------------------------------------
    def profile_enable(self, on: bool = True):
        self._chk(self.lib.moloch_b200_profile_enable(self.ctx, int(on)))

    def profile_read(self) -> dict:
        cap = 64
        names = ((C.c_char * 48) * cap)()
        ms = (C.c_double * cap)()
        cnt = (C.c_int64 * cap)()
        n = self.lib.moloch_b200_profile_read(self.ctx, cap, names, ms, cnt)
        if n < 0:
            raise MolochError(self.lib.moloch_b200_last_error().decode())
        return {names[q].value.decode(): {"ms": ms[q], "launches": cnt[q]} for q in range(min(n, cap))}

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.moloch_b200_launch_count(self.ctx, int(reset)))

    def device_bytes(self) -> int:
        return int(self.lib.moloch_b200_device_bytes(self.ctx))

    def close(self):
        if self.ctx:
            self.lib.moloch_b200_destroy(self.ctx)
            self.ctx = C.c_void_p()
        for p in self._pinned:
            self.lib.moloch_b200_host_free(p)
        self._pinned = []
        for a in self._registered:
            self.lib.moloch_b200_host_unregister(C.c_void_p(a.ctypes.data))
        self._registered = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


STATIC_FIELDS = ["fmz", "fmzf", "rfmzu", "rfmzv", "zeta", "hx", "hy", "msfx", "msfu", "msfv", "coru", "corv",
                 "bdywtu", "bdywtv", "bdywtw"]
STATE_FIELDS = ["u", "v", "w", "pai", "tetav", "t", "qx", "trac", "ux", "vx", "tvirt", "p", "rho", "qsat", "ps"]
PROFILE_NAMES = ["gzitak", "gzitakh", "ffilt", "xkdamp", "xknu"]
