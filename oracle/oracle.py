"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see moloch_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (regcm_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OracleConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "jx", "iy", "kz", "nqx", "ntr", "i_band", "i_crm", "px", "py", "mo_nadv", "mo_nsound",
        "mo_divdamp", "mo_divfilter", "lrotllr", "ipptls", "nspgx")] + [
        (n, C.c_double) for n in ("dtsec", "dx", "mo_ztop", "mo_h", "mo_a0")]


class OracleExtConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "do_bdy", "present_qc", "present_qi", "mo_top_nudge", "mo_spectral_nudge", "ichem", "ichebdy",
        "ibltyp", "icldmstrat", "do_slice", "idiag", "irceideal")] + [
        (n, C.c_double) for n in ("dtbdys", "dtrad", "rhmin", "rhmax", "tkemin", "calday", "dayspy")] + [
        (n, C.c_int) for n in ("ichdiag", "reserved")]


def build(force: bool = False) -> None:
    """Compile the oracle shared libraries (g++, a few seconds)."""
    if force or not (os.path.exists(os.path.join(_HERE, "libmoloch_oracle.so"))
                     and os.path.exists(os.path.join(_HERE, "libmoloch_oracle_chk.so"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


_libs = {}


def _lib(checked: bool):
    key = "chk" if checked else "fast"
    if key not in _libs:
        build()
        name = "libmoloch_oracle_chk.so" if checked else "libmoloch_oracle.so"
        lib = C.CDLL(os.path.join(_HERE, name))
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.POINTER(OracleConfig)]
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_last_error.restype = C.c_char_p
        lib.oracle_global_size.restype = C.c_long
        lib.oracle_global_size.argtypes = [C.c_void_p, C.c_char_p]
        lib.oracle_set_global.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        lib.oracle_get_global.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        for f in ("oracle_setup_static", "oracle_init_state", "oracle_reset_tendencies", "oracle_sound",
                  "oracle_advection", "oracle_dynamical_core", "oracle_diagnostics", "oracle_status_update"):
            getattr(lib, f).argtypes = [C.c_void_p]
        for f in ("oracle_boundary", "oracle_bdyval", "oracle_mkslice"):
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.oracle_set_ext.argtypes = [C.c_void_p, C.POINTER(OracleExtConfig)]
        lib.oracle_massck.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_ps_check.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_set_xbctime.argtypes = [C.c_void_p, C.c_double]
        lib.oracle_get_xbctime.argtypes = [C.c_void_p]
        lib.oracle_get_xbctime.restype = C.c_double
        lib.oracle_get_int.argtypes = [C.c_void_p, C.c_char_p]
        lib.oracle_step.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_wafone.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _libs[key] = lib
    return _libs[key]


class Oracle:
    """One oracle world: the global domain split into px x py subdomains."""

    def __init__(self, wl, px: int = 1, py: int = 1, checked: bool = False):
        self.lib = _lib(checked)
        self.wl = wl
        self.cfg = OracleConfig(jx=wl.jx, iy=wl.iy, kz=wl.kz, nqx=wl.nqx, ntr=wl.ntr, i_band=wl.i_band,
                                i_crm=wl.i_crm, px=px, py=py, mo_nadv=wl.mo_nadv, mo_nsound=wl.mo_nsound,
                                mo_divdamp=wl.mo_divdamp, mo_divfilter=wl.mo_divfilter, lrotllr=wl.lrotllr,
                                ipptls=wl.ipptls, nspgx=wl.nspgx, dtsec=wl.dt, dx=wl.dx, mo_ztop=wl.mo_ztop,
                                mo_h=wl.mo_h, mo_a0=wl.mo_a0)
        self.h = self.lib.oracle_create(C.byref(self.cfg))
        if not self.h:
            raise RuntimeError(self.lib.oracle_last_error().decode())
        self.ext = None
        if getattr(wl, "needs_ext", False):
            self.ext = OracleExtConfig(do_bdy=wl.do_bdy, present_qc=wl.present_qc, present_qi=wl.present_qi,
                                       mo_top_nudge=wl.mo_top_nudge, mo_spectral_nudge=wl.mo_spectral_nudge,
                                       ichem=int(wl.ntr > 0), ichebdy=wl.ichebdy, ibltyp=wl.ibltyp,
                                       icldmstrat=wl.icldmstrat, do_slice=wl.do_slice, idiag=wl.idiag, ichdiag=wl.ichdiag, reserved=0,
                                       irceideal=wl.irceideal, dtbdys=wl.dtbdys, dtrad=wl.dtrad, rhmin=wl.rhmin,
                                       rhmax=wl.rhmax, tkemin=wl.tkemin, calday=wl.calday, dayspy=wl.dayspy)
            self._chk(self.lib.oracle_set_ext(self.h, C.byref(self.ext)))

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.oracle_last_error().decode())

    def set(self, name: str, arr) -> None:
        a = np.ascontiguousarray(arr, dtype=np.float64)
        n = self.lib.oracle_global_size(self.h, name.encode())
        if n != a.size:
            raise ValueError(f"{name}: expected {n} values, got {a.size}")
        self._chk(self.lib.oracle_set_global(self.h, name.encode(), a.ctypes.data))

    def get(self, name: str) -> np.ndarray:
        wl = self.wl
        n = self.lib.oracle_global_size(self.h, name.encode())
        if n <= 0:
            raise KeyError(name)
        out = np.zeros(n)
        self._chk(self.lib.oracle_get_global(self.h, name.encode(), out.ctypes.data))
        plane = wl.jx * wl.iy
        if n % plane == 0 and n >= plane:
            nk = n // plane
            if name in ("qx", "qxten"):
                return out.reshape(wl.nqx, wl.kz, wl.iy, wl.jx)
            if name in ("trac", "chiten", "chib0", "chib1", "chiten0", "cadvhdiag", "cbdydiag"):
                return out.reshape(wl.ntr, wl.kz, wl.iy, wl.jx)
            return out.reshape(wl.iy, wl.jx) if nk == 1 else out.reshape(nk, wl.iy, wl.jx)
        return out

    def load_primary(self, P: dict) -> None:
        """Feed the file-like inputs of regcm_b200.synthetic.make_primary, then run
        the oracle's own restatement of the set-up code."""
        for k in ("ht", "htu", "htv", "msfx", "msfu", "msfv", "ulat", "vlat", "rlat", "ps", "t", "qx", "u", "v"):
            self.set(k, P[k])
        if self.ext is not None and "xlat" in P:
            self.set("xlat", P["xlat"])
        if "hefc" in P and self.wl.nspgx > 0:
            self.set("hefc", P["hefc"])
        if "trac" in P and self.wl.ntr > 0:
            self.set("trac", P["trac"])
        if "fcx" in P and self.wl.nspgx > 0 and self.ext is not None:
            self.set("fcx", P["fcx"])
        self._chk(self.lib.oracle_setup_static(self.h))
        self._chk(self.lib.oracle_init_state(self.h))
        if self.ext is not None and self.wl.ibltyp == 2 and "tke" in P:
            self.set("tke", P["tke"])

    def load_boundary(self, B: dict) -> None:
        """b0/b1 buffers of regcm_b200.synthetic.make_boundary (ICBC stand-in)."""
        for k, v in B.items():
            self.set(k, v)

    def step(self, n: int = 1): self._chk(self.lib.oracle_step(self.h, n))
    def reset_tendencies(self): self._chk(self.lib.oracle_reset_tendencies(self.h))
    def sound(self): self._chk(self.lib.oracle_sound(self.h))
    def advection(self): self._chk(self.lib.oracle_advection(self.h))
    def dynamical_core(self): self._chk(self.lib.oracle_dynamical_core(self.h))
    def diagnostics(self): self._chk(self.lib.oracle_diagnostics(self.h))
    def status_update(self): self._chk(self.lib.oracle_status_update(self.h))
    def boundary(self): self._chk(self.lib.oracle_boundary(self.h))
    def bdyval(self): self._chk(self.lib.oracle_bdyval(self.h))
    def mkslice(self): self._chk(self.lib.oracle_mkslice(self.h))
    def massck(self) -> np.ndarray:
        out = np.zeros(4)
        self._chk(self.lib.oracle_massck(self.h, out.ctypes.data))
        return out

    def ps_check(self):
        mm = np.zeros(2)
        bad = C.c_int(0)
        self._chk(self.lib.oracle_ps_check(self.h, mm.ctypes.data, C.byref(bad)))
        return float(mm[0]), float(mm[1]), int(bad.value)

    def set_xbctime(self, t: float): self._chk(self.lib.oracle_set_xbctime(self.h, float(t)))
    def get_xbctime(self) -> float: return float(self.lib.oracle_get_xbctime(self.h))
    def get_int(self, name: str) -> int: return int(self.lib.oracle_get_int(self.h, name.encode()))

    def wafone(self, field: str, n: int = 1): self._chk(self.lib.oracle_wafone(self.h, field.encode(), n))

    def set_threads(self, n: int): self.lib.oracle_set_threads(int(n))
    def get_threads(self) -> int: return int(self.lib.oracle_get_threads())
