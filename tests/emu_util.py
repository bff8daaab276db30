"""TEST INFRASTRUCTURE: the host-compiled instantiation of the boundary / mkslice
/ TKE cell functions (tests/emu/emu_bdy.cpp) behind the MolochB200 interface, so
that the CPU tests can drive the device code's per-cell arithmetic through the
same Python calls the GPU tests use.  See the header of emu_bdy.cpp."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from regcm_b200.moloch import Config, MolochB200

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu_bdy.cpp")
LIB = os.path.join(HERE, "emu", "libemu_bdy.so")
DEPS = [SRC, os.path.join(ROOT, "regcm_b200", "csrc", "bdy_cells.h"), os.path.join(ROOT, "regcm_b200", "csrc", "geo.h"),
        os.path.join(ROOT, "include", "moloch_b200.h")]


def build():
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared",
                           "-Wall", "-Wextra", "-I", os.path.join(ROOT, "regcm_b200", "csrc"),
                           "-I", os.path.join(ROOT, "include"), "-o", LIB, SRC])
    return LIB


class _Proxy:
    """Maps moloch_b200_<name> onto emu_b200_<name>."""

    def __init__(self, lib):
        self._lib = lib
        ctx = C.c_void_p
        lib.emu_b200_last_error.restype = C.c_char_p
        lib.emu_b200_create.argtypes = [C.POINTER(Config), C.POINTER(ctx)]
        xf = [ctx, C.c_int, C.c_int, C.c_void_p] + [C.c_int] * 6
        lib.emu_b200_set_field.argtypes = xf
        lib.emu_b200_get_field.argtypes = xf
        lib.emu_b200_set_profile.argtypes = [ctx, C.c_int, C.c_void_p, C.c_int]
        lib.emu_b200_set_table.argtypes = [ctx, C.c_int, C.c_void_p, C.c_int]
        lib.emu_b200_set_ibnd.argtypes = [ctx, C.c_int, C.c_void_p] + [C.c_int] * 4
        lib.emu_b200_set_xbctime.argtypes = [ctx, C.c_double]
        lib.emu_b200_get_xbctime.argtypes = [ctx]
        lib.emu_b200_get_xbctime.restype = C.c_double
        lib.emu_b200_set_order.argtypes = [ctx, C.c_int]
        lib.emu_b200_set_calday.argtypes = [ctx, C.c_double, C.c_double]
        lib.emu_b200_massck.argtypes = [ctx, C.c_void_p]
        lib.emu_b200_ps_check.argtypes = [ctx, C.c_void_p, C.c_void_p]
        for f in ("destroy", "init", "bdyval", "boundary", "boundary_pre", "boundary_post", "mkslice", "tke_destagger",
                  "tke_restagger", "tke_update"):
            getattr(lib, "emu_b200_" + f).argtypes = [ctx]

    def __getattr__(self, name):
        if name.startswith("moloch_b200_"):
            return getattr(self._lib, "emu_b200_" + name[len("moloch_b200_"):])
        raise AttributeError(name)


class EmuMoloch(MolochB200):
    def __init__(self, wl, bdy=None, order=0, rank=0, nranks=1, px=None, py=None):
        lib = C.CDLL(build())
        super().__init__(wl, rank=rank, nranks=nranks, px=px, py=py, bdy=bdy, lib=_Proxy(lib))
        self._emu = lib
        self._order = order

    def allocate_moloch(self):
        super().allocate_moloch()
        self._emu.emu_b200_set_order(self.ctx, self._order)
        return self

    def boundary_pre(self): self._chk(self._emu.emu_b200_boundary_pre(self.ctx))
    def boundary_post(self): self._chk(self._emu.emu_b200_boundary_post(self.ctx))

    def tke_destagger(self): self._chk(self._emu.emu_b200_tke_destagger(self.ctx))
    def tke_restagger(self): self._chk(self._emu.emu_b200_tke_restagger(self.ctx))
    def tke_update(self): self._chk(self._emu.emu_b200_tke_update(self.ctx))

    def close(self):
        if self.ctx:
            self._emu.emu_b200_destroy(self.ctx)
            self.ctx = C.c_void_p()
