"""Host-side domain decomposition and index ranges.

In a RegCM build these values come from the unchanged Fortran host code
(`set_nproc`, Main/mpplib/mod_mppparam.F90:1250-1641, and
`setup_model_indexes`, Main/mod_atm_interface.F90:182-382) and are handed to
the C ABI in `moloch_b200_config`.  This module is the Python stand-in for that
host code so that tests and benchmarks can drive the C ABI without Fortran.
"""
from __future__ import annotations

import math
from dataclasses import dataclass


def default_cpus_per_dim(nproc: int, jx: int, iy: int) -> tuple[int, int]:
    """Process grid (njxcpus, niycpus) chosen by set_nproc (:1381-1401)."""
    if nproc < 4:
        return nproc, 1
    c1 = (int(round(math.sqrt(float(nproc)))) // 2) * 2
    if iy > int(1.5 * jx):
        c1 -= 1
        while nproc % c1 != 0:
            c1 -= 1
    elif jx > int(1.5 * iy):
        c1 += 1
        while nproc % c1 != 0:
            c1 += 1
    else:
        while nproc % c1 != 0:
            c1 += 1
    return c1, nproc // c1


@dataclass
class Geom:
    jx: int
    iy: int
    kz: int
    band: bool
    crm: bool
    px: int
    py: int
    rank: int
    locj: int = 0
    loci: int = 0
    left: int = -1
    right: int = -1
    bottom: int = -1
    top: int = -1
    bl: bool = False
    br: bool = False
    bb: bool = False
    bt: bool = False
    jde1: int = 0
    jde2: int = 0
    ide1: int = 0
    ide2: int = 0
    jce1: int = 0
    jce2: int = 0
    ice1: int = 0
    ice2: int = 0

    # ghost multipliers ma%jbl1.. (1 where a neighbour exists)
    @property
    def gl(self): return 0 if self.bl else 1
    @property
    def gr(self): return 0 if self.br else 1
    @property
    def gb(self): return 0 if self.bb else 1
    @property
    def gt(self): return 0 if self.bt else 1
    # internal ranges
    @property
    def jci1(self): return self.jce1 + (1 if self.bl else 0)
    @property
    def jci2(self): return self.jce2 - (1 if self.br else 0)
    @property
    def ici1(self): return self.ice1 + (1 if self.bb else 0)
    @property
    def ici2(self): return self.ice2 - (1 if self.bt else 0)
    @property
    def jdi1(self): return self.jde1 + (1 if self.bl else 0)
    @property
    def jdi2(self): return self.jde2 - (1 if self.br else 0)
    @property
    def idi1(self): return self.ide1 + (1 if self.bb else 0)
    @property
    def idi2(self): return self.ide2 - (1 if self.bt else 0)

    def ext(self, stag: str, ghost_j: int = 0, ghost_i: int = 0):
        """(jlo, jhi, ilo, ihi) of the owned box of a staggering, widened by
        ghost_j/ghost_i points on sides that have a neighbour (the reference's
        `ga`/`gb` suffixes)."""
        if stag == "cross":
            b = (self.jce1, self.jce2, self.ice1, self.ice2)
        elif stag == "u":
            b = (self.jde1, self.jde2, self.ice1, self.ice2)
        elif stag == "v":
            b = (self.jce1, self.jce2, self.ide1, self.ide2)
        elif stag == "dot":
            b = (self.jde1, self.jde2, self.ide1, self.ide2)
        else:
            raise ValueError(stag)
        return (b[0] - ghost_j * self.gl, b[1] + ghost_j * self.gr,
                b[2] - ghost_i * self.gb, b[3] + ghost_i * self.gt)


def make_geom(jx: int, iy: int, kz: int, i_band: int, i_crm: int, px: int, py: int, rank: int) -> Geom:
    band = (i_band == 1 or i_crm == 1)
    crm = (i_crm == 1)
    g = Geom(jx, iy, kz, band, crm, px, py, rank)
    nproc = px * py
    if nproc == 1:
        gdj1, gdi1, gdj2, gdi2 = 1, 1, jx, iy
        gcj2 = jx if band else jx - 1
        gci2 = iy if crm else iy - 1
        if crm:
            g.left = g.right = g.top = g.bottom = 0
        else:
            g.bt = g.bb = True
            if band:
                g.left = g.right = 0
            else:
                g.bl = g.br = True
    else:
        g.locj, g.loci = rank // py, rank % py

        def cart(lj, li):
            if lj < 0 or lj >= px:
                if not band:
                    return -1
                lj %= px
            if li < 0 or li >= py:
                if not crm:
                    return -1
                li %= py
            return lj * py + li
        g.left, g.right = cart(g.locj - 1, g.loci), cart(g.locj + 1, g.loci)
        g.bottom, g.top = cart(g.locj, g.loci - 1), cart(g.locj, g.loci + 1)
        g.bt, g.bb, g.br, g.bl = g.top < 0, g.bottom < 0, g.right < 0, g.left < 0
        jxp, iyp = jx // px, iy // py
        gdj1, gdi1 = g.locj * jxp + 1, g.loci * iyp + 1
        if jxp * px < jx:
            imiss = jx - jxp * px
            if g.locj < imiss:
                gdj1 += g.locj
                jxp += 1
            else:
                gdj1 += imiss
        if iyp * py < iy:
            imiss = iy - iyp * py
            if g.loci < imiss:
                gdi1 += g.loci
                iyp += 1
            else:
                gdi1 += imiss
        gdj2, gdi2 = gdj1 + jxp - 1, gdi1 + iyp - 1
        if jxp < 3 or iyp < 3:
            raise ValueError("Cannot have one processor with less than 3x3 points")
        gci2 = gdi2 - 1 if (not crm and gdi2 == iy) else gdi2
        gcj2 = gdj2 - 1 if (not band and gdj2 == jx) else gdj2
    g.jde1, g.jde2, g.ide1, g.ide2 = gdj1, gdj2, gdi1, gdi2
    g.jce1, g.jce2, g.ice1, g.ice2 = gdj1, gcj2, gdi1, gci2
    return g
