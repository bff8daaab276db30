"""Per-rank host arrays with RegCM's bounds and ghost widths.

In a RegCM build the host arrays are the Fortran module variables of
`mod_atm_interface` (allocation: Main/mod_atm_interface.F90:579-624, 844-848)
and of `mod_moloch` (Main/mod_moloch.F90:159-199).  This module is the Python
stand-in: it cuts a rank's arrays out of global (nk, iy, jx) arrays, with
periodic wrap where the reference's communicator is periodic, and assembles
global arrays back from the ranks' owned boxes.
"""
from __future__ import annotations

import numpy as np

from .decomp import Geom

# field -> (staggering, ghost_j, ghost_i, levels)  levels: "kz" | "kzp1" | 1
ALLOC = {
    "u": ("u", 2, 1, "kz"), "v": ("v", 1, 2, "kz"),
    "ux": ("cross", 2, 1, "kz"), "vx": ("cross", 1, 2, "kz"),
    "w": ("cross", 0, 0, "kzp1"), "pai": ("cross", 1, 1, "kz"), "tetav": ("cross", 1, 1, "kz"),
    "t": ("cross", 1, 1, "kz"), "qx": ("cross", 1, 1, "kz"), "trac": ("cross", 1, 1, "kz"),
    "zeta": ("cross", 2, 2, "kz"), "tvirt": ("cross", 0, 0, "kz"), "p": ("cross", 0, 0, "kz"),
    "rho": ("cross", 0, 0, "kz"), "qsat": ("cross", 0, 0, "kz"), "ps": ("cross", 0, 0, 1),
    "fmz": ("cross", 1, 1, "kz"), "fmzf": ("cross", 0, 0, "kzp1"),
    "rfmzu": ("u", 1, 1, "kz"), "rfmzv": ("v", 1, 1, "kz"),
    "hx": ("u", 1, 0, 1), "hy": ("v", 0, 1, 1),
    "msfx": ("dot", 1, 1, 1), "msfu": ("dot", 1, 1, 1), "msfv": ("dot", 1, 1, 1),
    "coru": ("u", 0, 0, 1), "corv": ("v", 0, 0, 1),
    "bdywtu": ("u", 0, 0, "kz"), "bdywtv": ("v", 0, 0, "kz"), "bdywtw": ("cross", 0, 0, "kz"),
    "tten": ("cross", 0, 0, "kz"), "uten": ("cross", 0, 0, "kz"), "vten": ("cross", 0, 0, "kz"),
    "qxten": ("cross", 0, 0, "kz"), "chiten": ("cross", 0, 0, "kz"),
    "s": ("cross", 0, 0, "kzp1"), "zdiv2": ("cross", 1, 1, "kz"), "wx": ("cross", 1, 1, "kz"),
    "wz": ("cross", 2, 2, "kz"), "p0": ("cross", 2, 2, "kz"), "tetavf": ("cross", 0, 0, "kz"),
    # UW-PBL TKE (Main/mod_atm_interface.F90:609-612, Main/mod_moloch.F90:184)
    "tke": ("cross", 0, 0, "kzp1"), "tketen": ("cross", 0, 0, "kzp1"), "tkex": ("cross", 0, 0, "kz"),
    # v3dbound/v2dbound b0, b1 (Main/mod_atm_interface.F90:547-577)
    "dub0": ("u", 2, 2, "kz"), "dub1": ("u", 2, 2, "kz"), "dvb0": ("v", 2, 2, "kz"), "dvb1": ("v", 2, 2, "kz"),
    "xtb0": ("cross", 1, 1, "kz"), "xtb1": ("cross", 1, 1, "kz"),
    "xpaib0": ("cross", 1, 1, "kz"), "xpaib1": ("cross", 1, 1, "kz"),
    "xqb0": ("cross", 1, 1, "kz"), "xqb1": ("cross", 1, 1, "kz"),
    "xlb0": ("cross", 1, 1, "kz"), "xlb1": ("cross", 1, 1, "kz"),
    "xib0": ("cross", 1, 1, "kz"), "xib1": ("cross", 1, 1, "kz"),
    "xpsb0": ("cross", 1, 1, 1), "xpsb1": ("cross", 1, 1, 1),
    "chib0": ("cross", 1, 1, "kz"), "chib1": ("cross", 1, 1, "kz"),
    # mkslice outputs (Main/mod_atm_interface.F90:599,964-1012) and zetaf (:602)
    "pf3d": ("cross", 0, 0, "kzp1"), "th3d": ("cross", 0, 0, "kz"), "rhb3d": ("cross", 0, 0, "kz"),
    "wpx3d": ("cross", 0, 0, "kz"), "rhox2d": ("cross", 0, 0, 1), "tp2d": ("cross", 0, 0, 1),
    "th700": ("cross", 0, 0, 1), "zetaf": ("cross", 0, 0, "kzp1"),
    "xlat": ("cross", 0, 0, 1), "ptrop": ("cross", 0, 0, 1), "ktrop": ("cross", 0, 0, 1),
    "kmxpbl": ("cross", 0, 0, 1),
    # tendency diagnostics (Main/mod_moloch.F90:187-192; tdiag/qdiag, chemistry diagnostics)
    "ten0": ("cross", 0, 0, "kz"), "qen0": ("cross", 0, 0, "kz"), "tdiag_adh": ("cross", 0, 0, "kz"),
    "qdiag_adh": ("cross", 0, 0, "kz"), "tdiag_bdy": ("cross", 0, 0, "kz"), "qdiag_bdy": ("cross", 0, 0, "kz"),
    "chiten0": ("cross", 0, 0, "kz"), "cadvhdiag": ("cross", 0, 0, "kz"), "cbdydiag": ("cross", 0, 0, "kz"),
}


def bounds(g: Geom, name: str):
    """(jlo, jhi, ilo, ihi) of the rank's host array `name` (owned + ghosts)."""
    stag, gj, gi, _ = ALLOC[name]
    return g.ext(stag, gj, gi)


def owned(g: Geom, name: str):
    stag = ALLOC[name][0]
    return g.ext(stag, 0, 0)


def cut(glob: np.ndarray, g: Geom, box) -> np.ndarray:
    """Cut box (jlo,jhi,ilo,ihi) (global 1-based indices, inclusive) out of a
    global array whose last two axes are (iy, jx); indices outside the grid
    wrap (they only occur in periodic directions)."""
    jlo, jhi, ilo, ihi = box
    jv = (np.arange(jlo, jhi + 1) - 1) % g.jx
    iv = (np.arange(ilo, ihi + 1) - 1) % g.iy
    return np.ascontiguousarray(glob[..., iv[:, None], jv[None, :]])


def paste(glob: np.ndarray, local: np.ndarray, local_box, own_box) -> None:
    """Write the owned part of a rank's array into the global array."""
    jlo, jhi, ilo, ihi = local_box
    j1, j2, i1, i2 = own_box
    glob[..., i1 - 1:i2, j1 - 1:j2] = local[..., i1 - ilo:i2 - ilo + 1, j1 - jlo:j2 - jlo + 1]
