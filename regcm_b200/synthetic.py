"""Synthetic RegCM-side model state for the MOLOCH dycore (host stand-in).

A real deployment keeps RegCM's Fortran host code: the DOMAIN/ICBC readers fill
`mddom`/`mo_atm`, `compute_moloch_static` (Main/mod_params.F90:3316-3395)
derives the metric terms and `init` (Main/mod_init.F90:156-221,941-1026) the
initial Exner function.  There is no Fortran toolchain (and no NetCDF input)
in this environment, so this module generates the same arrays analytically, in
NumPy, on the *global* grid; `regcm_b200.hostmodel` then cuts them into the
per-rank arrays (with RegCM's bounds and ghost widths) that cross the C ABI.

Global arrays are C-ordered (nk, iy, jx) / (iy, jx): RegCM's (j,i,k) with j
fastest.  Index [i-1, j-1] is RegCM's (j,i).

The named workloads are the BASELINE.json configs.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import math

import numpy as np

# ---- Share/mod_constants.F90 (non-RCEMIP) ---------------------------------
egrav = 9.80665
rgas = ((6.02214076e23 * 1.3806490e-23) / 28.96454) * 1000.0
cpd = 3.5 * rgas
cvd = 2.5 * rgas
regrav = 1.0 / egrav
rovcp = rgas * (1.0 / cpd)
rdrcv = rgas / cvd
cpovr = cpd / rgas
govcp = egrav / cpd
govr = egrav / rgas
p00 = 1.0e5
lrate = 0.00649
stdt = 288.15
stdp = 1.013250e5
tzero = 273.15
ep1 = 28.96454 / 18.01528 - 1.0
ep2 = 18.01528 / 28.96454
mathpi = 3.14159265358979323846
degrad = mathpi / 180.0
raddeg = 180.0 / mathpi
eomeg2 = 2.0 * 7.2921159e-5
earthrad = 6.371229e6
mo_zfilt_fac = 0.8


@dataclass(frozen=True)
class Workload:
    name: str
    jx: int
    iy: int
    kz: int
    nqx: int = 5
    ntr: int = 0
    i_band: int = 0
    i_crm: int = 0
    ds_km: float = 2.0
    dt: float = 30.0
    clat: float = 0.0
    oro: str = "flat"        # flat | gauss | sine
    oro_h: float = 1500.0
    msf_amp: float = 0.0     # map factor = 1 + amp*sin*cos
    nspgx: int = 0           # sponge width (0: no sponge masks)
    lrotllr: int = 0
    ipptls: int = 2
    mo_nadv: int = 2
    mo_nsound: int = 5
    mo_divdamp: int = 1
    mo_divfilter: int = 1
    mo_ztop: float = 30000.0
    mo_h: float = 8000.0
    mo_a0: float = 0.0
    u0: float = 10.0
    v0: float = 2.0
    seed: int = 20240613
    # lateral boundary / physics hand-off / TKE
    do_bdy: int = 0          # do_apply_bdy (Main/mod_moloch.F90:305,341)
    present_qc: int = 0      # ICBC carries qc / qi (Main/mod_bdycod.F90:695,699)
    present_qi: int = 0
    mo_top_nudge: int = 0    # Share/mod_dynparam.F90:209
    mo_spectral_nudge: int = 0   # :207
    ichebdy: int = 1         # tracer boundary (Main/mod_params.F90:592): 1 chib0/chib1, 0 flux dependent
    ibltyp: int = 1          # 2: UW PBL, TKE advected by the dycore
    icldmstrat: int = 0
    do_slice: int = 0        # mkslice inside moloch()
    do_massck: int = 0       # keep zq on the device for massck (debug_level > 0)
    irceideal: int = 0       # 1: mkslice keeps ptrop (Main/mod_slice.F90:345)
    idiag: int = 0           # > 0: tendency diagnostics tdiag%adh/bdy, qdiag%adh/bdy
    ichdiag: int = 0         # > 0: tracer diagnostics cadvhdiag, cbdydiag
    calday: float = 172.25   # calendar day used by mkslice's tropopause pressure
    dayspy: float = 365.2422
    dtbdys: float = 21600.0
    dtrad: float = 1800.0
    rhmin: float = 0.01      # Main/mod_params.F90:381-382
    rhmax: float = 1.01
    tkemin: float = 1.0e-8

    @property
    def dx(self) -> float:
        return self.ds_km * 1000.0

    @property
    def nfields(self) -> int:
        """F = number of wafone-advected fields."""
        return 5 + 1 + (self.nqx - 1 if self.ipptls > 0 else 0) + (1 if self.ibltyp == 2 else 0) + self.ntr

    @property
    def needs_ext(self) -> bool:
        return bool(self.do_bdy or self.do_slice or self.do_massck or self.ibltyp == 2 or self.idiag or self.ichdiag)

    @property
    def cells(self) -> int:
        return self.jx * self.iy * self.kz

    def bytes_per_cell_update(self) -> int:
        """Algorithmic bytes B(F,nqx,ntr) of SURVEY.md section 8(d)."""
        return 8 * (self.mo_nadv * (self.mo_nsound * 38 + 6 + 16 + 17 * self.nfields)
                    + 50 + 4 * (self.nqx + self.ntr))


# BASELINE.json configs
WORKLOADS = {
    # 1: Testing/ideal.in shape with ds from isc24.in, doubly periodic
    "ideal": Workload("ideal", 500, 100, 60, i_band=1, i_crm=1, ds_km=2.0, dt=30.0),
    # 2: Testing/isc24_small.in
    "isc24_small": Workload("isc24_small", 100, 50, 30, i_band=1, i_crm=1, ds_km=2.0, dt=30.0),
    # 3: CORDEX-like 25 km limited-area, 10 tracers
    "cordex25": Workload("cordex25", 400, 400, 41, ntr=10, ds_km=25.0, dt=150.0, clat=45.0,
                         oro="gauss", msf_amp=0.05, nspgx=12),
    # 4: convection-permitting 3 km
    "cp3km": Workload("cp3km", 1536, 1536, 41, ds_km=3.0, dt=30.0, clat=45.0, oro="gauss",
                      msf_amp=0.05, nspgx=12),
    # 5: large tracer load
    "tracer40": Workload("tracer40", 1024, 1024, 41, ntr=40, ds_km=12.0, dt=90.0, clat=45.0,
                         oro="gauss", msf_amp=0.05, nspgx=12),
}


def small(wl: Workload, jx: int, iy: int, kz: int, **kw) -> Workload:
    """A reduced-size variant of a workload for parity tests."""
    return replace(wl, name=f"{wl.name}_{jx}x{iy}x{kz}", jx=jx, iy=iy, kz=kz, **kw)


# ---- deterministic hash noise in [-1, 1) (identical for any decomposition) --
def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def noise(shape, seed: int, salt: int) -> np.ndarray:
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64((seed * 1000003 + salt * 7919) & 0xFFFFFFFFFFFF)
        r = _splitmix64(idx)
    u = (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (2.0 * u - 1.0).reshape(shape)


# ---- Share/mod_zita.F90:40-178 ----------------------------------------------
def model_zitaf(kz, ztop):
    z = np.zeros(kz + 1)
    dz = ztop / float(kz)
    z[kz] = 0.0
    z[0] = ztop
    for k in range(kz - 1, 0, -1):       # do k = kz, 2, -1 (1-based)
        z[k] = z[k + 1] + dz
    return z


def model_zitah(kz, ztop):
    z = np.zeros(kz)
    dz = ztop / float(kz)
    z[kz - 1] = dz * 0.5
    z[0] = ztop - dz * 0.5
    for k in range(kz - 2, 0, -1):
        z[k] = z[k + 1] + dz
    return z


def _zfz(ztop, zh): return ztop / (np.exp(ztop / zh) - 1.0)
def bzita(zita, ztop, zh): return _zfz(ztop, zh) * (np.exp(zita / zh) - 1.0)
def bzitap(zita, ztop, zh): return _zfz(ztop, zh) * np.exp(zita / zh) / zh


def gzita(zita, ztop, a0):
    ratio = zita / ztop
    return ((0.0 - 1.0 * a0) * ratio - (3.0 - 2.0 * a0) * (ratio * ratio)
            + (2.0 - 1.0 * a0) * (ratio * ratio * ratio)) + 1.0


def gzitap(zita, ztop, a0):
    ratio = zita / ztop
    return ((0.0 - 1.0 * a0) * 1.0 - (6.0 - 4.0 * a0) * ratio + (6.0 - 3.0 * a0) * (ratio * ratio)) / ztop


def md_fmz(zita, geopot, ztop, zh, a0):
    return 1.0 / (gzitap(zita, ztop, a0) * (geopot * regrav) + bzitap(zita, ztop, zh))


def md_zeta(zita, geopot, ztop, zh, a0):
    return (geopot * regrav) * (gzita(zita, ztop, a0) - 1.0) + bzita(zita, ztop, zh)


def pfwsat(t, p):
    """Share/pfwsat.inc"""
    a = [0.611213476e+03, 0.444007856e+02, 0.143064234e+01, 0.264461437e-01, 0.305903558e-03,
         0.196237241e-05, 0.892344772e-08, -0.373208410e-10, 0.209339997e-13]
    c = [0.611123516e+03, 0.503109514e+02, 0.188369801e+01, 0.420547422e-01, 0.614396778e-03,
         0.602780717e-05, 0.387940929e-07, 0.149436277e-09, 0.262655803e-12]
    td = np.minimum(np.maximum(t - tzero, -75.0), 100.0)

    def horner(co):
        r = co[8]
        for n in range(7, -1, -1):
            r = co[n] + td * r
        return r
    es = np.where(td >= 0.0, np.minimum(horner(a), 0.15 * p), np.minimum(horner(c), 0.15 * p))
    return ep2 * (es / (p - es))


# ---- primary (file-like) inputs ----------------------------------------------
def _height(wl: Workload, x, y):
    """Analytic terrain height [m] at grid coordinates (x=j, y=i)."""
    if wl.oro == "flat":
        return np.zeros(np.broadcast(x, y).shape)
    if wl.oro == "sine":   # periodic hills
        return wl.oro_h * 0.25 * (1.0 + np.sin(2 * mathpi * x / wl.jx)) * (1.0 + np.cos(2 * mathpi * y / wl.iy))
    if wl.oro == "gauss":
        h = 0.0
        for (cx, cy, sx, sy, a) in ((0.30, 0.35, 0.08, 0.10, 1.0), (0.62, 0.55, 0.12, 0.07, 0.7),
                                    (0.45, 0.78, 0.06, 0.06, 0.5)):
            h = h + a * np.exp(-(((x / wl.jx - cx) / sx) ** 2 + ((y / wl.iy - cy) / sy) ** 2))
        return wl.oro_h * h / 1.0
    raise ValueError(wl.oro)


def _msf(wl: Workload, x, y):
    if wl.msf_amp == 0.0:
        return np.ones(np.broadcast(x, y).shape)
    return 1.0 + wl.msf_amp * np.sin(2 * mathpi * x / wl.jx) * np.cos(2 * mathpi * y / wl.iy)


def spline1d(xold, yold, y2, xnew):
    """Cubic spline through (xold, yold) with end second derivatives y2[0], y2[-1], evaluated at xnew: a
    restatement of spline1d, Share/mod_spline.F90:387-459 (1-based there), operation for operation."""
    xold, yold, y2, xnew = (np.array(v, dtype=np.float64) for v in (xold, yold, y2, xnew))
    nold, nnew = xold.size, xnew.size
    afac = 6.0
    bfac = 1.0 / afac
    X = lambda k: xold[k - 1]      # noqa: E731  (Fortran indices below)
    Y = lambda k: yold[k - 1]      # noqa: E731
    p, q = np.zeros(max(nold - 2, 1) + 1), np.zeros(max(nold - 2, 1) + 1)
    dxl = X(2) - X(1)
    dxr = X(3) - X(2)
    dydxl = (Y(2) - Y(1)) / dxl
    dydxr = (Y(3) - Y(2)) / dxr
    rtdxc = 0.5 / (dxl + dxr)
    p[1] = rtdxc * (bfac * (dydxr - dydxl) - dxl * y2[0])
    q[1] = -rtdxc * dxr
    if nold > 3:
        # the reference's loop runs k = 3 .. nold and would read xold(nold+1): only nold == 3 is ever used
        raise NotImplementedError("spline1d with more than three nodes")
    for k in range(nold - 1, 1, -1):
        y2[k - 1] = p[k - 1] + q[k - 1] * y2[k]
    ynew = np.zeros(nnew)
    k = -1
    ak = bk = ck = 0.0
    for k1 in range(1, nnew + 1):
        xk = xnew[k1 - 1]
        if xk < X(1):
            ynew[k1 - 1] = Y(1)
            continue
        if xk >= X(nold):
            ynew[k1 - 1] = Y(nold)
            continue
        kold = None
        for k2 in range(2, nold + 1):
            if X(k2) <= xk:
                continue
            kold = k2 - 1
            break
        if k != kold:
            k = kold
            y2k, y2kp1 = y2[k - 1], y2[k]
            dx = X(k + 1) - X(k)
            rdx = 1.0 / dx
            ak = bfac * rdx * (y2kp1 - y2k)
            bk = 0.5 * y2k
            ck = rdx * (Y(k + 1) - Y(k)) - bfac * dx * (y2kp1 + y2k + y2k)
        x = xk - X(k)
        xsq = x * x
        ynew[k1 - 1] = ak * xsq * x + bk * xsq + ck * x + Y(k)
    return ynew


def exponential_nudging(kz: int, ztop: float, high_nudge=3.0, medium_nudge=2.0, low_nudge=1.0) -> np.ndarray:
    """anudge(1:kz): exponential_nudging, Main/mpplib/mod_runparams.F90:611-636, on MOLOCH's sigma = 1 - zita/ztop
    (Main/mod_params.F90:2460-2465, Share/mod_zita.F90:40-45); defaults of Share/mod_dynparam.F90:150-152."""
    sigma = 1.0 - model_zitaf(kz, ztop) / ztop          # sigma(1:kz+1)
    hsigma = 1.0 - model_zitah(kz, ztop) / ztop         # hsigma(1:kz)
    kw = kz + 1                                          # findwhere(0.40): the loop index after a full loop
    for k in range(2, kz + 1):
        if sigma[k - 1] > 0.40:
            kw = k
            break
    zcin = [sigma[0], sigma[kw - 1], sigma[kz]]
    return spline1d(zcin, [high_nudge, medium_nudge, low_nudge], [0.0, 0.0, 0.0], hsigma)


def hefc_table(wl: Workload) -> np.ndarray:
    """Sponge coefficients hefc(n,k), the exponential branch of setup_bdycon's MOLOCH part
    (Main/mod_bdycod.F90:520-545): hefc(1) = 1, hefc(nspgx) = 0, exp(-(n-1)/anudge(k)) in between."""
    nsp, kz = wl.nspgx, wl.kz
    h = np.zeros((kz, nsp))
    anudge = exponential_nudging(kz, wl.mo_ztop)
    for k in range(kz):
        h[k, 0] = 1.0
        h[k, nsp - 1] = 0.0
        for n in range(2, nsp):
            h[k, n - 1] = math.exp(-float(n - 1) / anudge[k])
    return h


def make_primary(wl: Workload) -> dict:
    """Global 'file-like' inputs: terrain, map factors, latitudes, sponge table
    and the initial t, qx, u, v, trac, ps."""
    jx, iy, kz = wl.jx, wl.iy, wl.kz
    J, I = np.meshgrid(np.arange(1, jx + 1, dtype=np.float64), np.arange(1, iy + 1, dtype=np.float64))
    P = {}
    P["ht"] = _height(wl, J, I) * egrav            # geopotential, as mddom%ht (mod_params.F90:2438)
    P["htu"] = _height(wl, J - 0.5, I) * egrav
    P["htv"] = _height(wl, J, I - 0.5) * egrav
    P["msfx"] = _msf(wl, J, I)
    P["msfu"] = _msf(wl, J - 0.5, I)
    P["msfv"] = _msf(wl, J, I - 0.5)
    dl = raddeg * wl.dx / earthrad                  # Main/mod_params.F90:2122-2141
    P["xlat"] = wl.clat - dl * (float(iy) * 0.5 - I + 0.5)
    P["ulat"] = P["xlat"].copy()
    P["vlat"] = wl.clat - dl * (float(iy) * 0.5 - I + 1.0)
    P["rlat"] = wl.clat - dl * (float(iy) * 0.5 - np.arange(1, iy + 2, dtype=np.float64) + 1.0)
    if wl.nspgx > 0:
        P["hefc"] = hefc_table(wl)
    zitah = model_zitah(kz, wl.mo_ztop)
    zeta = md_zeta(zitah[:, None, None], P["ht"][None], wl.mo_ztop, wl.mo_h, wl.mo_a0)
    # base state: initideal-like profile + warm bubble + noise
    t = np.maximum(stdt - lrate * (zeta + P["ht"][None] * regrav), 210.0)
    r2 = ((J - 0.5 * jx) ** 2 + (I - 0.5 * iy) ** 2)[None] / 100.0 + ((zeta - 2000.0) / 1500.0) ** 2
    t = t + 2.0 * np.exp(-r2) + 1.0e-3 * noise(t.shape, wl.seed, 1)
    qx = np.zeros((wl.nqx, kz, iy, jx))
    qx[0] = 0.012 * np.exp(-(zeta + P["ht"][None] * regrav) / 2500.0)
    P["t"], P["qx"] = t, qx
    P["u"] = wl.u0 * (1.0 + 0.1 * noise(t.shape, wl.seed, 2))
    P["v"] = wl.v0 * (1.0 + 0.1 * noise(t.shape, wl.seed, 3))
    hsurf = P["ht"] * regrav
    P["ps"] = stdp * (1.0 - lrate * hsurf / stdt) ** (egrav / (rgas * lrate))
    if wl.ntr > 0:
        tr = np.full((wl.ntr, kz, iy, jx), 1.0e-9)
        rng = np.random.default_rng(wl.seed)
        for n in range(wl.ntr):
            cx, cy, cz = rng.uniform(0.15, 0.85), rng.uniform(0.15, 0.85), rng.uniform(500.0, 6000.0)
            rr = ((J - cx * jx) ** 2 + (I - cy * iy) ** 2)[None] / 64.0 + ((zeta - cz) / 1000.0) ** 2
            tr[n] += 1.0e-6 * np.exp(-rr)
        P["trac"] = tr
    return P


def make_boundary(wl: Workload, base: dict) -> dict:
    """ICBC stand-in: the b0/b1 buffers `bdyin` keeps for the lateral boundary
    (Main/mod_bdycod.F90:1079-1423; types v3dbound/v2dbound,
    Main/mpplib/mod_regcm_types.F90).  b0 is the large-scale state at the
    start of the boundary interval (here: the model's own initial state
    `base` with keys u, v, t, pai, qx, ps), b1 the state dtbdys later (b0
    plus a smooth large-scale change).  Global (nk, iy, jx) arrays."""
    jx, iy, kz = wl.jx, wl.iy, wl.kz
    J, I = np.meshgrid(np.arange(1, jx + 1, dtype=np.float64), np.arange(1, iy + 1, dtype=np.float64))
    wave = np.sin(2 * mathpi * J / jx)[None] * np.cos(2 * mathpi * I / iy)[None]
    lev = np.linspace(1.0, 0.3, kz)[:, None, None]
    B = {}
    B["dub0"] = np.array(base["u"], dtype=np.float64)
    B["dub1"] = B["dub0"] + 1.5 * wave * lev
    B["dvb0"] = np.array(base["v"], dtype=np.float64)
    B["dvb1"] = B["dvb0"] - 1.0 * wave * lev
    B["xtb0"] = np.array(base["t"], dtype=np.float64)
    B["xtb1"] = B["xtb0"] + 0.8 * wave * lev
    B["xpaib0"] = np.array(base["pai"], dtype=np.float64)
    B["xpaib1"] = B["xpaib0"] * (1.0 + 2.0e-4 * wave * lev)
    B["xqb0"] = np.array(base["qx"][0], dtype=np.float64)
    B["xqb1"] = B["xqb0"] * (1.0 + 0.05 * wave)
    cloud = np.exp(-((np.arange(kz)[:, None, None] - 0.6 * kz) / (0.15 * kz)) ** 2)
    B["xlb0"] = 2.0e-5 * cloud * (1.0 + 0.5 * wave)
    B["xlb1"] = 3.0e-5 * cloud * (1.0 - 0.5 * wave)
    B["xib0"] = 1.0e-5 * cloud * (1.0 - 0.3 * wave)
    B["xib1"] = 0.5e-5 * cloud * (1.0 + 0.3 * wave)
    B["xpsb0"] = np.array(base["ps"], dtype=np.float64)
    B["xpsb1"] = B["xpsb0"] * (1.0 + 1.0e-3 * wave[0])
    if wl.ntr > 0 and wl.ichebdy != 0:
        nz = noise((wl.ntr, kz, iy, jx), wl.seed, 11)
        B["chib0"] = 1.0e-9 * (1.0 + 0.2 * nz)
        B["chib1"] = 1.0e-9 * (1.0 - 0.2 * nz)
    return B


def make_tke(wl: Workload, zetaf: np.ndarray) -> np.ndarray:
    """Initial TKE on the kz+1 interfaces (UW PBL, ibltyp == 2): a shallow
    boundary-layer profile with noise, >= tkemin."""
    shp = (wl.kz + 1, wl.iy, wl.jx)
    return wl.tkemin + 0.4 * np.exp(-np.maximum(zetaf, 0.0) / 800.0) * (1.0 + 0.3 * noise(shp, wl.seed, 21))


def bdycon_setup(wl: Workload, zeta: np.ndarray | None = None) -> dict:
    """Host-side set-up of the lateral boundary: setup_bdycon, idynamic == 3
    branch (Main/mod_bdycod.F90:478-568), lowpass_init (:3844-3896) and the
    three ba%ibnd planes of setup_boundaries (Main/mod_atm_interface.F90:384-532).
    NumPy stand-in for the Fortran host code; arrays are global.  `zeta`: the
    (kz, iy, jx) level heights (computed from the workload's terrain if absent)."""
    jx, iy, kz = wl.jx, wl.iy, wl.kz
    perj = wl.i_band == 1 or wl.i_crm == 1
    peri = wl.i_crm == 1
    njc, nic = (jx if perj else jx - 1), (iy if peri else iy - 1)
    if zeta is None:
        J, I = np.meshgrid(np.arange(1, jx + 1, dtype=np.float64), np.arange(1, iy + 1, dtype=np.float64))
        ht = _height(wl, J, I) * egrav
        zeta = md_zeta(model_zitah(kz, wl.mo_ztop)[:, None, None], ht[None], wl.mo_ztop, wl.mo_h, wl.mo_a0)
    T = {"rtb": 1.0 / wl.dtbdys, "nztop": 0, "km": 0, "lm": 0}
    zztop = 18000.0
    gmeanz = np.zeros(kz)
    if wl.mo_top_nudge or wl.mo_spectral_nudge:
        gmeanz = zeta[:, :nic, :njc].reshape(kz, -1).sum(axis=1) / float(njc * nic)
        T["nztop"] = int((gmeanz > zztop).sum())
    T["gmeanz"] = gmeanz
    T["hefc"] = hefc_table(wl) if wl.nspgx > 0 else None
    T["fcx"] = chem_fcx(wl) if wl.nspgx > 0 else None
    k = np.arange(1, kz + 1)
    T["tnudge"] = np.where(k <= T["nztop"], np.sin(0.5 * mathpi * (gmeanz - zztop) / (wl.mo_h - zztop)) ** 2, 0.0) \
        if wl.mo_top_nudge else np.zeros(kz)
    if wl.mo_spectral_nudge:
        km = max(int(np.rint((njc * wl.ds_km) / 1500.0)), 1)
        lm = max(int(np.rint((nic * wl.ds_km) / 750.0)), 1)
        dx, dy = mathpi / float(njc - 1), mathpi / float(nic - 1)
        kk = np.arange(1, 2 * km + 1, dtype=np.float64)[:, None]
        ll = np.arange(1, 2 * lm + 1, dtype=np.float64)[:, None]
        jj = np.arange(1, jx + 1)[None, :]
        ii = np.arange(1, iy + 1)[None, :]
        T["km"], T["lm"] = km, lm
        T["bvx"] = np.sqrt(2.0 / float(jx - 1) * np.exp(-(kk / km) ** 2)) * np.sin((kk * (jj - 2)) * dx)
        T["bvy"] = np.sqrt(2.0 / float(iy - 1) * np.exp(-(ll / lm) ** 2)) * np.sin((ll * (ii - 2)) * dy)
        T["cnudge"] = (wl.dtrad * T["rtb"]) * np.minimum((gmeanz / wl.mo_h) ** 2, 1.0)
    T["ibnd"] = {"cr": _ibnd(wl, False, False), "ud": _ibnd(wl, True, False), "vd": _ibnd(wl, False, True)}
    return T


def chem_fcx(wl: Workload) -> np.ndarray:
    """Tracer relaxation weights fcx(1:nspgx): setup_che_bdycon, idynamic == 3 branch
    (Main/chemlib/mod_che_bdyco.F90:1036-1039): fcx(1) = 1, fcx(nspgx) = 0, Lehmann coefficients in
    between.  The reference stops with a fatal error unless nspgx-2 is a power of two; for the other
    sponge widths of the synthetic workloads a linear ramp stands in."""
    m = wl.nspgx - 2
    if m >= 1 and (m & (m - 1)) == 0:
        f = np.zeros(wl.nspgx)
        f[0] = 1.0
        f[1:wl.nspgx - 1] = relax_coefficients(m, 0.1, 1.0)
        return f
    n = np.arange(1, wl.nspgx + 1, dtype=np.float64)
    return np.where((n >= 2) & (n <= wl.nspgx - 1), 0.5 * (wl.nspgx - n) / max(wl.nspgx - 2, 1), 0.0)


def relax_coefficients(npts: int, gmmin: float, gmmax: float) -> np.ndarray:
    """Lehmann (1993) optimal relaxation coefficients: host-side restatement of
    relax_coefficients (Main/mpplib/mod_runparams.F90:645-697) for
    bdy_use_lehmann runs (Main/mod_bdycod.F90:524-535)."""
    npmax = 32
    p, q = np.zeros(2 * npmax + 1), np.zeros(2 * npmax + 1)
    n = 1
    p[1], q[0] = 1.0, 1.0
    my = np.sqrt(gmmax / gmmin)
    while n < npts:
        my = np.sqrt((my + 1.0 / my) / 2.0)
        pp, qq = np.zeros(2 * npmax + 1), np.zeros(2 * npmax + 1)
        for i in range(n + 1):
            for j in range(n + 1):
                pp[i + j] += p[i] * p[j] + q[i] * q[j]
                qq[i + j] += 2.0 * my * p[i] * q[j]
        p[:2 * n + 1], q[:2 * n + 1] = pp[:2 * n + 1], qq[:2 * n + 1]
        n = 2 * n
    if n != npts and npts != 1:
        raise ValueError("nbl np not a power of 2")       # fatal(...,'INVALID BOUNDARY POINT NUMBER.')
    coeff = np.zeros(npts)
    for i in range(n, 0, -1):
        kk = p[i] / q[i - 1]
        for j in range(i, 0, -1):
            xxx = q[j]
            q[j] = p[j] - kk * q[j - 1]
            p[j] = xxx
        xxx = q[0]
        q[0] = p[0]
        p[0] = xxx
        kdt2 = kk * np.sqrt(gmmin * gmmax)
        coeff[i - 1] = kdt2 / (1.0 + kdt2)
    return coeff


def hefc_lehmann(wl: Workload) -> np.ndarray:
    """hefc(n,k) of the bdy_use_lehmann branch (Main/mod_bdycod.F90:520-535); nspgx-2 must be a power of 2."""
    nsp, kz = wl.nspgx, wl.kz
    cflmax = min(0.999, (300.0 * (wl.dt / float(wl.mo_nadv))) / (2.0 * wl.dx))
    cflmin = max(0.001, (10.0 * (wl.dt / float(wl.mo_nadv))) / (2.0 * wl.dx))
    h = np.zeros((kz, nsp))
    h[:, 0] = 1.0
    h[:, 1:nsp - 1] = relax_coefficients(nsp - 2, cflmin, cflmax)[None, :]
    return h


# ---- derived static fields (compute_moloch_static + init_moloch) ------------
def _ibnd(wl: Workload, ldotx: bool, ldoty: bool) -> np.ndarray:
    """ba%ibnd of setup_boundaries (Main/mod_atm_interface.F90:384-532)."""
    jx, iy, nsp = wl.jx, wl.iy, wl.nspgx
    ib = np.full((iy, jx), -1, dtype=np.int64)
    crm = wl.i_crm == 1
    band = wl.i_band == 1 or crm
    if nsp <= 0 or crm:
        return ib
    J, I = np.meshgrid(np.arange(1, jx + 1), np.arange(1, iy + 1))
    jcx, icy = (0 if ldotx else 1), (0 if ldoty else 1)
    igbb1, igbb2, jgbl1, jgbl2 = 2, nsp - 1, 2, nsp - 1
    igbt1, igbt2 = iy - icy - nsp + 2, iy - 1 - icy
    jgbr1, jgbr2 = jx - jcx - nsp + 2, jx - 1 - jcx
    if band:
        jgbl1, jgbr2 = 1, jx - jcx
        south = (I >= igbb1) & (I <= igbb2) & ~((J < jgbl1) & (J > jgbr2))
        ib[south] = (I - igbb1 + 2)[south]
        north = (I >= igbt1) & (I <= igbt2) & ~((J < jgbl1) & (J > jgbr2))
        ib[north] = (igbt2 - I + 2)[north]
        return ib
    inj = (J >= jgbl1) & (J <= jgbr2)
    south = ((I >= igbb1) & (I <= igbb2) & inj & ~((J <= jgbl2) & (I >= J))
             & ~((J >= jgbr1) & (I >= (jgbr2 - J + 2))))
    ib[south] = (I - igbb1 + 2)[south]
    north = ((I >= igbt1) & (I <= igbt2) & inj & ~((J <= jgbl2) & (J <= (igbt2 - I + 2)))
             & ~((J >= jgbr1) & ((igbt2 - I) >= (jgbr2 - J))))
    ib[north] = (igbt2 - I + 2)[north]
    mid = ~(north | south) & (I >= igbb1) & (I <= igbt2)
    west = mid & (J >= jgbl1) & (J <= jgbl2)
    ib[west] = (J - jgbl1 + 2)[west]
    east = mid & (J >= jgbr1) & (J <= jgbr2)
    ib[east] = (jgbr2 - J + 2)[east]
    return ib


def _bdywt(wl: Workload, ib: np.ndarray, hefc) -> np.ndarray:
    """setup_bdywt (Main/mod_bdycod.F90:4033-4047)."""
    m = np.ones((wl.kz,) + ib.shape)
    if wl.nspgx > 0 and hefc is not None:
        sel = ib > 0
        idx = np.where(sel, ib - 1, 0)
        m = np.where(sel[None], 1.0 - hefc[:, idx], 1.0)
    return m


def derive_static(wl: Workload, P: dict) -> dict:
    """compute_moloch_static (Main/mod_params.F90:3316-3395), the zita levels
    (:2461-2463), setup_bdywt and ffilt (Main/mod_init.F90:1008-1026) on the
    global grid.  Rows/columns outside a field's owned range hold values that
    are never read."""
    jx, iy, kz = wl.jx, wl.iy, wl.kz
    ztop, zh, a0 = wl.mo_ztop, wl.mo_h, wl.mo_a0
    S = {}
    zita, zitah = model_zitaf(kz, ztop), model_zitah(kz, ztop)
    S["zita"], S["zitah"] = zita, zitah
    S["mo_dzita"] = float(zita[kz - 1])
    rdx = 1.0 / wl.dx
    ht, htu, htv = P["ht"], P["htu"], P["htv"]
    perj = wl.i_band == 1 or wl.i_crm == 1
    peri = wl.i_crm == 1
    htm1 = np.roll(ht, 1, axis=1) if perj else np.concatenate([ht[:, :1], ht[:, :-1]], axis=1)
    hx = rdx * regrav * P["msfu"] * (ht - htm1)
    hx[:, 0] = 2.0 * rdx * regrav * P["msfu"][:, 0] * (ht[:, 0] - htu[:, 0])      # j == 1
    htm1 = np.roll(ht, 1, axis=0) if peri else np.concatenate([ht[:1], ht[:-1]], axis=0)
    mfv = 1.0 if wl.lrotllr else P["msfv"]
    hy = rdx * regrav * mfv * (ht - htm1)
    hy[0] = (2.0 * rdx * regrav * mfv * (ht - htv))[0]                            # i == 1
    S["hx"], S["hy"] = hx, hy
    S["zeta"] = md_zeta(zitah[:, None, None], ht[None], ztop, zh, a0)
    S["fmz"] = md_fmz(zitah[:, None, None], ht[None], ztop, zh, a0)
    S["rfmzu"] = 1.0 / md_fmz(zitah[:, None, None], htu[None], ztop, zh, a0)
    S["rfmzv"] = 1.0 / md_fmz(zitah[:, None, None], htv[None], ztop, zh, a0)
    S["fmzf"] = md_fmz(zita[:, None, None], ht[None], ztop, zh, a0)
    S["zetaf"] = md_zeta(zita[:, None, None], ht[None], ztop, zh, a0)
    hefc = P.get("hefc")
    S["bdywtw"] = _bdywt(wl, _ibnd(wl, False, False), hefc)
    S["bdywtu"] = _bdywt(wl, _ibnd(wl, True, False), hefc)
    S["bdywtv"] = _bdywt(wl, _ibnd(wl, False, True), hefc)
    # ffilt: sponge in the implicit solver above 18 km
    njc = jx if perj else jx - 1
    nic = iy if peri else iy - 1
    gmeanz = S["zeta"][:, :nic, :njc].reshape(kz, -1).sum(axis=1) / float(njc * nic)
    zzi = (gmeanz - 18000.0) / (ztop - 18000.0)
    S["ffilt"] = np.where(gmeanz < 18000.0, 0.0, mo_zfilt_fac * np.sin(0.5 * mathpi * zzi) ** 2)
    return S


def init_state(wl: Workload, P: dict, S: dict) -> dict:
    """paicompute (Main/mod_bdycod.F90:3762-3796) + Main/mod_init.F90:941-953."""
    kz = wl.kz
    t, q, z = P["t"], P["qx"][0], S["zeta"]
    pai = np.zeros_like(t)
    zdelta = z[kz - 1] * egrav
    tv1 = t[kz - 1] * (1.0 + ep1 * q[kz - 1])
    tv2 = t[kz - 2] * (1.0 + ep1 * q[kz - 2])
    lrt = (tv2 - tv1) / (z[kz - 2] - z[kz - 1])
    lrt = np.where(lrt > govcp, govcp, np.where(lrt < -0.005, 0.5 * lrt - 0.5 * lrate, lrt))
    tv = tv1 - 0.5 * z[kz - 1] * lrt
    zz = 1.0 / (rgas * tv)
    p = P["ps"] * np.exp(-zdelta * zz)
    paikp1 = (p / p00) ** rovcp
    pai[kz - 1] = paikp1
    for k in range(kz - 2, -1, -1):
        tv1 = t[k] * (1.0 + ep1 * q[k])
        tv2 = t[k + 1] * (1.0 + ep1 * q[k + 1])
        zb = 2.0 * egrav * S["mo_dzita"] / (S["fmzf"][k + 1] * cpd) + tv1 - tv2
        zdelta = np.sqrt(zb * zb + 4.0 * tv2 * tv1)
        paikp1 = -paikp1 / (2.0 * tv2) * (zb - zdelta)
        pai[k] = paikp1
    st = {"pai": pai}
    st["p"] = pai ** cpovr * p00
    st["qsat"] = pfwsat(t, st["p"])
    st["rho"] = st["p"] / (rgas * t)
    st["tvirt"] = t * (1.0 + ep1 * q)
    st["tetav"] = st["tvirt"] / pai
    st["w"] = np.zeros((kz + 1,) + t.shape[1:])
    return st


def model_inputs(wl: Workload):
    """Everything `init_moloch` hands to the device, on the global grid:
    (fields, profiles) for MolochB200.init_moloch.  NumPy stand-in for the
    Fortran host set-up (compute_moloch_static, init_moloch's 1-D tables,
    paicompute); used by bench.py, where no oracle is involved."""
    P = make_primary(wl)
    St = derive_static(wl, P)
    X = init_state(wl, P, St)
    kz = wl.kz
    F = {k: St[k] for k in ("fmz", "fmzf", "rfmzu", "rfmzv", "zeta", "hx", "hy", "bdywtu", "bdywtv", "bdywtw")}
    F["msfx"], F["msfu"], F["msfv"] = P["msfx"], P["msfu"], P["msfv"]
    F["coru"] = eomeg2 * np.sin(P["ulat"] * degrad)      # Main/mod_moloch.F90:260-261
    F["corv"] = eomeg2 * np.sin(P["vlat"] * degrad)
    F.update(u=P["u"], v=P["v"], t=P["t"], qx=P["qx"], ps=P["ps"])
    if wl.ntr > 0:
        F["trac"] = P["trac"]
    F.update(pai=X["pai"], tetav=X["tetav"], tvirt=X["tvirt"], p=X["p"], rho=X["rho"], qsat=X["qsat"], w=X["w"])
    k = np.arange(1, kz + 1, dtype=np.float64)
    prof = {
        "gzitak": gzita(St["zita"], wl.mo_ztop, wl.mo_a0),
        "gzitakh": gzita(St["zitah"], wl.mo_ztop, wl.mo_a0),
        "ffilt": St["ffilt"],
        "xkdamp": 0.125 * 0.850 * (1.0 / (k + 1.0) - 1.0 / (kz + 2.0)),          # :295-296
        "xknu": 0.125 * (0.55 + 0.45 * ((kz - k + 1.0) - 1.0) / (kz - 1.0)),     # :297-298
    }
    if wl.lrotllr:
        prof["rlat"] = P["rlat"]
    return F, prof


# ---- rank-local generation (large grids: no global 3-D arrays) ---------------
def _noise_at(lin: np.ndarray, seed: int, salt: int) -> np.ndarray:
    """noise() evaluated at global linear indices `lin` (same values)."""
    with np.errstate(over="ignore"):
        idx = lin.astype(np.uint64) + np.uint64((seed * 1000003 + salt * 7919) & 0xFFFFFFFFFFFF)
        r = _splitmix64(idx)
    u = (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return 2.0 * u - 1.0


def model_inputs_local(wl: Workload, g):
    """(fields, profiles, boxes) of ONE rank (geometry `g`, regcm_b200.decomp.Geom)
    without ever building a global 3-D array: what model_inputs(wl) gives, cut to
    the rank's bounds.  2-D fields are still computed globally (they are cheap)."""
    from . import hostmodel as H
    jx, iy, kz = wl.jx, wl.iy, wl.kz
    ztop, zh, a0 = wl.mo_ztop, wl.mo_h, wl.mo_a0
    Jg, Ig = np.meshgrid(np.arange(1, jx + 1, dtype=np.float64), np.arange(1, iy + 1, dtype=np.float64))
    P2 = {"ht": _height(wl, Jg, Ig) * egrav, "htu": _height(wl, Jg - 0.5, Ig) * egrav,
          "htv": _height(wl, Jg, Ig - 0.5) * egrav, "msfx": _msf(wl, Jg, Ig), "msfu": _msf(wl, Jg - 0.5, Ig),
          "msfv": _msf(wl, Jg, Ig - 0.5)}
    dl = raddeg * wl.dx / earthrad
    xlat = wl.clat - dl * (float(iy) * 0.5 - Ig + 0.5)
    vlat = wl.clat - dl * (float(iy) * 0.5 - Ig + 1.0)
    rlat = wl.clat - dl * (float(iy) * 0.5 - np.arange(1, iy + 2, dtype=np.float64) + 1.0)
    hsurf = P2["ht"] * regrav
    ps2 = stdp * (1.0 - lrate * hsurf / stdt) ** (egrav / (rgas * lrate))
    # 2-D statics exactly as derive_static does
    rdx = 1.0 / wl.dx
    perj = wl.i_band == 1 or wl.i_crm == 1
    peri = wl.i_crm == 1
    ht, htu, htv = P2["ht"], P2["htu"], P2["htv"]
    htm1 = np.roll(ht, 1, axis=1) if perj else np.concatenate([ht[:, :1], ht[:, :-1]], axis=1)
    hx = rdx * regrav * P2["msfu"] * (ht - htm1)
    hx[:, 0] = 2.0 * rdx * regrav * P2["msfu"][:, 0] * (ht[:, 0] - htu[:, 0])
    htm1 = np.roll(ht, 1, axis=0) if peri else np.concatenate([ht[:1], ht[:-1]], axis=0)
    mfv = 1.0 if wl.lrotllr else P2["msfv"]
    hy = rdx * regrav * mfv * (ht - htm1)
    hy[0] = (2.0 * rdx * regrav * mfv * (ht - htv))[0]
    zita, zitah = model_zitaf(kz, ztop), model_zitah(kz, ztop)
    mo_dzita = float(zita[kz - 1])
    njc = jx if perj else jx - 1
    nic = iy if peri else iy - 1
    gmeanz = np.array([md_zeta(zitah[k], ht[:nic, :njc], ztop, zh, a0).sum() / float(njc * nic) for k in range(kz)])
    zzi = (gmeanz - 18000.0) / (ztop - 18000.0)
    ffilt = np.where(gmeanz < 18000.0, 0.0, mo_zfilt_fac * np.sin(0.5 * mathpi * zzi) ** 2)
    hefc = hefc_table(wl) if wl.nspgx > 0 else None

    # the rank's working box: dot range + 2 ghosts wherever a neighbour exists
    box = g.ext("dot", 2, 2)
    jlo, jhi, ilo, ihi = box
    jv = (np.arange(jlo, jhi + 1) - 1) % jx            # 0-based wrapped global indices
    iv = (np.arange(ilo, ihi + 1) - 1) % iy
    cut2 = lambda a: np.ascontiguousarray(a[iv[:, None], jv[None, :]])
    J = (jv + 1).astype(np.float64)[None, :] + 0.0 * iv[:, None]
    I = (iv + 1).astype(np.float64)[:, None] + 0.0 * jv[None, :]
    htl, htul, htvl = cut2(ht), cut2(htu), cut2(htv)
    F3 = {}
    F3["zeta"] = md_zeta(zitah[:, None, None], htl[None], ztop, zh, a0)
    F3["fmz"] = md_fmz(zitah[:, None, None], htl[None], ztop, zh, a0)
    F3["rfmzu"] = 1.0 / md_fmz(zitah[:, None, None], htul[None], ztop, zh, a0)
    F3["rfmzv"] = 1.0 / md_fmz(zitah[:, None, None], htvl[None], ztop, zh, a0)
    F3["fmzf"] = md_fmz(zita[:, None, None], htl[None], ztop, zh, a0)
    for name, ldx, ldy in (("bdywtw", False, False), ("bdywtu", True, False), ("bdywtv", False, True)):
        ib = cut2(_ibnd(wl, ldx, ldy))
        if wl.nspgx > 0:
            sel = ib > 0
            idx = np.where(sel, ib - 1, 0)
            F3[name] = np.where(sel[None], 1.0 - hefc[:, idx], 1.0)
        else:
            F3[name] = np.ones((kz,) + ib.shape)
    zeta = F3["zeta"]
    lin = (np.arange(kz, dtype=np.int64)[:, None, None] * iy + iv[None, :, None]) * jx + jv[None, None, :]
    t = np.maximum(stdt - lrate * (zeta + htl[None] * regrav), 210.0)
    r2 = ((J - 0.5 * jx) ** 2 + (I - 0.5 * iy) ** 2)[None] / 100.0 + ((zeta - 2000.0) / 1500.0) ** 2
    t = t + 2.0 * np.exp(-r2) + 1.0e-3 * _noise_at(lin, wl.seed, 1)
    qx = np.zeros((wl.nqx, kz) + htl.shape)
    qx[0] = 0.012 * np.exp(-(zeta + htl[None] * regrav) / 2500.0)
    u = wl.u0 * (1.0 + 0.1 * _noise_at(lin, wl.seed, 2))
    v = wl.v0 * (1.0 + 0.1 * _noise_at(lin, wl.seed, 3))
    F3.update(t=t, qx=qx, u=u, v=v)
    if wl.ntr > 0:
        tr = np.full((wl.ntr, kz) + htl.shape, 1.0e-9)
        rng = np.random.default_rng(wl.seed)
        for n in range(wl.ntr):
            cx, cy, cz = rng.uniform(0.15, 0.85), rng.uniform(0.15, 0.85), rng.uniform(500.0, 6000.0)
            rr = ((J - cx * jx) ** 2 + (I - cy * iy) ** 2)[None] / 64.0 + ((zeta - cz) / 1000.0) ** 2
            tr[n] += 1.0e-6 * np.exp(-rr)
        F3["trac"] = tr
    St = {"zeta": zeta, "fmzf": F3["fmzf"], "mo_dzita": mo_dzita}
    X = init_state(wl, {"t": t, "qx": qx, "ps": cut2(ps2)}, St)
    F3.update(pai=X["pai"], tetav=X["tetav"], tvirt=X["tvirt"], p=X["p"], rho=X["rho"], qsat=X["qsat"], w=X["w"])
    F2 = {"hx": cut2(hx), "hy": cut2(hy), "msfx": cut2(P2["msfx"]), "msfu": cut2(P2["msfu"]),
          "msfv": cut2(P2["msfv"]), "coru": cut2(eomeg2 * np.sin(xlat * degrad)),
          "corv": cut2(eomeg2 * np.sin(vlat * degrad)), "ps": cut2(ps2)}
    fields, boxes = {}, {}
    for name, arr in list(F3.items()) + list(F2.items()):
        b = H.bounds(g, name)
        sub = arr[..., b[2] - ilo:b[3] - ilo + 1, b[0] - jlo:b[1] - jlo + 1]
        fields[name] = np.ascontiguousarray(sub)
        boxes[name] = b
    k = np.arange(1, kz + 1, dtype=np.float64)
    prof = {"gzitak": gzita(zita, ztop, a0), "gzitakh": gzita(zitah, ztop, a0), "ffilt": ffilt,
            "xkdamp": 0.125 * 0.850 * (1.0 / (k + 1.0) - 1.0 / (kz + 2.0)),
            "xknu": 0.125 * (0.55 + 0.45 * ((kz - k + 1.0) - 1.0) / (kz - 1.0))}
    if wl.lrotllr:
        prof["rlat"] = rlat
    return fields, prof, boxes
