"""CPU tests of the host side: the C-ABI library loads and exports every
symbol of include/moloch_b200.h, fails loudly without a GPU, and the
decomposition / halo-plan logic (the N>1 path) is right -- including a
world_size-2 run over gloo that moves real halos between two processes."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from regcm_b200 import hostmodel as H
from regcm_b200 import moloch as M
from regcm_b200 import synthetic as S
from regcm_b200.decomp import default_cpus_per_dim, make_geom

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "moloch_b200.h")).read()
    declared = set(re.findall(r"\b(moloch_b200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(M.ABI_SYMBOLS), declared ^ set(M.ABI_SYMBOLS)
    lib = ctypes.CDLL(M.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    assert M.load_library().moloch_b200_abi_version() == 3


def test_plain_c_client_compiles_links_and_gets_the_error_contract(tmp_path):
    """The boundary is a C ABI: a C99 translation unit (-Wall -Wextra -pedantic -Werror) that includes the header,
    takes the address of every declared entry point and links against the library; it then runs the calls that
    need no GPU (version, struct size, device count) and the refusal of an invalid configuration with its message
    (RegCM's `fatal` text comes from moloch_b200_last_error)."""
    hdr = open(os.path.join(ROOT, "include", "moloch_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(moloch_b200_[a-z_0-9]+)\s*\(", hdr)))
    src = tmp_path / "abi_client.c"
    src.write_text(
        '#include "moloch_b200.h"\n#include <stdio.h>\n#include <string.h>\n'
        "typedef void (*any_fn)(void);\n"
        "int main(void) {\n"
        "  any_fn tab[] = {" + ", ".join(f"(any_fn){n}" for n in declared) + "};\n"
        "  size_t n = sizeof tab / sizeof tab[0], q;\n"
        "  moloch_b200_config cfg;\n  moloch_b200_ctx* ctx = NULL;\n"
        "  for (q = 0; q < n; ++q) if (!tab[q]) return 1;\n"
        "  if (moloch_b200_abi_version() != 3) return 2;\n"
        "  if ((size_t)moloch_b200_config_size() != sizeof cfg) return 3;\n"
        "  memset(&cfg, 0, sizeof cfg);\n"
        "  if (moloch_b200_create(&cfg, &ctx) == 0 || ctx != NULL) return 4;   /* an all-zero configuration is refused */\n"
        '  if (!strstr(moloch_b200_last_error(), "jx, iy, kz")) return 5;\n'
        "  if (moloch_b200_create(NULL, &ctx) == 0) return 6;\n"
        '  printf("%d entry points, %d devices\\n", (int)n, moloch_b200_device_count());\n'
        "  return 0;\n}\n")
    exe = tmp_path / "abi_client"
    libdir = os.path.dirname(M.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(M.LIB_PATH),
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith(f"{len(declared)} entry points")


def test_enums_match_header():
    hdr = open(os.path.join(ROOT, "include", "moloch_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", re.search(r"enum moloch_b200_field \{(.*?)\};", hdr, re.S).group(1), flags=re.S)
    names = [n.strip().split("=")[0].strip() for n in body.replace("\n", " ").split(",") if n.strip()]
    names = [n for n in names if n != "MB_NFIELDS"]
    assert [n[3:].lower() for n in names] == M.FIELDS
    body = re.sub(r"/\*.*?\*/", "", re.search(r"enum moloch_b200_table \{(.*?)\};", hdr, re.S).group(1), flags=re.S)
    names = [n.strip().split("=")[0].strip() for n in body.split(",")]
    names = [n for n in names if n and n != "MB_NTABLES"]
    assert [n[7:].lower() for n in names] == M.TABLES
    body = re.search(r"enum moloch_b200_profile \{(.*?)\};", hdr, re.S).group(1)
    names = [re.sub(r"/\*.*?\*/", "", n, flags=re.S).strip().split("=")[0].strip() for n in body.split(",")]
    names = [n for n in names if n and n != "MB_NPROFILES"]
    assert [n[3:].lower() for n in names] == M.PROFILES


def test_config_struct_layout():
    # 34 int32 + 3 double (ABI v1) + 14 int32 + 5 double + 4 int32 (ABI v2)
    assert ctypes.sizeof(M.Config) == 34 * 4 + 3 * 8 + 14 * 4 + 5 * 8 + 4 * 4
    lib = M.load_library()
    assert int(lib.moloch_b200_config_size()) == ctypes.sizeof(M.Config)      # the C side agrees


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (this container has no
    GPU; on the GPU box the test checks the bad-config error path instead)."""
    lib = M.load_library()
    wl = S.small(S.WORKLOADS["isc24_small"], 16, 12, 8)
    m = M.MolochB200(wl)
    if lib.moloch_b200_device_count() == 0:
        with pytest.raises(M.MolochError, match="no CUDA device"):
            m.allocate_moloch()
    bad = M.MolochB200(wl)
    bad.cfg.kz = 1
    with pytest.raises(M.MolochError, match="jx, iy, kz"):
        bad.allocate_moloch()
    tiny = M.MolochB200(wl)
    tiny.cfg.jde2 = tiny.cfg.jde1 + 1
    with pytest.raises(M.MolochError, match="less than 3x3"):
        tiny.allocate_moloch()


def test_default_process_grid_matches_set_nproc():
    # Main/mpplib/mod_mppparam.F90:1381-1401
    assert default_cpus_per_dim(1, 400, 400) == (1, 1)
    assert default_cpus_per_dim(2, 400, 400) == (2, 1)
    assert default_cpus_per_dim(4, 400, 400) == (2, 2)
    assert default_cpus_per_dim(8, 400, 400) == (2, 4)
    assert default_cpus_per_dim(8, 1536, 1536) == (2, 4)


@pytest.mark.parametrize("band,crm", [(0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("px,py", [(1, 1), (2, 1), (2, 2), (3, 2), (2, 4)])
def test_geometry_tiles_the_domain(px, py, band, crm):
    jx, iy = 47, 53
    seen_dot = np.zeros((iy, jx), dtype=int)
    seen_cross = np.zeros((iy, jx), dtype=int)
    geoms = [make_geom(jx, iy, 10, band, crm, px, py, r) for r in range(px * py)]
    for g in geoms:
        seen_dot[g.ide1 - 1:g.ide2, g.jde1 - 1:g.jde2] += 1
        seen_cross[g.ice1 - 1:g.ice2, g.jce1 - 1:g.jce2] += 1
        # neighbour symmetry
        if g.left >= 0: assert geoms[g.left].right == g.rank
        if g.right >= 0: assert geoms[g.right].left == g.rank
        if g.top >= 0: assert geoms[g.top].bottom == g.rank
        if g.bottom >= 0: assert geoms[g.bottom].top == g.rank
        assert g.bl == (g.left < 0) and g.bt == (g.top < 0)
    assert (seen_dot == 1).all()
    nj = jx if (band or crm) else jx - 1
    ni = iy if crm else iy - 1
    assert (seen_cross[:ni, :nj] == 1).all() and seen_cross[ni:, :].sum() == 0 and seen_cross[:, nj:].sum() == 0


def _emulate_exchange(wl, px, py, stag, name, nex, lr, bt):
    """Apply the library's halo plan with NumPy on all ranks of a px x py grid
    and compare every received ghost with the global field."""
    rng = np.random.default_rng(1)
    glob = rng.standard_normal((wl.iy, wl.jx))
    geoms = [make_geom(wl.jx, wl.iy, wl.kz, wl.i_band, wl.i_crm, px, py, r) for r in range(px * py)]
    nbr_of = lambda g: [g.left, g.right, g.bottom, g.top]
    loc, box, plan = [], [], []
    for g in geoms:
        b = g.ext(name, 2, 2)
        arr = np.full((b[3] - b[2] + 1, b[1] - b[0] + 1), np.nan)
        own = g.ext(name, 0, 0)
        arr[own[2] - b[2]:own[3] - b[2] + 1, own[0] - b[0]:own[1] - b[0] + 1] = H.cut(glob, g, own)
        loc.append(arr); box.append(b)
        plan.append(M.halo_plan(M.make_config(wl, g), stag, nex, lr, bt))
    cutb = lambda r, bx: loc[r][bx[2] - box[r][2]:bx[3] - box[r][2] + 1, bx[0] - box[r][0]:bx[1] - box[r][0] + 1]
    opposite = [1, 0, 3, 2]
    checked = 0
    for r, g in enumerate(geoms):
        send, recv = plan[r]
        for sd in range(4):
            n = nbr_of(g)[sd]
            if recv[sd][0] > recv[sd][1]:
                assert n < 0 or not ((sd < 2 and lr) or (sd >= 2 and bt))
                continue
            src = cutb(n, plan[n][0][opposite[sd]])       # the neighbour's send box towards me
            dst = cutb(r, recv[sd])
            assert src.shape == dst.shape
            dst[...] = src
            want = H.cut(glob, g, tuple(recv[sd]))
            assert np.array_equal(dst, want), (r, sd)
            checked += 1
    return checked


@pytest.mark.parametrize("crm", [0, 1])
def test_halo_plan_matches_reference_exchange_semantics(crm):
    """ghost (j1-iex) <- neighbour's (j2-(iex-1)) etc.
    (Main/mpplib/mod_mppparam.F90:3843-3874)."""
    wl = S.small(S.WORKLOADS["isc24_small"], 23, 19, 6, i_band=crm, i_crm=crm)
    n = 0
    for px, py in ((2, 1), (2, 2), (3, 2)):
        for stag, name in ((0, "cross"), (1, "u"), (2, "v"), (3, "dot")):
            for nex in (1, 2):
                n += _emulate_exchange(wl, px, py, stag, name, nex, True, True)
                n += _emulate_exchange(wl, px, py, stag, name, nex, True, False)
                n += _emulate_exchange(wl, px, py, stag, name, nex, False, True)
    assert n > 100


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from regcm_b200 import hostmodel as H, moloch as M, synthetic as S
from regcm_b200.decomp import make_geom
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
wl = S.small(S.WORKLOADS["cordex25"], 30, 22, 5, ntr=0, nspgx=0)
g = make_geom(wl.jx, wl.iy, wl.kz, wl.i_band, wl.i_crm, world, 1, rank)
glob = np.random.default_rng(7).standard_normal((wl.kz, wl.iy, wl.jx))
box = g.ext("cross", 2, 2)
loc = np.full((wl.kz, box[3] - box[2] + 1, box[1] - box[0] + 1), np.nan)
own = g.ext("cross", 0, 0)
loc[:, own[2]-box[2]:own[3]-box[2]+1, own[0]-box[0]:own[1]-box[0]+1] = H.cut(glob, g, own)
send, recv = M.halo_plan(M.make_config(wl, g), 0, 2, True, False)
view = lambda b: loc[:, b[2]-box[2]:b[3]-box[2]+1, b[0]-box[0]:b[1]-box[0]+1]
nbrs = [g.left, g.right]
reqs, bufs = [], {}
for sd in (0, 1):                      # sends in side order L, R
    if send[sd][0] <= send[sd][1]:
        t = torch.from_numpy(np.ascontiguousarray(view(send[sd])))
        reqs.append(dist.isend(t, nbrs[sd]))
for sd in (1, 0):                      # receives in order R, L (see halo.cu)
    if recv[sd][0] <= recv[sd][1]:
        bufs[sd] = torch.empty(view(recv[sd]).shape, dtype=torch.float64)
        reqs.append(dist.irecv(bufs[sd], nbrs[sd]))
for r in reqs: r.wait()
ok = True
for sd, t in bufs.items():
    view(recv[sd])[...] = t.numpy()
    ok &= np.array_equal(view(recv[sd]), H.cut(glob, g, tuple(recv[sd])))
# one value per rank: all ranks must agree that every halo is right
flag = torch.tensor([1.0 if ok and len(bufs) == 1 else 0.0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("GLOO_HALO_OK" if flag.item() == 1.0 else "GLOO_HALO_BAD")
dist.destroy_process_group()
'''


def test_world_size_2_halo_exchange_over_gloo(tmp_path):
    """Two processes, gloo backend: each packs the boxes the library's halo
    plan names, sends them to its Cartesian neighbour and checks the received
    ghosts against the global field."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert "GLOO_HALO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_hostmodel_cut_paste_roundtrip():
    wl = S.small(S.WORKLOADS["isc24_small"], 21, 17, 4)
    glob = np.random.default_rng(3).standard_normal((wl.kz, wl.iy, wl.jx))
    out = np.zeros_like(glob)
    for r in range(6):
        g = make_geom(wl.jx, wl.iy, wl.kz, 1, 1, 3, 2, r)
        b = H.bounds(g, "tetav")
        loc = H.cut(glob, g, b)
        assert loc.shape == (wl.kz, b[3] - b[2] + 1, b[1] - b[0] + 1)
        H.paste(out, loc, b, H.owned(g, "tetav"))
    assert np.array_equal(out, glob)


def test_rank_local_inputs_equal_global_inputs():
    """bench.py builds each rank's arrays directly (model_inputs_local); they
    must equal the globally generated arrays cut to the rank's bounds."""
    for wl, px, py in ((S.small(S.WORKLOADS["cordex25"], 45, 41, 9, ntr=2, nspgx=6), 2, 2),
                       (S.small(S.WORKLOADS["isc24_small"], 30, 22, 8, oro="sine"), 3, 2)):
        Fg, Pg = S.model_inputs(wl)
        for r in range(px * py):
            g = make_geom(wl.jx, wl.iy, wl.kz, wl.i_band, wl.i_crm, px, py, r)
            Fl, Pl, B = S.model_inputs_local(wl, g)
            for n in Fl:
                own, b = H.owned(g, n), B[n]
                sl = (Ellipsis, slice(own[2] - b[2], own[3] - b[2] + 1), slice(own[0] - b[0], own[1] - b[0] + 1))
                assert np.array_equal(H.cut(np.asarray(Fg[n]), g, b)[sl], Fl[n][sl]), (wl.name, r, n)
            for n in Pl:
                assert np.array_equal(Pl[n], Pg[n]), n


def test_bench_reference_arm_prints_contract_line():
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--workload", "isc24_small"], capture_output=True, text=True, timeout=600)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_bench_reference_arm_keeps_the_asked_counts_or_says_what_it_cut():
    import json
    env = dict(os.environ)
    for budget, want in (("600", (3, 2)), ("0.001", (1, 1))):
        env["BENCH_CPU_BUDGET_S"] = budget
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                            "--warmup", "2", "--workload", "isc24_small"], capture_output=True, text=True, timeout=600,
                           env=env)
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert (line["steps"], line["warmup"]) == want, line
        assert ("asked: 3 after 2" in line["cpu_baseline"]["sample"]) == (want != (3, 2))

