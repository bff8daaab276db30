"""GPU tests of the pipelined physics hand-off (moloch_b200_handoff, include/moloch_b200.h): slab-wise D2H of
the state and H2D of the tendencies on two copy streams, against the plain get_field/set_field transfers.

(Sorts after the other GPU test files on purpose: written when the round's GPU budget was already spent; the
same test bodies run on the CPU against the host build of the CUDA sources, tests/test_emu_full.py.)"""
import numpy as np
import pytest

from test_gpu_parity import CASES
from util import make_gpu, make_oracle, oracle_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nslabs", [1, 3, 1000])
def test_handoff_matches_field_transfers(nslabs):
    """moloch_b200_handoff (slab-pipelined D2H of the state / H2D of the tendencies on two copy streams) moves
    exactly what get_field/set_field move, calls the physics callback once per slab in row order with that
    slab's state already on the host, and leaves the uploaded tendencies visible to status_update."""
    from regcm_b200 import hostmodel as H
    from regcm_b200.moloch import STATE_FIELDS
    wl = CASES["limited_area"]
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    m = make_gpu(wl, fields, profiles)
    m.reset_tendencies(); m.dynamical_core(); m.diagnostics()
    rng = np.random.default_rng(5)
    g = m.g

    def species(n):
        return wl.nqx if n in ("qx", "qxten") else wl.ntr if n in ("trac", "chiten") else 0

    down_items, up_items, want_down = [], [], []
    for n in STATE_FIELDS:
        box = H.bounds(g, n)
        if n == "pai":      # host array wider than the device box: only the overlap moves
            box = (box[0] - 6, box[1] + 6, box[2] - 5, box[3] + 5)
        for s in range(1, species(n) + 1) if species(n) else [0]:
            shp = (m._levels(n), box[3] - box[2] + 1, box[1] - box[0] + 1)
            a = m.pinned_empty(shp); a[...] = -7.0
            down_items.append((n, s, a, box))
            ref = np.full(shp, -7.0)
            m.get_local(n, box, s, out=ref)
            want_down.append(ref)
    for n in ["tten", "uten", "vten", "qxten", "chiten"]:
        box = H.bounds(g, n)
        for s in range(1, species(n) + 1) if species(n) else [0]:
            shp = (m._levels(n), box[3] - box[2] + 1, box[1] - box[0] + 1)
            if n == "uten":     # an array the host already owns, page-locked in place
                a = np.zeros(shp)
                m.host_register(a)
            else:
                a = m.pinned_empty(shp)
            a[...] = rng.standard_normal(a.shape) * 1e-4
            up_items.append((n, s, a, box))
    seen = []

    def physics(i1, i2):
        # slab rows of the first state array must already be on the host
        n, s, a, box = down_items[0]
        lo, hi = max(i1, box[2]) - box[2], min(i2, box[3]) - box[2]
        # a slab that lies outside this array's rows (the widened `pai` box starts the slabs below
        # row box[2] of `u`) has nothing to compare: hi < lo must not become a negative slice
        ok = True if hi < lo else bool(np.array_equal(a[:, lo:hi + 1, :], want_down[0][:, lo:hi + 1, :]))
        # (exceptions do not propagate out of a ctypes callback: record, assert below)
        seen.append((i1, i2, ok))
        return 0

    m.handoff(m.xfer_list(down_items), m.xfer_list(up_items), nslabs=nslabs, physics=physics)
    assert seen and all(seen[q][1] + 1 == seen[q + 1][0] for q in range(len(seen) - 1)), seen
    assert all(ok for _, _, ok in seen), "a slab's state had not arrived when its physics callback ran"
    assert len(seen) == min(nslabs, seen[-1][1] - seen[0][0] + 1)
    for (n, s, a, box), ref in zip(down_items, want_down):
        assert np.array_equal(a, ref), f"down {n}[{s}]"
    for n, s, a, box in up_items:
        got = m.get_local(n, box, s)
        assert np.array_equal(got.reshape(a.shape), a), f"up {n}[{s}]"
    # the uploaded tendencies are what status_update applies: same result as set_local + status_update
    m.status_update()
    t1 = m.get_global("t")
    m2 = make_gpu(wl, fields, profiles)
    m2.reset_tendencies(); m2.dynamical_core(); m2.diagnostics()
    for n, s, a, box in up_items:
        m2.set_local(n, a, box, s)
    m2.status_update()
    assert np.array_equal(t1, m2.get_global("t"))
    compare_pair = [(f, m.get_global(f), m2.get_global(f)) for f in ("u", "v", "qx", "trac", "tetav")]
    assert all(np.array_equal(a, b) for _, a, b in compare_pair)
    m.close(); m2.close()


def test_handoff_rejects_bad_arguments():
    from regcm_b200.moloch import MolochError
    wl = CASES["periodic_flat"]
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    m = make_gpu(wl, fields, profiles)
    from regcm_b200 import hostmodel as H
    box = H.bounds(m.g, "t")
    a = np.zeros((wl.kz, box[3] - box[2] + 1, box[1] - box[0] + 1))
    with pytest.raises(MolochError, match="not allocated"):
        m.handoff(m.xfer_list([("trac", 1, a, box)]), m.xfer_list([]))
    with pytest.raises(MolochError, match="species"):
        m.handoff(m.xfer_list([("qx", 9, a, box)]), m.xfer_list([]))
    with pytest.raises(MolochError, match="physics callback"):
        m.handoff(m.xfer_list([("t", 0, a, box)]), m.xfer_list([]), physics=lambda i1, i2: 1)

    def broken_physics(i1, i2):
        raise ZeroDivisionError("the host physics failed")
    up = m.pinned_empty(a.shape); up[...] = 5.0
    m.set_local("tten", np.zeros_like(a), box)
    with pytest.raises(ZeroDivisionError):      # re-raised after the C call, and nothing was uploaded
        m.handoff(m.xfer_list([("t", 0, a, box)]), m.xfer_list([("tten", 0, up, box)]), nslabs=1, physics=broken_physics)
    assert (m.get_local("tten", box) == 0.0).all()
    m.close()


def test_restart_from_the_save_set_is_bit_exact():
    """the MOLOCH restart contract (Main/mod_savefile.F90:232-242, 618-627: atm_u, atm_v, atm_w,
    atm_t, atm_pai, atm_qx, trac, ps).  Two steps, the save set staged to the host in one asynchronous batch,
    a NEW context initialised from it the way `init` does on a restart (Main/mod_init.F90:420-449 copies the
    save set, :941-953 rebuilds p, qs, rho, tvirt, tetav; ux, vx follow from u, v in `advection`), two more
    steps: every prognostic field equals the uninterrupted four-step run bit for bit."""
    from regcm_b200 import hostmodel as H
    from regcm_b200.moloch import STATE_FIELDS, STATIC_FIELDS
    from util import PROGNOSTIC
    wl = CASES["limited_area"]
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    ref = make_gpu(wl, fields, profiles)
    ref.moloch(4)
    m = make_gpu(wl, fields, profiles)
    m.moloch(2)
    save = {}
    m.set_async(True)                      # selective D2H at alarm_out_sav: enqueue everything, one sync
    for n in ("u", "v", "w", "t", "pai", "qx", "trac", "ps"):
        box = H.bounds(m.g, n)
        nspec = wl.nqx if n == "qx" else wl.ntr if n == "trac" else 0
        shp = ((nspec,) if nspec else ()) + (m._levels(n), box[3] - box[2] + 1, box[1] - box[0] + 1)
        buf = m.pinned_empty(shp)
        for s in range(nspec or 1):
            m.get_local(n, box, s + 1 if nspec else 0, out=buf[s] if nspec else buf)
        save[n] = (buf, box)
    m.sync(); m.set_async(False)
    # restart: statics as in a cold start, the save set, and init's rebuild of the derived state
    ep1 = 28.96454 / 18.01528 - 1.0
    t, pai, qv = save["t"][0], save["pai"][0], save["qx"][0][0]
    tvirt = t * (1.0 + ep1 * qv)           # Main/mod_init.F90:947-950
    derived = {"tvirt": tvirt, "tetav": tvirt / pai}
    r = type(m)(wl, lib=m.lib).allocate_moloch()
    for n in STATIC_FIELDS:
        r.set_global(n, fields[n])
    for n, (buf, box) in save.items():
        r.set_local(n, buf.reshape(buf.shape[-3:]) if n == "ps" else buf, box)
    for n, a in derived.items():
        r.set_local(n, a, save["t"][1])
    for n in ("ux", "vx", "p", "rho", "qsat"):      # rebuilt by the dycore before they are read; any finite value
        r.set_global(n, fields[n])
    for n, v in profiles.items():
        r.set_profile(n, v)
    r._chk(r.lib.moloch_b200_init(r.ctx))
    r.moloch(2)
    bad = [f for f in PROGNOSTIC + ["trac"] if not np.array_equal(ref.get_global(f), r.get_global(f))]
    assert not bad, bad
    for x in (ref, m, r):
        x.close()
