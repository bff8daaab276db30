#!/usr/bin/env python
"""Times moloch_b200_handoff alone (no dycore): state down only, tendencies up only, both; per slab count."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from regcm_b200 import synthetic as S, hostmodel as H
from regcm_b200.moloch import MolochB200, STATE_FIELDS

wl = S.WORKLOADS[os.environ.get("WL", "cordex25")]
m = MolochB200(wl).allocate_moloch()
fields, profiles, boxes = S.model_inputs_local(wl, m.g)
m.init_moloch(fields, profiles, boxes)
g = m.g
down = [n for n in STATE_FIELDS if not (n == "trac" and wl.ntr == 0)]
up = ["tten", "uten", "vten", "qxten"] + (["chiten"] if wl.ntr > 0 else [])
four_d = {"qx", "trac", "qxten", "chiten"}
hbuf, nb = {}, {"down": 0, "up": 0}
for n in down + up:
    box = H.bounds(g, n)
    nspec = wl.nqx if n in ("qx", "qxten") else wl.ntr if n in ("trac", "chiten") else 1
    shp = (nspec, m._levels(n), box[3] - box[2] + 1, box[1] - box[0] + 1)
    hbuf[n] = (m.pinned_empty(shp), box)
    hbuf[n][0][...] = 0.0
    nb["down" if n in down else "up"] += int(np.prod(shp)) * 8
xd = m.xfer_list([(n, s + 1 if n in four_d else 0, hbuf[n][0][s], hbuf[n][1]) for n in down for s in range(hbuf[n][0].shape[0])])
xu = m.xfer_list([(n, s + 1 if n in four_d else 0, hbuf[n][0][s], hbuf[n][1]) for n in up for s in range(hbuf[n][0].shape[0])])
none = m.xfer_list([])
print("bytes down %.3f GB up %.3f GB" % (nb["down"] / 1e9, nb["up"] / 1e9))
for ns in [int(x) for x in os.environ.get("SLABS", "1,2,4,8,16").split(",")]:
    row = []
    for name, a, b, nbytes in (("down", xd, none, nb["down"]), ("up", none, xu, nb["up"]), ("both", xd, xu, nb["down"] + nb["up"])):
        m.handoff(a, b, nslabs=ns); m.sync()
        t0 = time.perf_counter()
        for _ in range(3):
            m.handoff(a, b, nslabs=ns)
        m.sync()
        dt = (time.perf_counter() - t0) / 3
        row.append("%s %.2f ms (%.1f GB/s)" % (name, dt * 1e3, nbytes / dt / 1e9))
    print("slabs %2d: " % ns + "   ".join(row))
m.close()
