#!/usr/bin/env python
"""bench.py -- MOLOCH dycore throughput (cell-updates/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one full MOLOCH model step (reset_tendencies + dynamical_core +
diagnostics + status_update with zero physics tendencies) of the named
synthetic workload; value = jx*iy*kz*K / device time (max over ranks).  N > 1
is a STRONG-scaling run: the same global grid is split with RegCM's own 2-D
block decomposition (set_nproc) and halos travel over NCCL/NVLink.

The JSON line also carries
  e2e          the same metric through the reference-facing `moloch` hand-off:
               every step the physics-facing state is copied device->host and
               the tendencies host->device (pinned buffers), inside the timing
  roofline     dominant kernel: algorithmic bytes / CUDA-event duration vs the
               measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline the CPU oracle (C++ restatement of mod_moloch.F90) on the host
               cores of this box, bounded sample (rank 0, N=1 only)
`--impl reference` times only that CPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from regcm_b200 import synthetic as S  # noqa: E402

DEFAULT_WORKLOAD = "cordex25"

# algorithmic doubles per cell per launch; per-field kernels
# are multiplied by the number of fields in the launch
ALG_DOUBLES = {
    "sound_pre": 9, "divdamp_filter": 7, "wsolve": 12, "uvupdate": 10, "sfinish": 4, "tetavf_init": 2,
    "destagger": 6, "waf_vertical": 5, "waf_meridional": 6, "waf_zonal": 6, "waf_horizontal": 12, "curvature": 4, "restagger": 6,
    "tvirt_temp": 10, "diag_prq": 5, "status_update": None, "reset_tendencies": None,
}
PER_FIELD = {"waf_vertical", "waf_meridional", "waf_zonal", "waf_horizontal"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a
    thread (10 ms period, started before the warm-up so that it is running
    when the short timed region begins); nvidia-smi as fall-back."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
               "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}

    def __init__(self, index: int):
        self.index, self.rows, self.stop_flag, self.thread = index, [], False, None
        self.window = [None, None]
        self.max_mhz = None

    def _loop_nvml(self):
        import pynvml as N
        N.nvmlInit()
        uuid = None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids):
                if ids[self.index].isdigit():
                    idx = int(ids[self.index])
                else:
                    uuid = ids[self.index]
        h = N.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid) if uuid \
            else N.nvmlDeviceGetHandleByIndex(idx)
        self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        while not self.stop_flag:
            try:
                mhz = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                try:
                    rs = int(N.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    rs = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def _loop_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in out.stdout.strip().split(",")]
                self.max_mhz = float(f[1])
                self.rows.append((time.perf_counter(), float(f[0]), int(f[2], 16)))
            except Exception:
                time.sleep(0.05)

    def _loop(self):
        try:
            self._loop_nvml()
        except Exception:
            self._loop_smi()

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def mark_begin(self):
        self.window[0] = time.perf_counter()

    def mark_end(self):
        self.window[1] = time.perf_counter()

    def stop(self) -> dict:
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=3)
        t0, t1 = self.window
        rows = [r for r in self.rows if t0 is not None and t1 is not None and t0 <= r[0] <= t1]
        note = "samples inside the timed region"
        if not rows and self.rows and t0 is not None:
            # region shorter than one sampling period: nearest samples around it
            rows = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:2]
            note = "timed region shorter than the sampling period: nearest samples"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples (NVML and nvidia-smi unavailable)"]}
        bits = 0
        for r in rows:
            bits |= r[2]
        reasons = sorted(n for n, b in self.REASONS.items() if bits & b)
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(rows), "note": note}


def cpu_sample_workload(wl):
    """Bounded CPU sample: the workload itself while a step of it stays in the seconds on the host
    cores (cordex25: the full 400x400x41 grid, same config as the GPU arm); the larger grids
    (cp3km, tracer40) cropped horizontally, same species, sponge, dt, dx."""
    if wl.cells * wl.nfields <= 8_000_000 * 20:
        return wl
    n = 256
    return S.small(wl, min(wl.jx, n), min(wl.iy, n), wl.kz)


def host_cores() -> int:
    """Cores this process may run on.  torch.distributed.run exports OMP_NUM_THREADS=1 to its
    workers: the CPU arm must not inherit that, it sets the OpenMP team size itself."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_cpu_reference(wl, steps: int, warmup: int, budget_s: float | None = None) -> dict:
    """The CPU arm: oracle (C++ restatement of Main/mod_moloch.F90, OpenMP over
    all host cores) on a bounded sample of the workload.  `budget_s`: wall-clock budget of the whole
    call; the first warm-up step is timed and the step count is cut (never below 1) when
    `warmup + steps` steps would not fit -- the record names the counts that were run."""
    from oracle.oracle import Oracle
    swl = cpu_sample_workload(wl)
    o = Oracle(swl)
    o.load_primary(S.make_primary(swl))
    o.set_threads(host_cores())     # omp_set_num_threads: overrides an inherited OMP_NUM_THREADS
    cores = o.get_threads()
    warmup = max(warmup, 1)
    t0 = time.perf_counter()
    o.step(1)
    t1 = time.perf_counter() - t0
    asked = (steps, warmup)
    if budget_s is not None and (warmup + steps) * t1 > budget_s:
        warmup = min(warmup, 3 if 4 * t1 <= budget_s else 1)
        steps = max(1, min(steps, int(budget_s / t1) - warmup))
    if warmup > 1:
        o.step(warmup - 1)
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    ok = bool(np.isfinite(o.get("pai")).all())
    cut = "" if (steps, warmup) == asked else f" (asked: {asked[0]} after {asked[1]}; cut to the {budget_s:.0f} s budget)"
    return {"value": swl.cells * steps / dt, "unit": "cell-updates/s", "cores": cores, "kind": "port",
            "sample": f"{wl.name} " + ("full grid " if swl is wl else "cropped to ") +
                      f"{swl.jx}x{swl.iy}x{swl.kz} (F={swl.nfields}), {steps} steps "
                      f"after {warmup} warm-up{cut}, {dt:.2f} s on {cores} OpenMP threads; finite={ok}",
            "ms_per_step": dt / steps * 1e3, "grid": [swl.jx, swl.iy, swl.kz], "steps": steps, "warmup": warmup}


def measure_e2e(m, wl, ksteps, barrier=lambda: None, allreduce_max=lambda x: x):
    """The metric through the reference-facing hand-off with HOST buffers: every step the state the
    physics reads goes device -> host and the tendencies host -> device (pinned memory), inside the
    timed region.  `m`: an initialised MolochB200; wall-clock timing between two barriers."""
    from regcm_b200 import hostmodel as H
    from regcm_b200.moloch import STATE_FIELDS
    g = m.g
    down = [n for n in STATE_FIELDS if not (n == "trac" and wl.ntr == 0)]
    up = ["tten", "uten", "vten", "qxten"] + (["chiten"] if wl.ntr > 0 else [])
    hbuf, bytes_d2h, bytes_h2d = {}, 0, 0
    for n in down + up:
        box = H.bounds(g, n)
        nk = m._levels(n)
        nspec = wl.nqx if n in ("qx", "qxten") else wl.ntr if n in ("trac", "chiten") else 1
        shp = (nspec, nk, box[3] - box[2] + 1, box[1] - box[0] + 1)
        hbuf[n] = (m.pinned_empty(shp), box)
        hbuf[n][0][...] = 0.0
        if n in down:
            bytes_d2h += int(np.prod(shp)) * 8
        else:
            bytes_h2d += int(np.prod(shp)) * 8

    def e2e_step_sequential():
        m.reset_tendencies()
        m.dynamical_core()
        m.diagnostics()
        m.set_async(True)       # batch the hand-off: one sync per direction
        for n in down:          # device -> host: what mkslice/physics read
            buf, box = hbuf[n]
            for s in range(buf.shape[0]):
                m.get_local(n, box, s + 1 if n in ("qx", "trac") else 0, out=buf[s])
        m.sync()                # the host physics would run here, on the downloaded state
        for n in up:            # host -> device: the physics tendencies
            buf, box = hbuf[n]
            for s in range(buf.shape[0]):
                m.set_local(n, buf[s], box, s + 1 if n in ("qxten", "chiten") else 0)
        m.set_async(False)
        m.status_update()

    # Pipelined hand-off (moloch_b200_handoff): the rank's rows in slabs, state down on one copy
    # stream, the tendencies of a slab up on a second one as soon as that slab's state (and the
    # column physics on it -- none here) is done: both directions of the link overlap.
    nslabs = int(os.environ.get("BENCH_HANDOFF_SLABS", "8"))
    four_d = {"qx", "trac", "qxten", "chiten"}
    xl_down = m.xfer_list([(n, s + 1 if n in four_d else 0, hbuf[n][0][s], hbuf[n][1])
                           for n in down for s in range(hbuf[n][0].shape[0])])
    xl_up = m.xfer_list([(n, s + 1 if n in four_d else 0, hbuf[n][0][s], hbuf[n][1])
                         for n in up for s in range(hbuf[n][0].shape[0])])

    def e2e_step_pipelined():
        m.reset_tendencies()
        m.dynamical_core()
        m.diagnostics()
        m.handoff(xl_down, xl_up, nslabs=nslabs)
        m.status_update()

    # untimed self-check of the pipelined path against the plain field transfers
    handoff_mode, handoff_note = "pipelined", None
    try:
        m.reset_tendencies(); m.dynamical_core(); m.diagnostics()
        hbuf["tten"][0][...] = 1.0e-5
        m.handoff(xl_down, xl_up, nslabs=nslabs)
        ok = True
        for n in ("pai", "t", "qx", "ux"):
            buf, box = hbuf[n]
            for s in range(buf.shape[0]):
                ok = ok and bool(np.array_equal(m.get_local(n, box, s + 1 if n in four_d else 0).reshape(buf[s].shape),
                                                buf[s]))
        buf, box = hbuf["tten"]
        ok = ok and bool(np.array_equal(m.get_local("tten", box).reshape(buf[0].shape), buf[0]))
        hbuf["tten"][0][...] = 0.0
        m.handoff(xl_down, xl_up, nslabs=nslabs)
        ok = ok and bool((m.get_local("tten", box) == 0.0).all())
        m.status_update()
        if not ok:
            handoff_mode, handoff_note = "sequential", "pipelined hand-off failed its self-check"
    except Exception as exc:  # noqa: BLE001
        hbuf["tten"][0][...] = 0.0
        handoff_mode, handoff_note = "sequential", f"pipelined hand-off unavailable: {exc}"
    if os.environ.get("BENCH_HANDOFF", "") == "sequential":
        handoff_mode, handoff_note = "sequential", "BENCH_HANDOFF=sequential"
    def run(step):
        step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            step()
        m.sync()
        barrier()
        return allreduce_max(time.perf_counter() - t0)

    # both hand-offs are timed (a few steps each); the line reports the faster one and names it
    times = {"sequential": run(e2e_step_sequential)}
    if handoff_mode == "pipelined":
        times["pipelined"] = run(e2e_step_pipelined)
    handoff_mode = min(times, key=times.get)
    dt = times[handoff_mode]
    return {"value": wl.cells * ksteps / dt, "unit": "cell-updates/s",
           "h2d_bytes_per_step": bytes_h2d, "d2h_bytes_per_step": bytes_d2h, "steps": ksteps,
           "ms_per_step": dt / ksteps * 1e3,
           "handoff": handoff_mode, "handoff_slabs": nslabs if handoff_mode == "pipelined" else 1,
           "ms_per_step_by_handoff": {k: v / ksteps * 1e3 for k, v in times.items()},
           "handoff_note": handoff_note,
           # filled in by main(): the link rate over the hand-off window (e2e step minus the device-resident step)
           "link_gbs": None,
           "what": "moloch(): device dycore, D2H of u,v,w,ux,vx,pai,tetav,t,tvirt,p,rho,qsat,ps,qx,trac to "
                   "pinned host arrays, H2D of tten,uten,vten,qxten,chiten, device status_update; pipelined: "
                   "rows in slabs, both copy directions overlapped (moloch_b200_handoff)"}



WSOLVE_SMALL = 2            # CTA-parallel coefficients + one-warp sweeps out of shared memory: slower than the thread-per-
                            # column variants on a full GPU (13.3 vs 12.1 ms/step, r2a) but not bound by the latency of one
                            # column sweep per warp: a candidate where a rank holds few columns (strong scaling, config 2)
WSOLVE_SMALL_COLUMNS = 60000
WSOLVE_NEW = (8, 9, 12)     # round-2 variants (row tiles, cp.async.cg ring, re-partitioned upward ring; 12: sweeps in TMEM)


def variants_agree(wl, device, lib=None) -> bool:
    """The round-2 variants of the column solver against variant 5 on a small case of the same configuration (same
    species, boundary type, level count): two steps, every prognostic field bit for bit."""
    from regcm_b200.moloch import MolochB200
    swl = S.small(wl, min(wl.jx, 72), min(wl.iy, 56), wl.kz)
    out = []
    for v in (5,) + WSOLVE_NEW + (WSOLVE_SMALL,):
        m = MolochB200(swl, device=device, lib=lib).allocate_moloch()
        fields, profiles, boxes = S.model_inputs_local(swl, m.g)
        m.init_moloch(fields, profiles, boxes)
        m.set_option("wsolve", v)
        m.moloch(2)
        out.append([m.get_local(n) for n in ("u", "v", "w", "pai", "tetav", "t", "qx")])
        m.close()
    return all(np.array_equal(a, b) for other in out[1:] for a, b in zip(out[0], other))


def autotune_wsolve(m, wl, timed, all_min=lambda x: x, device=-1, lib=None):
    """(variant, record): times two steps of the benchmark model per admissible variant and picks the faster.
    `all_min`: minimum over the ranks (every rank must take the same decisions)."""
    g = m.g
    small = (g.jce2 - g.jce1 + 1) * (g.ice2 - g.ice1 + 1) <= WSOLVE_SMALL_COLUMNS
    rec = {"candidates": [5], "rejected": {"2": "time on a full GPU (13.3 vs 12.1 ms/step, r2a); a candidate again when a "
                                                 f"rank holds <= {WSOLVE_SMALL_COLUMNS} columns", "6": "time (246 vs 217 us, r2a)",
                                            "7": "time (248 vs 217 us, r2a)", "10": "time (246 vs 201 us, r2ws)",
                                            "11": "time (168 vs 158 us for 12, r2ws3)", "13": "time (160 vs 158 us, r2ws4)"}}
    try:
        ok = bool(variants_agree(wl, device, lib))
    except Exception as exc:  # noqa: BLE001
        ok = False
        rec["note"] = f"variants {WSOLVE_NEW} not considered: {exc}"
    rec["new_bit_exact_vs_v5"] = bool(all_min(1.0 if ok else 0.0) > 0.5)
    if rec["new_bit_exact_vs_v5"]:
        rec["candidates"] += list(WSOLVE_NEW) + ([WSOLVE_SMALL] if small else [])
    rec["ms_per_step"] = {}
    for v in rec["candidates"]:
        m.set_option("wsolve", v)
        m.moloch(1)
        m.sync()
        # max over ranks; the better of two short timings (a neighbour process starting up must not decide)
        reps = max(1, int(os.environ.get("BENCH_TUNE_REPS", "2")))
        rec["ms_per_step"][str(v)] = min(timed(lambda: m.moloch(1), 2) for _ in range(reps)) / 2.0
    best = min(rec["candidates"], key=lambda v: rec["ms_per_step"][str(v)])
    return best, rec


# ---- parity of the path that is about to be timed ------------------------------------------
# Before the timed region every rank runs PARITY_STEPS steps of a reduced cordex25 (same levels,
# species, sponge, dt, dx; 64x64 columns) on the SAME decomposition, transport, halo fusion level
# and kernel variants the timed region uses; rank 0 gathers the owned cells, hashes the prognostic
# fields (SHA-256) and compares with tests/golden/bench_parity.json -- digests the CPU oracle
# produced from the same inputs (scripts/make_bench_golden.py).  A mismatch fails the run.
PARITY_STEPS = 2
PARITY_FIELDS = ["u", "v", "w", "pai", "tetav", "t", "qx", "trac", "ux", "vx"]


def parity_workload():
    return S.small(S.WORKLOADS["cordex25"], 64, 64, S.WORKLOADS["cordex25"].kz)


def parity_descriptor(wl):
    return {"base": "cordex25", "grid": [wl.jx, wl.iy, wl.kz], "nqx": wl.nqx, "ntr": wl.ntr, "nspgx": wl.nspgx,
            "dt_s": wl.dt, "dx_m": wl.dx, "advected_fields": wl.nfields}


def digest(a) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def run_parity(make_model, gather, rank: int) -> dict | None:
    """`make_model(wl)`: an initialised MolochB200 of this rank configured like the timed model;
    `gather(obj)`: list of every rank's obj on rank 0 (None elsewhere).  Returns the record on rank 0."""
    from regcm_b200 import hostmodel as H
    wl = parity_workload()
    gold_path = os.path.join(ROOT, "tests", "golden", "bench_parity.json")
    gold = json.load(open(gold_path))
    m, inputs = make_model(wl)
    m.moloch(PARITY_STEPS)
    m.sync()
    mine = {}
    for n in PARITY_FIELDS:
        own = H.owned(m.g, n)
        mine[n] = (own, m.get_local(n, own))
    shape = {n: m.global_shape(n) for n in PARITY_FIELDS}
    inputs_ok = all(gold["inputs"].get(n) == d for n, d in inputs.items()) if inputs is not None else None
    m.close()
    parts = gather((mine, inputs_ok))
    if rank != 0:
        return None
    got = {}
    for n in PARITY_FIELDS:
        glob = np.zeros(shape[n])
        for part, _ in parts:
            own, loc = part[n]
            H.paste(glob, loc, own, own)
        got[n] = digest(glob)
    bad = [n for n in PARITY_FIELDS if got[n] != gold["fields"][n]]
    ins = [ok for _, ok in parts if ok is not None]
    return {"n_ranks": len(parts), "case": parity_descriptor(wl), "steps": PARITY_STEPS, "fields": PARITY_FIELDS,
            "bit_exact": not bad, "mismatch": bad, "inputs_match_golden": (all(ins) if ins else None),
            "golden": "tests/golden/bench_parity.json (CPU oracle, scripts/make_bench_golden.py)",
            "how": "SHA-256 of the gathered global fields"}


def base_line(wl, args, n_gpus):
    return {"metric": "MOLOCH dycore cell-updates/s", "unit": "cell-updates/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "grid": [wl.jx, wl.iy, wl.kz], "advected_fields": wl.nfields,
                       "nqx": wl.nqx, "ntr": wl.ntr, "mo_nadv": wl.mo_nadv, "mo_nsound": wl.mo_nsound,
                       "dt_s": wl.dt, "dx_m": wl.dx, "periodic": bool(wl.i_crm),
                       "bytes_per_cell_update": wl.bytes_per_cell_update(),
                       "l2": "state (>1 GB) exceeds the 126 MB L2; no explicit flush"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(S.WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"],
                    help="strict: no FMA contraction, bit-identical to the reference arithmetic (the default and the "
                         "parity build); fast: the FMA-contracting build of the same sources (within the tolerances "
                         "of tests/test_gpu_zz_fastmode.py)")
    ap.add_argument("--transport", default=os.environ.get("BENCH_TRANSPORT", "p2p"), choices=["p2p", "nccl"])
    ap.add_argument("--px", type=int, default=int(os.environ.get("BENCH_PX", "0")))
    ap.add_argument("--py", type=int, default=int(os.environ.get("BENCH_PY", "0")))
    args = ap.parse_args()
    wl = S.WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the same --steps K --warmup W as the GPU arm while the whole run stays within a few minutes of host time
        # (cordex25: 1.3 s per step on 16 cores); a slower host gets fewer steps, and the record says so
        r = run_cpu_reference(wl, max(1, args.steps), args.warmup,
                              budget_s=float(os.environ.get("BENCH_CPU_BUDGET_S", "150")))
        line = base_line(wl, args, args.gpus)
        line.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                     "warmup": r["warmup"],
                     "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                     "e2e": {"value": r["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0},
                     "gpu_launches": 0,
                     "note": "reference Fortran/MPI build unavailable (no Fortran compiler, MPI or NetCDF in the "
                             "image); this is the C++ restatement of Main/mod_moloch.F90 on the host cores"})
        line["config"]["cpu_sample_grid"] = r["grid"]
        print(json.dumps(line), flush=True)
        return 0

    if args.mode == "fast":      # chosen before the binding module looks for its library
        os.environ["MOLOCH_B200_LIB"] = os.path.join(ROOT, "regcm_b200", "libmoloch_b200_fast.so")
        args.no_parity = True    # not bit-identical by construction: its check is the tolerance test
    import torch
    import torch.distributed as dist
    from regcm_b200.moloch import MolochB200, STATE_FIELDS
    from regcm_b200 import hostmodel as H

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # Decomposition: RegCM lets the user fix it (njxcpus/niycpus in &dimparam); its
    # automatic choice (set_nproc) is the most square one.  On NVLink the 1-D split
    # along i measures faster (contiguous rows only, two neighbours), so that is the
    # bench default; --px/--py (or BENCH_PX/BENCH_PY) select any other, e.g. 2x4.
    px = args.px or (1 if world > 1 else None)
    py = args.py or (world if world > 1 else None)
    if world > 1 and wl.iy // py < 3:
        px, py = None, None     # too thin: fall back to set_nproc's choice
    stream = torch.cuda.Stream()

    def build_model(wl=wl, global_inputs=False):
        m = MolochB200(wl, rank=rank, nranks=world, px=px, py=py, device=local_rank).allocate_moloch()
        if world > 1:
            if args.transport == "nccl":
                ids = [MolochB200.comm_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                m.comm_init(ids[0])
            else:   # direct NVLink peer stores between the ranks' arenas (CUDA IPC)
                blobs = [None] * world
                dist.all_gather_object(blobs, m.p2p_export())
                m.p2p_connect(blobs)
                dist.barrier()
        m.set_stream(stream.cuda_stream)
        if global_inputs:   # the small parity case: global arrays, cut to the rank by init_moloch
            F, prof = S.model_inputs(wl)
            m.init_moloch(F, prof)
            return m, {n: digest(a) for n, a in {**F, **prof}.items()}
        # the rank's own arrays, generated locally (no global 3-D arrays: the large
        # workloads would not fit the host otherwise)
        fields, profiles, boxes = S.model_inputs_local(wl, m.g)
        m.init_moloch(fields, profiles, boxes)
        return m

    # Peer-store transport: how many of the exchanges are fused into the kernels around them is a library
    # switch (MOLOCH_B200_FUSE_HALO: 2 = all, the default; 1 = sub-steps 2.. of the sound loop only, the
    # configuration the round-1 scaling numbers were measured with; 0 = none).  One untimed step checks that
    # every rank gets through its waits and stays finite; otherwise the next lower level is used and named.
    from regcm_b200.moloch import MolochError
    levels = [os.environ.get("MOLOCH_B200_FUSE_HALO", "2")]
    if world > 1 and args.transport == "p2p":
        levels += [x for x in ("1", "0") if x < levels[0]]
    m, fusion_note = None, None
    for lv in levels:
        os.environ["MOLOCH_B200_FUSE_HALO"] = lv
        m = build_model()
        ok = 1.0
        try:
            m.moloch(1)
            m.sync()
            ok = 1.0 if bool(np.isfinite(m.get_local("pai")).all()) else 0.0
        except MolochError as exc:
            ok, fusion_note = 0.0, f"level {lv}: {exc}"
        t = torch.tensor([ok], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if float(t.item()) > 0.5:
            break
        fusion_note = fusion_note or f"level {lv}: non-finite state or a timed-out wait on another rank"
        if lv != levels[-1]:
            m.close()
            m = None
            if world > 1:
                dist.barrier()
    halo_fusion = lv

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(k):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    # ---- kernel variants: measured, not guessed ---------------------------------------
    # The implicit-w column solver exists in several bit-identical variants (include/moloch_b200.h: set_option
    # "wsolve").  The candidates are first checked against variant 5 on a small case on THIS device (every
    # prognostic field bit for bit), then timed on the benchmark model; the fastest runs the timed region and the
    # record names every variant that was rejected earlier with the reason (config.variant_tuning).
    # MOLOCH_B200_WSOLVE fixes the variant instead.
    tuning = {"wsolve": {}}
    wsolve_variant = int(os.environ.get("MOLOCH_B200_WSOLVE", "0"))
    if wsolve_variant == 0:
        def all_min(x):
            t = torch.tensor([float(x)], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return float(t.item())
        wsolve_variant, tuning["wsolve"] = autotune_wsolve(m, wl, timed, all_min, local_rank)
    m.set_option("wsolve", wsolve_variant)
    # Same for the halo fusion level at N > 1: level 2 (first sub-step, advection, wz and status_update exchanges
    # fused too) runs the timed region only if it is not slower than level 1 here (they tie at 2 GPUs, level 2
    # wins at 4 and 8).
    if world > 1 and args.transport == "p2p" and int(halo_fusion) == 2 and "MOLOCH_B200_FUSE_HALO_FIXED" not in os.environ:
        tuning["fuse_halo"] = {}
        for lv in (2, 1):
            m.set_option("fuse_halo", lv)
            m.moloch(1)
            m.sync()
            tuning["fuse_halo"][str(lv)] = timed(lambda: m.moloch(1), 3) / 3.0
        halo_fusion = min((2, 1), key=lambda lv: tuning["fuse_halo"][str(lv)])
        m.set_option("fuse_halo", int(halo_fusion))
        m.moloch(1)
        m.sync()

    # ---- parity of exactly this configuration (untimed) ------------------------------
    def make_parity_model(pwl):
        pm, ins = build_model(pwl, global_inputs=True)
        pm.set_option("wsolve", wsolve_variant)
        if world > 1 and args.transport == "p2p":
            pm.set_option("fuse_halo", int(halo_fusion))
        return pm, ins

    def gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out if rank == 0 else None
    parity = None
    if not args.no_parity:
        parity = run_parity(make_parity_model, gather, rank)
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput -----------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m.moloch(args.warmup)
    m.sync()
    m.launch_count(reset=True)
    sampler.mark_begin()
    ms = timed(lambda: m.moloch(1), args.steps)
    sampler.mark_end()
    launches = m.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    value = wl.cells * args.steps / (ms * 1e-3)

    # ---- per-kernel device timing (CUDA events on the launching stream) ----------
    m.profile_enable(True)
    psteps = max(1, min(3, args.steps))
    m.moloch(psteps)
    prof = m.profile_read()
    m.profile_enable(False)
    peak, peak_src = measured_peak_gbs()
    g = m.g
    cells_local = (g.jce2 - g.jce1 + 1) * (g.ice2 - g.ice1 + 1) * wl.kz
    kernels = []
    for name, r in prof.items():
        avg_ms = r["ms"] / max(r["launches"], 1)
        d = ALG_DOUBLES.get(name)
        if name == "status_update":
            d = 9 + 3 * (wl.nqx + wl.ntr) + 7 + 7
        alg = None
        if d is not None:
            alg = d * 8 * cells_local * (wl.nfields if name in PER_FIELD else 1)
        kernels.append({"kernel": name, "launches_per_step": r["launches"] / psteps, "avg_ms": avg_ms,
                        "share": None, "alg_bytes": alg,
                        "gbs": (alg / (avg_ms * 1e-3) / 1e9) if alg and avg_ms > 0 else None})
    tot = sum(k["avg_ms"] * k["launches_per_step"] for k in kernels) or 1.0
    for k in kernels:
        k["share"] = k["avg_ms"] * k["launches_per_step"] / tot
    kernels.sort(key=lambda k: -k["share"])
    # the dominant kernel; kernels whose shares lie within 2 % of the step of the largest count as tied (the two
    # WAF kernels do: 25.9 % each) and the one with the LOWER fraction of the peak is reported
    top = next((k for k in kernels if k["gbs"]), None)
    dom = top and min((k for k in kernels if k["gbs"] and top["share"] - k["share"] <= 0.02), key=lambda k: k["gbs"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and world == 1 and os.path.exists(tpath):   # ncu figures are per launch on the whole grid
        try:
            traffic = json.load(open(tpath)).get(wl.name, {}).get(dom["kernel"])
        except Exception:
            traffic = None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": "moloch_" + dom["kernel"], "achieved": dom["gbs"], "peak": peak,
                    "unit": "GB/s", "frac": dom["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": dom["alg_bytes"], "avg_launch_ms": dom["avg_ms"],
                    "share_of_step": dom["share"],
                    # `achieved` counts the ALGORITHMIC bytes of the reference loops the kernel replaces
                    # (SURVEY.md App. D); a fused kernel moves fewer (`traffic`, ncu) and can exceed 1.0
                    "traffic_gbs": (traffic / (dom["avg_ms"] * 1e-3) / 1e9) if traffic and world == 1 else None,
                    "note": "fused kernel: frac is against the unfused loops' algorithmic bytes; the kernel itself "
                            "is bound by FP64 instruction issue (profiles/), not by HBM",
                    "whole_step": {"achieved": wl.bytes_per_cell_update() * value / 1e9,
                                   "frac": wl.bytes_per_cell_update() * value / 1e9 / peak / world}}

    # ---- end to end through the reference-facing hand-off --------------------------
    e2e = None
    if not args.no_e2e:
        def allmax(x):
            t = torch.tensor([x], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        e2e = measure_e2e(m, wl, max(1, min(args.steps, 5)), barrier, allmax)
        win = (e2e["ms_per_step"] - ms / args.steps) * 1e-3     # per rank: bytes are this rank's
        if win > 0:
            e2e["link_gbs"] = {"d2h": e2e["d2h_bytes_per_step"] / win / 1e9, "h2d": e2e["h2d_bytes_per_step"] / win / 1e9,
                               "window_ms": win * 1e3,
                               "note": "bytes of one direction / (e2e step - device-resident step); both directions "
                                       "run inside the same window when the hand-off is pipelined"}

    finite = bool(np.isfinite(m.get_local("pai")).all())
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu_reference(wl, 4, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = base_line(wl, args, world)
        line["config"]["decomposition"] = f"{m.g.px}x{m.g.py}"
        line["config"]["halo_transport"] = ("none" if world == 1 else args.transport)
        line["config"]["arith_mode"] = ("strict (-fmad=false: bit-identical to the reference's operation order)"
                                        if args.mode == "strict" else
                                        "fast (-fmad=true; tolerances of tests/test_gpu_zz_fastmode.py)")
        line["config"]["wsolve_variant"] = wsolve_variant
        line["config"]["variant_tuning"] = tuning
        if world > 1 and args.transport == "p2p":
            line["config"]["halo_fusion_level"] = int(halo_fusion)
            line["config"]["halo_signal"] = ("producer's last edge CTA" if os.environ.get("MOLOCH_B200_PSIGNAL", "0") == "1"
                                             else "consumer's first CTA")
            line["config"]["halo_wz_fused"] = (os.environ.get("MOLOCH_B200_FUSE_WZ", "1") != "0" and int(halo_fusion) >= 2
                                               and m.g.px == 1)
            if fusion_note:
                line["config"]["halo_fusion_note"] = fusion_note
        line.update({"value": value, "ms_per_step": ms / args.steps, "clocks": clocks, "e2e": e2e,
                     "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
                     "kernels": kernels[:12], "finite": finite, "device_bytes": m.device_bytes()})
        print(json.dumps(line), flush=True)
    m.close()
    bad = torch.tensor([1.0 if (parity is not None and not parity["bit_exact"]) else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
    if float(bad.item()) > 0.5:
        if rank == 0:
            print("bench.py: the parity case does NOT match the oracle's digests: " + ", ".join(parity["mismatch"]),
                  file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
