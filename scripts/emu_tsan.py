"""Race check of the halo protocol without GPUs: decomposed runs (ranks = threads) on the ThreadSanitizer build
of the emulated library.  TSan reports accesses of different rank threads to one address that no
release/acquire flag (or NCCL mailbox lock) orders.  Run through scripts/emu_tsan.sh."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import emu_lib  # noqa: E402
import util  # noqa: E402
from regcm_b200 import moloch as M  # noqa: E402

lib = M.bind_library(C.CDLL(emu_lib.build(sanitize="thread"), mode=C.RTLD_LOCAL), 'tsan')
util.LIB = lib
import test_gpu_multi as Mu  # noqa: E402


class MP:
    def setenv(self, k, v): os.environ[k] = v


def run(name, fn, *a):
    t = time.time()
    try:
        fn(*a)
    except BaseException as exc:  # noqa: BLE001
        if type(exc).__name__ != "Skipped":     # pytest.skip inside the test body: not a representative case
            raise
        print(name, "skipped", flush=True)
        return
    print(name, "ok", round(time.time() - t, 1), flush=True)


cases = {c[0]: c for c in Mu.CASES}
sel = sys.argv[1:] or ["limited_area", "limited_area_2x2", "periodic_2x2", "band_2x4", "limited_area_1x4"]
for nm in sel:
    name, wl, px, py = cases[nm]
    for tr in ("p2p", "p2p_sound", "p2p_unfused", "nccl", "p2p_psignal", "p2p_nowz"):
        run(f"{name} {tr}", Mu.test_decomposed_bit_exact, name, wl, px, py, tr, MP())
run("boundary 2x2", Mu.test_decomposed_boundary_bit_exact, 2, 2, MP())
run("spectral 2x2", Mu.test_decomposed_spectral_nudging_bit_exact, 2, 2, "p2p+nccl")
print("DONE")
