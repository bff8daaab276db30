"""Memcheck without a GPU: representative GPU-test bodies on the AddressSanitizer build of the emulated
library (tests/emu_lib.py, sanitize=True).  Run through scripts/emu_asan.sh (ASan must be preloaded into python)."""
import sys, os, ctypes as C, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import util, emu_lib
from regcm_b200 import moloch as M
lib = M.bind_library(C.CDLL(emu_lib.build(sanitize=True), mode=C.RTLD_LOCAL), 'asan')
util.LIB = lib
import test_gpu_parity as P, test_gpu_zbdy as Z, test_gpu_multi as Mu, test_gpu_zz_handoff as H
class MP:
    def setenv(self, k, v): os.environ[k] = v
def run(name, fn, *a):
    t=time.time(); fn(*a); print(name, "ok", round(time.time()-t,1), flush=True)
run("steps limited_area", P.test_steps_bit_exact, "limited_area")
run("steps periodic", P.test_steps_bit_exact, "periodic_hills")
run("bdy lam_full", Z.test_steps_with_boundary_bit_exact, "lam_full")
run("spectral", Z.test_steps_with_boundary_bit_exact, "spectral")
run("tke", Z.test_steps_with_boundary_bit_exact, "lam_tke")
run("slice", Z.test_mkslice)
run("massck", Z.test_massck_and_ps_guard)
run("diag golden", Z.test_reference_golden_more, "limited_area_diag")
run("handoff", H.test_handoff_matches_field_transfers, 3)
name, wl, px, py = Mu.CASES[4]
run("2x2 p2p", Mu.test_decomposed_bit_exact, name, wl, px, py, "p2p", MP())
run("2x2 nccl", Mu.test_decomposed_bit_exact, name, wl, px, py, "nccl", MP())
run("spectral 2x2", Mu.test_decomposed_spectral_nudging_bit_exact, 2, 2, "p2p+nccl")
os.environ["MOLOCH_B200_WSOLVE"]="6"
run("wsolve6", P.test_steps_bit_exact, "limited_area")
print("ALL OK")
