"""Diagnostics (GPU): cycles per tcgen05.mma for the decoder's instruction shapes (csrc/nb_tc_bench.cu)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from neuralbody_b200 import capi  # noqa: E402

lib = capi.load()
lib.nb_debug_mma_rate.restype = C.c_int
lib.nb_debug_mma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
names = {0: "1 CTA  SS M=128", 1: "1 CTA  TS M=128", 2: "pair   SS M=256", 3: "pair   TS M=256"}
for N in (256, 128, 16):
    for variant in (0, 1, 2, 3):
        for n in (32, 128):
            for rep in range(2):
                capi.check(lib.nb_debug_mma_rate(variant, n, N, out.data_ptr(), None), "nb_debug_mma_rate")
                torch.cuda.synchronize()
            issue, done = [int(v) for v in out.tolist()]
            print("%s  N=%3d  n=%3d : issue %7d cycles (%6.1f / MMA)   until commit %7d (%6.1f / MMA)" % (
                names[variant], N, n, issue, issue / n, done, done / n))
