"""GPU: the tcgen05 building blocks in isolation (csrc/nb_tc_probe.cu) against fp32 matmuls of the
same fp16-rounded operands.  Pins the descriptor / TMEM-operand conventions the fused kernel uses."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def pack_kmajor(w):
    """(N,K) fp16 -> the packed K-major no-swizzle layout of csrc/nb_layout.h::umma_kmajor_offset."""
    N, K = w.shape
    out = torch.empty(N * K, dtype=w.dtype)
    n = torch.arange(N)[:, None].expand(N, K)
    k = torch.arange(K)[None, :].expand(N, K)
    off = ((k // 8) * (N // 8) + n // 8) * 64 + (n % 8) * 8 + (k % 8)
    out[off.reshape(-1)] = w.reshape(-1)
    return out


def run_probe(variant, seed=0):
    from neuralbody_b200 import capi
    lib = capi.load()
    g = torch.Generator().manual_seed(seed)
    a0 = (torch.randn((128, 64), generator=g)).half()
    w0 = (torch.randn((128, 64), generator=g) * 0.2).half()
    bias = torch.randn((128,), generator=g)
    b_hi = bias.half()
    b_lo = (bias - b_hi.float()).half()
    w0p = torch.zeros((128, 80), dtype=torch.float16)
    w0p[:, :64] = w0
    w0p[:, 64] = b_hi
    w0p[:, 65] = b_lo
    w1 = (torch.randn((64, 128), generator=g) * 0.2).half()
    dev = "cuda:0"
    a0d, w0d, w1d = a0.to(dev), pack_kmajor(w0p).to(dev), pack_kmajor(w1).to(dev)
    d0 = torch.zeros((128, 128), dtype=torch.float32, device=dev)
    d1 = torch.zeros((128, 64), dtype=torch.float32, device=dev)
    st = lib.nb_debug_tc_probe(a0d.data_ptr(), w0d.data_ptr(), w1d.data_ptr(), d0.data_ptr(), d1.data_ptr(),
                               variant, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    capi.check(st, "nb_debug_tc_probe")
    torch.cuda.synchronize()
    ref0 = a0.float() @ w0.float().t() + b_hi.float() + b_lo.float()
    h = torch.relu(d0.cpu()).half().float()          # the kernel rounds ITS OWN layer-0 output
    ref1 = h @ w1.float().t()
    e0 = float((d0.cpu() - ref0).abs().max())
    e1 = float((d1.cpu() - ref1).abs().max())
    ebias = float((d0.cpu() - ref0 - 0).abs().max())
    return e0, e1, float(ref0.abs().max()), float(ref1.abs().max())


def test_tc_probe_default_variant():
    """Variant 0 is the convention the fused kernel uses.  (Other variants can fault by reading
    shared memory out of bounds; probe them one per process: `python tests/test_tc_probe_gpu.py <v>`.)"""
    e0, e1, m0, m1 = run_probe(0)
    print("variant 0: max|d0-ref|=%.3e max|d1-ref|=%.3e (|ref0|max %.2f, |ref1|max %.2f)" % (e0, e1, m0, m1))
    assert e0 < 2e-4, "SS MMA / smem descriptor / bias-as-K-step mismatch: %.3e" % e0
    assert e1 < 2e-3, "TS MMA / TMEM A operand mismatch: %.3e" % e1


def run_pair_probe(seed=0):
    from neuralbody_b200 import capi
    lib = capi.load()
    g = torch.Generator().manual_seed(seed)
    a0 = torch.randn((256, 64), generator=g).half()
    w0 = (torch.randn((256, 64), generator=g) * 0.2).half()
    bias = torch.randn((256,), generator=g)
    b_hi = bias.half()
    b_lo = (bias - b_hi.float()).half()
    w0p = torch.zeros((256, 80), dtype=torch.float16)
    w0p[:, :64], w0p[:, 64], w0p[:, 65] = w0, b_hi, b_lo
    w1 = (torch.randn((144, 128), generator=g) * 0.2).half()
    dev = "cuda:0"
    w0h = torch.cat([pack_kmajor(w0p[:128]), pack_kmajor(w0p[128:])]).to(dev)       # rank r: rows [128 r, 128 r + 128)
    w1h = torch.cat([pack_kmajor(w1[:72]), pack_kmajor(w1[72:])]).to(dev)           # rank r: rows [72 r, 72 r + 72)
    a0d = a0.to(dev)
    d0 = torch.zeros((256, 256), dtype=torch.float32, device=dev)
    d1 = torch.zeros((256, 144), dtype=torch.float32, device=dev)
    capi.check(lib.nb_debug_tc_probe2(a0d.data_ptr(), w0h.data_ptr(), w1h.data_ptr(), d0.data_ptr(), d1.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "nb_debug_tc_probe2")
    torch.cuda.synchronize()
    ref0 = a0.float() @ w0.float().t() + b_hi.float() + b_lo.float()
    h = torch.relu(d0.cpu()[:, :128]).half().float()
    ref1 = h @ w1.float().t()
    ref1[:, 64:72] += h @ w1[64:72].float().t()            # N = 16 MMA: rank 0's local rows 64..71 -> columns 64..71
    ref1[:, 72:80] += h @ w1[136:144].float().t()          #             rank 1's local rows 64..71 -> columns 72..79
    return d0.cpu(), ref0, d1.cpu(), ref1


def test_tc_probe_cta_pair():
    """cta_group::2: M = 256 over a CTA pair, B split by N halves (rank 0: rows [0, N/2) -> accumulator columns [0, N/2)),
    SS and TS forms, N = 256 / 144 / 16, multicast commit, remote arrives, relayed bulk-copy completion."""
    d0, ref0, d1, ref1 = run_pair_probe()
    e0, e1 = float((d0 - ref0).abs().max()), float((d1 - ref1).abs().max())
    print("pair probe: max|d0-ref|=%.3e max|d1-ref|=%.3e; per CTA / column half d0 errors: %s" % (
        e0, e1, [["%.1e" % float((d0[r * 128:(r + 1) * 128, c * 128:(c + 1) * 128] - ref0[r * 128:(r + 1) * 128, c * 128:(c + 1) * 128]).abs().max())
                  for c in range(2)] for r in range(2)]))
    assert e0 < 2e-4, "2-CTA SS MMA / B-half mapping mismatch: %.3e" % e0
    assert e1 < 2e-3, "2-CTA TS MMA / N = 144 / N = 16 mismatch: %.3e" % e1


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    v = int(sys.argv[1])
    print("variant %d: max|d0-ref|=%.3e max|d1-ref|=%.3e (|ref0|max %.2f, |ref1|max %.2f)" % ((v,) + run_probe(v)))
