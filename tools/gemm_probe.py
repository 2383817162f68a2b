"""A/B of the attention GEMM kernels behind tcv_gemm_tn_tc (bf16x3, fp32 out) on a B200:
    python tools/gemm_probe.py
For each shape and kernel (CTA pair 256x256 tiles / single CTA 128x256 / per-tile 128x128): max error against an fp64
reference on sampled rows, and the device time per call (CUDA events, 10 calls after 2 warm-ups)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tcvom_b200 import _cabi

L = _cabi.lib()
dev = torch.device("cuda:0")
KERNELS = (("pair 256x256 BK64", 0), ("pair 256x256 BK32", 262144), ("single 128x256", 1024), ("per-tile 128x128", 2048))


def planes(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous()


def run(M, N, K, batch, only=None):
    g = torch.Generator(device="cpu").manual_seed(1)
    A = torch.randn((batch, M, K), generator=g).to(dev)
    B = torch.randn((batch, N, K), generator=g).to(dev)
    Ap, Bp = planes(A), planes(B)
    ldc = (N + 63) // 64 * 64
    rows = torch.randint(0, M, (48,), generator=g).to(dev)
    rows[0], rows[1] = 0, M - 1
    Av = (Ap[0].double() + Ap[1].double())
    Bv = (Bp[0].double() + Bp[1].double())
    ref = torch.einsum("brk,bnk->brn", Av[:, rows], Bv)
    st = torch.cuda.current_stream().cuda_stream
    for name, flag in KERNELS:
        if only is not None and flag not in only:
            continue
        L.tcv_set_debug_flags(flag)
        C = torch.full((batch, M, ldc), float("nan"), device=dev)
        call = lambda: _cabi.check(L.tcv_gemm_tn_tc(Ap.data_ptr(), batch * M * K, Bp.data_ptr(), batch * N * K, C.data_ptr(),
                                                    M, N, K, ldc, M * ldc, batch, 3, 0, 0, st), "gemm")
        call(); call()
        torch.cuda.synchronize()
        err = float((C[:, rows, :N].double() - ref).abs().max())
        nan = bool(torch.isnan(C[:, :, :N]).any())
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.0 * batch * M * N * K / ms / 1e9
        print(f"M={M} N={N} K={K} b={batch}  {name:18s} {ms:8.3f} ms  {tf:7.1f} TFLOP/s algorithmic ({3*tf:7.1f} incl. split)"
              f"  max err {err:.2e}  unwritten={nan}", flush=True)
    L.tcv_set_debug_flags(0)


if __name__ == "__main__":
    run(600, 300, 64, 2)
    run(1000, 777, 128, 1)
    run(8160, 8349, 576, 3)
    run(8349, 512, 8384, 3)
