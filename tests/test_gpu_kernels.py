"""Kernel-level GPU tests through the C ABI: every tensor-core kernel path (tcgen05 GEMM with 1/3/6-term
splits, conv kernel generations v1/v2/v3 incl. stride-2, transposed-conv phases, folded 8-channel input)
against an fp64 reference / the CUDA-core fp32 kernel on the same inputs."""
import ctypes as C
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def _tc():
    import tc_check
    return tc_check


@pytest.mark.parametrize("nsplit,M,N,K,batch,bf16out", [
    (1, 128, 256, 64, 1, 0), (1, 300, 520, 192, 2, 0), (1, 300, 520, 192, 2, 1),
    (3, 128, 128, 64, 1, 0), (3, 1000, 1000, 576, 2, 0), (6, 500, 300, 576, 1, 0)])
def test_gemm_tn_tc(nsplit, M, N, K, batch, bf16out):
    _tc().case_gemm(nsplit, M, N, K, batch, bf16out)


@pytest.mark.parametrize("cin,cout,h,w,n,kind", [
    (64, 128, 20, 28, 2, "3x3"),        # v2, ragged tiles
    (256, 256, 10, 12, 3, "3x3"),       # v2, K = 2304
    (128, 64, 16, 16, 2, "1x1"),        # v2, 1x1
    (64, 64, 9, 14, 2, "deconv"),       # v2, transposed-conv phase, odd sizes
    (32, 32, 16, 48, 1, "3x3"),         # v3 narrow
    (64, 32, 36, 44, 2, "3x3"),         # v3, two K blocks
    (32, 32, 18, 20, 1, "deconv"),      # v3, strided TMA-store view
    (64, 32, 20, 24, 1, "1x1"),         # v3, 1x1
    (8, 32, 40, 56, 2, "3x3"),          # v3, horizontal taps folded into K
    (64, 128, 40, 56, 2, "3x3s2"),      # v1, TMA traversal stride
    (32, 64, 36, 52, 1, "3x3s2"),
])
def test_conv_tc_matches_cuda_core_conv(cin, cout, h, w, n, kind):
    _tc().case_conv(cin, cout, h, w, n, kind)


@pytest.mark.parametrize("cin,cout,h,w,n,kind", [
    (128, 128, 136, 240, 3, "3x3"),     # N = 128 pair tiles, 3 rounds of work items
    (256, 256, 68, 120, 3, "3x3"),      # N = 256
    (512, 512, 34, 60, 3, "3x3"),       # two N tiles, K = 4608
    (256, 128, 10, 12, 3, "3x3"),       # image smaller than a pair tile (the lower CTA's tile is mostly outside)
    (128, 256, 18, 22, 2, "1x1"),       # no halo
    (256, 256, 9, 14, 2, "deconv"),     # transposed-conv phase: strided TMA-store view
    (64, 64, 40, 56, 2, "3x3"),         # N = 64: stacked [B_hi ; B_lo] operand, two MMAs per K step
    (128, 64, 272, 480, 1, "3x3"),      # the same at a full-size layer shape
    (64, 64, 9, 14, 2, "deconv"),
])
def test_conv_cta_pair_kernel(cin, cout, h, w, n, kind):
    """Wide layers on CTA pairs (conv_tc2p.cu, tcgen05.mma.cta_group::2) against fp64 torch and the CUDA-core kernel."""
    _tc().case_conv(cin, cout, h, w, n, kind, False, 4)


@pytest.mark.parametrize("cin,cout,h,w,n,prepad,flags,perm", [
    (32, 32, 40, 72, 2, False, 0, False),        # rows mode (one MMA = nine taps), ragged 64-pixel tiles
    (32, 32, 40, 72, 2, True, 0, False),         # pre-padded x (reflect layers): taps 0..2
    (8, 32, 24, 136, 2, True, 0, False),         # fold: 3 pixels x 8 channels per box row
    (8, 32, 24, 136, 2, False, 0, True),         # 8 channels, zero padding: rows mode without the fold, shuffled taps
    (32, 8, 18, 40, 3, False, 0, True),          # 8 output channels, 32 x 2 tiles
    (16, 32, 20, 24, 1, True, 0, False),         # 16 x 4 tiles
    (32, 32, 40, 72, 2, False, 1 << 20, False),  # the stacked-tap mode the rows mode replaces
    (64, 64, 20, 72, 2, False, 0, False),        # rows mode with 64-channel blocks: two accumulators
    (32, 64, 24, 136, 2, True, 0, True),         # ... 32 input channels in a 64-channel box, pre-padded, shuffled taps
    (64, 32, 18, 40, 3, False, 0, False),        # ... 32 output channels, 32 x 2 tiles
    (64, 64, 20, 72, 2, False, 1 << 21, False),  # the stacked-tap mode it replaces (64 channels)
    (128, 256, 12, 40, 2, False, 0, False),      # wide layers: one CTA per tap and 128 x 128 tile
])
def test_wgrad_nhwc_tc(cin, cout, h, w, n, prepad, flags, perm):
    """Weight gradient straight from the NHWC activations (conv_wgrad_tc.cu) against torch fp64 autograd."""
    _tc().case_wgrad(cin, cout, h, w, n, prepad, flags, perm)


@pytest.mark.parametrize("M,N,K,batch,a_mn,b_mn", [
    (1000, 520, 576, 2, False, False),     # K-major operands with padded pitches
    (1000, 520, 600, 2, False, True),      # dQ = dS . Kn : B as its producer left it; K not a multiple of 64
    (1035, 512, 1035, 2, True, True),      # dF = A2^T . dO2 : both operands MN-major, ragged M / K
    (640, 300, 520, 1, True, False),
])
def test_gemm_tc_ex_operand_layouts(M, N, K, batch, a_mn, b_mn):
    _tc().case_gemm_ex(M, N, K, batch, a_mn, b_mn)


@pytest.mark.parametrize("cin,cout,h,w,n,groups,dil", [
    (64, 64, 40, 56, 3, 3, 1),          # N = 64 stacked operand, ragged tiles
    (128, 128, 34, 60, 3, 3, 2),        # dilation 2 (FBA layer3)
    (256, 256, 18, 30, 4, 2, 1),        # N = 256; two images per statistics group (train-mode BatchNorm layout)
    (512, 512, 10, 12, 2, 2, 4),        # two N tiles, image smaller than a pair tile
])
def test_conv_epilogue_statistics(cin, cout, h, w, n, groups, dil):
    """Per-channel output statistics from the conv epilogue (tcv_conv_desc.stats) against sums over the stored output."""
    _tc().case_conv_stats(cin, cout, h, w, n, groups, dil)


def test_conv_kernel_generation_switch():
    """tcv_set_conv_tc_version selects the kernel generation; all generations agree."""
    from tcvom_b200 import _cabi
    L = _cabi.lib()
    try:
        for v in (1, 2, 3):
            L.tcv_set_conv_tc_version(v)
            _tc().case_conv(32, 32, 16, 48, 1, "3x3")
    finally:
        L.tcv_set_conv_tc_version(3)
