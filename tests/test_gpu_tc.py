"""GPU tests of the tcgen05 split-precision contraction against an fp64 matmul."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
native = importlib.import_module("speech-intent-recognizer_b200._native")


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 192), (6400, 1536, 1024), (275, 256, 512), (1, 128, 64),
                                   (4100, 512, 128), (6400, 768, 256), (9472, 1536, 512), (7690, 1536, 128), (9465, 1536, 64)])
# shapes with >= 32 tiles take the persistent kernel; its tile widths: 6400 x 1536 -> 176, 4100 x 512 and 6400 x 768 -> 160,
# 9472 / 9465 (ragged last tile) x 1536 -> 256, 7690 x 1536 -> 208 (tests/test_abi_cpu.py checks the choice itself)
def test_gemm_nt_split_f16_matches_fp64(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + K)
    a = torch.randn(M, K, device="cuda", generator=g) * 3.0
    a[:, ::7] *= 1e-3                                   # mixed magnitudes exercise the lo parts
    w = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    got = native.gemm_nt_split_f16(a, w, bias)
    want = (a.double() @ w.double().t() + bias.double())
    err = (got.double() - want).abs().max().item()
    scale = want.abs().max().item()
    # fp32 GEMM accuracy is ~1e-6 of the scale; a single fp16 pass would be ~1e-3
    assert err < 2e-6 * scale * (K / 64) ** 0.5 + 1e-6, (err, scale)


@pytest.mark.parametrize("cin,cout,H,W", [(32, 64, 32, 100), (64, 128, 16, 50), (128, 64, 16, 50), (64, 32, 32, 100),
                                          (64, 32, 9, 21)])
def test_tc_conv3x3_matches_conv2d(cin, cout, H, W):
    """Every instantiation of the implicit-GEMM convolution (forward shapes and data-gradient shapes)."""
    torch.manual_seed(cin + cout)
    B = 3
    x = torch.randn(B, H, W, cin, device="cuda")
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
    got = native.conv3x3_nhwc_split_f16(x, w.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous())
    want = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1)
    err = float((got.double() - want).abs().max() / want.abs().max())
    assert err < 1e-5, err          # fp32-level: K = 9 * C_in up to 1152 products per output
