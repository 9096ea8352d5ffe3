"""GPU parity of the tensor-core conv kernel family (through the C ABI test hook) against a plain
PyTorch fp32 convolution on the same bf16-rounded inputs and weights.

Tolerance: inputs/weights are identical bf16 values on both sides and both accumulate in fp32, so the
only differences are summation order and the final bf16 rounding of the output: |err| <= 2^-8 * |ref| +
1e-3 (one bf16 ulp of headroom plus accumulation-order noise).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref_conv(x_bf16_nhwc, cin, weight, bias, slope=None, prelu=None):
    x = x_bf16_nhwc[..., :cin].float().permute(0, 3, 1, 2).contiguous()
    w = weight.to(x_bf16_nhwc.dtype).float()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = F.conv2d(x.double(), w.double().to(x.device), bias.double().to(x.device), padding=1).float()
    if prelu is not None:
        a = prelu.to(x.device).view(1, -1, 1, 1)
        y = torch.where(y > 0, y, y * a)
    elif slope is not None:
        y = torch.where(y > 0, y, y * slope)
    return y.permute(0, 2, 3, 1).contiguous()


CASES = [
    # n, h, w, pitch, cin, cout, force_th, max_ctas
    (1, 8, 128, 64, 64, 64, 0, 0),
    (1, 8, 128, 192, 64, 32, 0, 0),
    (1, 20, 200, 192, 96, 32, 0, 0),
    (1, 20, 200, 192, 128, 32, 3, 0),
    (2, 37, 300, 192, 160, 32, 0, 0),
    (1, 37, 300, 192, 192, 64, 0, 0),
    (1, 33, 130, 192, 192, 64, 1, 0),
    (2, 40, 257, 192, 160, 32, 5, 3),   # few CTAs -> many tiles per persistent CTA
    (1, 64, 256, 64, 64, 64, 0, 2),
    (1, 5, 17, 192, 96, 32, 0, 0),      # tiny ragged frame
]


@pytest.mark.parametrize("n,h,w,pitch,cin,cout,force_th,max_ctas", CASES)
def test_conv_matches_fp32_reference(native_lib, n, h, w, pitch, cin, cout, force_th, max_ctas):
    from framewright_b200.engine import debug_conv3x3

    g = torch.Generator().manual_seed(1000 + h * 7 + w + cin + cout)
    x = (torch.randn(n, h, w, pitch, generator=g) * 0.5).to(torch.bfloat16).cuda()
    weight = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    out_pitch = 192 if cout == 32 else 64
    choff = 64 if cout == 32 and pitch == 192 and cin <= 128 else 0
    if cout == 32:
        choff = min(cin, 160)
    out = torch.full((n, h, w, out_pitch), 7.0, dtype=torch.bfloat16, device="cuda")
    debug_conv3x3(x, cin, weight, bias, out, out_choff=choff, slope=0.2, force_th=force_th, max_ctas=max_ctas)
    torch.cuda.synchronize()
    ref = _ref_conv(x, cin, weight, bias, slope=0.2)
    got = out[..., choff:choff + cout].float()
    err = (got - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-3
    bad = (err > tol)
    assert not bad.any(), f"{int(bad.sum())} / {bad.numel()} mismatches, max err {float(err.max()):.4g}"
    # channels outside the written slice are untouched
    mask = torch.ones(out_pitch, dtype=torch.bool)
    mask[choff:choff + cout] = False
    assert torch.all(out[..., mask.cuda()].float() == 7.0)


@pytest.mark.parametrize("cin,cout", [(64, 64), (96, 32)])
def test_conv_fp16_format(native_lib, cin, cout):
    """Same kernels with fp16 tensors/weights (HR tail and SRVGG layers); tolerance 2^-11 relative."""
    from framewright_b200.engine import debug_conv3x3

    g = torch.Generator().manual_seed(77 + cin)
    n, h, w = 1, 19, 140
    pitch = 64 if cout == 64 else 192
    x = (torch.randn(n, h, w, pitch, generator=g) * 0.5).to(torch.float16).cuda()
    weight = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    choff = 0 if cout == 64 else cin
    out = torch.zeros((n, h, w, pitch), dtype=torch.float16, device="cuda")
    debug_conv3x3(x, cin, weight, bias, out, out_choff=choff, slope=0.2)
    torch.cuda.synchronize()
    ref = _ref_conv(x, cin, weight, bias, slope=0.2)
    err = (out[..., choff:choff + cout].float() - ref).abs()
    assert torch.all(err <= ref.abs() * 2.0 ** -10 + 2e-4), float(err.max())


def test_conv_prelu(native_lib):
    from framewright_b200.engine import debug_conv3x3

    g = torch.Generator().manual_seed(5)
    n, h, w, cin, cout = 1, 24, 160, 64, 64
    x = (torch.randn(n, h, w, 64, generator=g) * 0.5).to(torch.bfloat16).cuda()
    weight = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    prelu = 0.25 + (torch.rand(64, generator=g) - 0.5) * 0.2
    out = torch.zeros((n, h, w, 64), dtype=torch.bfloat16, device="cuda")
    debug_conv3x3(x, cin, weight, bias, out, prelu=prelu)
    torch.cuda.synchronize()
    ref = _ref_conv(x, cin, weight, bias, prelu=prelu)
    err = (out.float() - ref).abs()
    assert torch.all(err <= ref.abs() * 2.0 ** -8 + 1e-3), float(err.max())
