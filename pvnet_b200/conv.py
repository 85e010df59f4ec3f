"""Host helpers for the tensor-core convolution primitive (pvnet_conv2d_nhwc).

Weight packing (done once at load time, not on the hot path): PyTorch's
[Cout,Cin,kh,kw] conv weight with an eval-mode BatchNorm folded in
(w' = w * gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps)) becomes the
K-major [Cout][kh*kw][Cin] matrix the kernel's weight tensor map reads, rounded to TF32.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
MODE_AUTO, MODE_PER_TAP, MODE_COLUMN = 0, 1, 2


def set_mode(mode: int):
    """Test hook (pvnet_conv_set_mode): which kernel runs layers both kernels support."""
    _native.check(_native.lib().pvnet_conv_set_mode(int(mode)), "pvnet_conv_set_mode")


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """Round fp32 to the nearest TF32 value (10 explicit mantissa bits), kept in fp32."""
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF          # round half away from zero on the magnitude bits
    return i.view(torch.float32)


def fold_bn(weight, bn_weight=None, bn_bias=None, bn_mean=None, bn_var=None, eps=1e-5, conv_bias=None):
    w = weight.detach().to(torch.float64)
    cout = w.shape[0]
    b = torch.zeros(cout, dtype=torch.float64, device=w.device) if conv_bias is None else conv_bias.detach().double()
    if bn_weight is not None:
        scale = bn_weight.detach().double() / torch.sqrt(bn_var.detach().double() + eps)
        w = w * scale[:, None, None, None]
        b = (b - bn_mean.detach().double()) * scale + bn_bias.detach().double()
    return w.float(), b.float()


def pack_weight(w: torch.Tensor, cin_pad: int | None = None, cout_pad: int | None = None, tf32: bool = True):
    """[Cout,Cin,kh,kw] -> [Cout_pad][kh*kw][Cin_pad] contiguous (zero padded)."""
    cout, cin, kh, kw = w.shape
    cin_pad = cin if cin_pad is None else cin_pad
    cout_pad = cout if cout_pad is None else cout_pad
    p = torch.zeros(cout_pad, kh * kw, cin_pad, dtype=torch.float32, device=w.device)
    p[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    return round_tf32(p) if tf32 else p.contiguous()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def conv2d_nhwc(inp, in_co, cin, w_packed, bias, out, out_co, cout, ksize, stride=1, dilation=1, act=ACT_NONE,
                res=None, res_co=0, round_out=False):
    """inp [b,H,W,in_cs] / out [b,Ho,Wo,out_cs] / res [b,Ho,Wo,res_cs]: contiguous NHWC fp32
    CUDA tensors; the conv reads channels [in_co,in_co+cin) and writes [out_co,out_co+cout)."""
    b, H, W, in_cs = inp.shape
    with torch.cuda.device(inp.device):
        _native.check(_native.lib().pvnet_conv2d_nhwc(
            _p(inp), in_cs, in_co, cin, _p(w_packed), _p(bias), _p(res), 0 if res is None else res.shape[3], res_co,
            _p(out), out.shape[3], out_co, cout, b, H, W, ksize, stride, dilation, act, int(round_out),
            ctypes.c_void_p(torch.cuda.current_stream(inp.device).cuda_stream)), "pvnet_conv2d_nhwc")
    return out
