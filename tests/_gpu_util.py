"""Helpers for the -m gpu parity tests: build slot-layout tensors with plain torch indexing (independently of the
pack kernels) and call the C-ABI through mmlf_b200._lib."""
import ctypes as C

import numpy as np
import torch

from mmlf_b200 import _lib
from mmlf_b200._lib import ConvArgs, call
from oracle.net import bf16_round, fp16_round

BF16, FP16 = 0, 1
TDT = {BF16: torch.bfloat16, FP16: torch.float16}
ROUND = {BF16: bf16_round, FP16: fp16_round}
ULP = {BF16: 2.0 ** -7, FP16: 2.0 ** -10}

DEV = 'cuda'


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pad16(x):
    return (x + 15) // 16 * 16


def to_slots(x_nhwc, ld, full_grid, Hp, Wp, dt=BF16):
    """numpy (B, h, w, C) -> torch bf16 [B*Hp*Wp, ld] on the device.  full_grid: the tensor covers the whole
    (H+1)x(W+1) slot grid; otherwise it is H x W and lives at slots (y+1, x+1) with a zero halo."""
    B, h, w, Cc = x_nhwc.shape
    s = torch.zeros((B, Hp, Wp, ld), dtype=torch.float32)
    t = torch.from_numpy(np.ascontiguousarray(x_nhwc))
    if full_grid:
        s[:, :, :, :Cc] = t
    else:
        s[:, 1:, 1:, :Cc] = t
    return s.reshape(B * Hp * Wp, ld).to(TDT[dt]).to(DEV)


def from_slots(t, B, Hp, Wp, Cc, full_grid):
    a = t.float().cpu().numpy().reshape(B, Hp, Wp, -1)
    return a[..., :Cc] if full_grid else a[:, 1:, 1:, :Cc]


def pack_weight(w, spatial=0, dgrad=0, groups=1, group_real=None, group_pad=None, n_pad=None, cin_pad=None, dt=BF16):
    cout, cin = w.shape[:2]
    group_real = group_real or cin
    group_pad = group_pad or pad16(cin)
    if not dgrad:
        n_pad = n_pad or pad16(cout)
        cin_pad = cin_pad or (groups * group_pad if groups > 1 else pad16(cin))
    kc = (cin_pad + 63) // 64
    wd = torch.from_numpy(np.ascontiguousarray(w)).to(DEV)
    out = torch.empty((n_pad, 4 * kc * 64), dtype=TDT[dt], device=DEV)
    call('mmlf_pack_conv_weight', ptr(wd), cout, cin, spatial, dgrad, groups, group_real, group_pad, ptr(out), n_pad,
         cin_pad, dt, stream())
    return out


def pack_bits(mask, ld_bits=None):
    """numpy bool (n_slots, C) -> torch int32 [n_slots][ld_bits] words, bit (c & 31) of word c // 32."""
    n, Cc = mask.shape
    words = (Cc + 31) // 32
    ld_bits = ld_bits or words
    m = np.zeros((n, ld_bits * 32), np.uint8)
    m[:, :Cc] = mask
    packed = np.packbits(m.reshape(n, ld_bits, 32), axis=2, bitorder='little').view(np.uint32).reshape(n, ld_bits)
    return torch.from_numpy(packed.view(np.int32).copy()).to(DEV)


def unpack_bits(t, Cc):
    a = t.cpu().numpy().view(np.uint32)
    bits = np.unpackbits(a.view(np.uint8).reshape(a.shape[0], -1), axis=1, bitorder='little')
    return bits[:, :Cc].astype(bool)


def run_conv(x_slots, ld_in, cin_pad, wpack, n_pad, B, H, W, ctype, *, bias=None, scale=None, shift=None, relu=False,
             gate_bits=None, relu_bits=None, ld_bits=0, out_mode=0, n_real=0, simt=False, ld_out=None, ab=BF16,
             out_dt=BF16, out2_dt=None, col_sums=None):
    """Returns out, or (out, out2) when out2_dt is given."""
    n_slots = B * (H + 1) * (W + 1)
    ld_out = ld_out or n_pad
    out2 = None
    if out_mode == 0:
        out = torch.full((n_slots, ld_out), float('nan'), dtype=TDT[out_dt], device=DEV)
        if out2_dt is not None:
            out2 = torch.full((n_slots, ld_out), float('nan'), dtype=TDT[out2_dt], device=DEV)
    elif out_mode == 1:
        out = torch.full((n_slots, ld_out), float('nan'), dtype=torch.float32, device=DEV)
    else:
        Ho, Wo = (H, W) if ctype else (H + 1, W + 1)
        out = torch.full((B, n_real, Ho, Wo), float('nan'), dtype=torch.float32, device=DEV)
    a = ConvArgs()
    a.in_, a.ld_in, a.cin_pad = x_slots.data_ptr(), ld_in, cin_pad
    a.wpack, a.n_pad = wpack.data_ptr(), n_pad
    a.B, a.H, a.W, a.type = B, H, W, ctype
    a.bias = bias.data_ptr() if bias is not None else None
    a.scale = scale.data_ptr() if scale is not None else None
    a.shift = shift.data_ptr() if shift is not None else None
    a.relu = int(relu)
    a.gate_bits = gate_bits.data_ptr() if gate_bits is not None else None
    a.relu_bits = relu_bits.data_ptr() if relu_bits is not None else None
    a.ld_bits = ld_bits
    a.out, a.ld_out, a.out_mode, a.n_real = out.data_ptr(), ld_out, out_mode, n_real
    a.out2 = out2.data_ptr() if out2 is not None else None
    a.ld_out2 = ld_out
    a.col_sums = col_sums.data_ptr() if col_sums is not None else None
    a.ab_dtype, a.out_dtype, a.out2_dtype = ab, out_dt, (out2_dt or 0)
    call('mmlf_conv2x2_simt' if simt else 'mmlf_conv2x2', C.byref(a), stream())
    torch.cuda.synchronize()
    return (out, out2) if out2 is not None else out


def dev_f32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)


def assert_close_bf16(got, want, what, ulps=1.0, atol=1e-6, dt=BF16):
    """|got - want| within `ulps` units-in-the-last-place (of the 16-bit storage format) of `want`."""
    want = np.asarray(want, np.float32)
    got = np.asarray(got, np.float32)
    tol = ulps * np.abs(want) * ULP[dt] + atol
    bad = np.abs(got - want) > tol
    assert not bad.any(), f'{what}: {bad.sum()} / {bad.size} outside {ulps} bf16 ulp; worst |d|={np.abs(got - want).max():.4g}'
