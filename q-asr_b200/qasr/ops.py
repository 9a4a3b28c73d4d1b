"""Thin torch-tensor wrappers over the C ABI (one Python function per entry point of
include/qasr.h).  Tensors are plumbing only: raw device pointers, sizes and the current
CUDA stream are what crosses the boundary."""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def round_up(a, b):
    return (a + b - 1) // b * b


def alloc_codes(B, Cn, T, device, dtype=torch.int8):
    """Zero-initialised [B][C][ld] code tensor, ld = T rounded up to 16 (the pitch padding
    must stay zero: kernels read whole aligned words / TMA boxes)."""
    return torch.zeros((B, Cn, round_up(T, 16)), dtype=dtype, device=device)


def chan_table(s_out, bias, m, e, device):
    """Pack per-channel (s_out f32, bias i32, m i32, e i32) into the 16-byte qasr_chan records."""
    s_out = np.ascontiguousarray(np.asarray(s_out, np.float32).reshape(-1))
    n = s_out.size
    tab = np.zeros((n, 4), np.int32)
    tab[:, 0] = s_out.view(np.int32)
    tab[:, 1] = 0 if bias is None else np.asarray(bias, np.int64).astype(np.int32)
    tab[:, 2] = np.asarray(m, np.int64).astype(np.int32)
    tab[:, 3] = np.asarray(e, np.int64).astype(np.int32)
    return torch.from_numpy(tab).to(device)


def make_epilogue(mode, ch0, flags=0, ch1=None, prev=None, qmin=-128, qmax=127, m2=1 << 30, e2=30,
                  qmin2=0, qmax2=0, out_lens=None, f32_mult=None):
    ep = L.Epilogue()
    ep.mode, ep.flags = mode, flags
    ep.ch0 = ch0.data_ptr()
    ep.ch1 = ch1.data_ptr() if ch1 is not None else None
    ep.prev = prev.data_ptr() if prev is not None else None
    ep.prev_ld = prev.shape[-1] if prev is not None else 0
    ep.qmin, ep.qmax, ep.m2, ep.e2, ep.qmin2, ep.qmax2 = qmin, qmax, int(m2), int(e2), qmin2, qmax2
    ep.out_lens = out_lens.data_ptr() if out_lens is not None else None
    ep.f32_mult = f32_mult.data_ptr() if f32_mult is not None else None
    ep._keep = (ch0, ch1, prev, out_lens, f32_mult)
    return ep


def dwconv(x, T_in, w, out, T_out, k, stride, dilation, pad, ep):
    B, Cn, x_ld = x.shape
    L.check(L.lib().qasr_dwconv1d_i8(L.ptr(x), x_ld, L.ptr(w), L.ptr(out), out.shape[-1], B, Cn, T_in,
                                     T_out, k, stride, dilation, pad, C.byref(ep), L.stream_ptr()))
    return out


def pack_dw_weights(w_int8, stride, dilation, pad):
    """w_int8: [C][k] int8 cuda -> banded A-fragment table for qasr_dwconv1d_i8."""
    Cn, k = w_int8.shape
    nbytes = int(L.load_library().qasr_dw_packed_bytes(Cn, k, stride, dilation, pad))
    if nbytes <= 0:
        raise NotImplementedError("qasr: depthwise conv k=%d stride=%d dilation=%d pad=%d exceeds the supported "
                                  "tap window" % (k, stride, dilation, pad))
    wp = torch.empty(nbytes, dtype=torch.int8, device=w_int8.device)
    L.check(L.lib().qasr_pack_dw_weights(L.ptr(w_int8.contiguous()), L.ptr(wp), Cn, k, stride, dilation, pad,
                                         L.stream_ptr()))
    return wp


def packed_weight_bytes(c_out, c_in, k):
    return int(L.load_library().qasr_packed_weight_bytes(c_out, c_in, k))


def pack_weights(w_int8):
    """w_int8: [C_out][C_in][k] int8 cuda -> packed [k][C_out_pad][C_in_pad] int8."""
    c_out, c_in, k = w_int8.shape
    wp = torch.empty(packed_weight_bytes(c_out, c_in, k), dtype=torch.int8, device=w_int8.device)
    L.check(L.lib().qasr_pack_conv_weights(L.ptr(w_int8.contiguous()), L.ptr(wp), c_out, c_in, k,
                                           L.stream_ptr()))
    return wp


def gemm_operand(x, T_in, wp, c_in, k, stride=1, dilation=1, pad=0):
    op = L.GemmOperand()
    op.x, op.wp = x.data_ptr(), wp.data_ptr()
    op.x_ld, op.C_in, op.T_in, op.k = x.shape[-1], c_in, T_in, k
    op.stride, op.dilation, op.pad = stride, dilation, pad
    op._keep = (x, wp)
    return op


def conv_gemm(op0, op1, out, B, c_out, T_out, ep):
    L.check(L.lib().qasr_conv1d_gemm_i8(C.byref(op0), C.byref(op1) if op1 is not None else None,
                                        L.ptr(out), out.shape[-1], B, c_out, T_out, C.byref(ep),
                                        L.stream_ptr()))
    return out


def conv_ntc(op0, op1, out, B, c_out, T_out, ep):
    """Channels-last implicit GEMM (any k / stride / dilation); operands are [B][T][ld] code tensors."""
    L.check(L.lib().qasr_conv1d_ntc_i8(C.byref(op0), C.byref(op1) if op1 is not None else None,
                                       L.ptr(out), out.shape[-1], B, c_out, T_out, C.byref(ep),
                                       L.stream_ptr()))
    return out


def alloc_codes_ntc(B, Cn, T, device):
    """Zero-initialised channels-last code tensor [B][T][ld], ld = C rounded up to 16."""
    return torch.zeros((B, T, round_up(Cn, 16)), dtype=torch.int8, device=device)


def quantize_f32_ntc(x, scale, mode, bit=8, lens=None):
    """(B,C,T) float -> channels-last codes; mode 0 = first-layer quantiser, 1 = rne(x / s)."""
    B, Cn, T = x.shape
    x = x.contiguous()
    codes = alloc_codes_ntc(B, Cn, T, x.device)
    L.check(L.lib().qasr_quantize_f32_ntc(L.ptr(x), T, L.ptr(scale), mode, bit, L.ptr(codes), codes.shape[-1],
                                          B, Cn, T, L.ptr(lens), L.stream_ptr()))
    return codes


def quantize_f32(x, scale, bit, lens=None, want_codes=True, want_float=True):
    B, Cn, T = x.shape
    x = x.contiguous()
    codes = alloc_codes(B, Cn, T, x.device) if want_codes else None
    xq = torch.empty_like(x) if want_float else None
    L.check(L.lib().qasr_quantize_f32(L.ptr(x), T, L.ptr(scale), bit, L.ptr(codes),
                                      codes.shape[-1] if codes is not None else 0, L.ptr(xq), B, Cn, T,
                                      L.ptr(lens), L.stream_ptr()))
    return codes, xq


def requant_f32(x, pre_sf, z_sf, qmin, qmax, identity=None, id_sf=None, lens=None, want_codes=True,
                want_float=True):
    B, Cn, T = x.shape
    x = x.contiguous()
    if identity is not None:
        identity = identity.contiguous()
    codes = alloc_codes(B, Cn, T, x.device) if want_codes else None
    xq = torch.empty_like(x) if want_float else None
    pre_sf = pre_sf.reshape(-1).contiguous()
    if id_sf is not None:
        id_sf = id_sf.reshape(-1).contiguous()
    L.check(L.lib().qasr_requant_f32(L.ptr(x), L.ptr(pre_sf), pre_sf.numel(), L.ptr(identity), L.ptr(id_sf),
                                     id_sf.numel() if id_sf is not None else 0, L.ptr(z_sf), qmin, qmax,
                                     L.ptr(codes), codes.shape[-1] if codes is not None else 0, L.ptr(xq),
                                     T, B, Cn, T, L.ptr(lens), L.stream_ptr()))
    return codes, xq


def dyadic_from_scales(pre_sf, z_sf):
    pre_sf = pre_sf.reshape(-1).contiguous()
    m = torch.empty(pre_sf.numel(), dtype=torch.int32, device=pre_sf.device)
    e = torch.empty_like(m)
    L.check(L.lib().qasr_dyadic_from_scales(L.ptr(pre_sf), pre_sf.numel(), L.ptr(z_sf), L.ptr(m), L.ptr(e),
                                            L.stream_ptr()))
    return m, e


def codes_from_f32(x, scale):
    B, Cn, T = x.shape
    x = x.contiguous()
    codes = alloc_codes(B, Cn, T, x.device)
    L.check(L.lib().qasr_codes_from_f32(L.ptr(x), T, L.ptr(scale), L.ptr(codes), codes.shape[-1], B, Cn, T,
                                        L.stream_ptr()))
    return codes


def noise_codes(codes, scale, in_unsigned):
    """LUT-mapped codes of the reference's un-rounded x_int (quant_modules.py:301): eps(q) * 2^22 as int8, same layout."""
    codes = codes.contiguous()
    out = torch.empty_like(codes)
    L.check(L.lib().qasr_noise_codes_i8(L.ptr(codes), int(bool(in_unsigned)), L.ptr(scale), L.ptr(out), codes.numel(),
                                        L.stream_ptr()))
    return out


def combine_noise(acc, dacc, T, bias_f32, s_out):
    """f32(f64(acc) + bias + f64(dacc) * 2^-22) * s_out[c] (quant_modules.py:305-307) -> (B, C, T) float32."""
    B, Cn, ld = acc.shape
    out = torch.empty((B, Cn, T), dtype=torch.float32, device=acc.device)
    s_out = s_out.reshape(-1).float().contiguous()
    bias = None if bias_f32 is None else bias_f32.reshape(-1).float().contiguous()
    L.check(L.lib().qasr_combine_noise_f32(L.ptr(acc), L.ptr(dacc), ld, L.ptr(bias), L.ptr(s_out), L.ptr(out), T,
                                           B, Cn, T, L.stream_ptr()))
    return out


def requant_i8(x, T, in_signed, m, e, qmin, qmax, lens=None):
    B, Cn, ld = x.shape
    out = torch.zeros_like(x)
    L.check(L.lib().qasr_requant_i8(L.ptr(x), ld, int(in_signed), int(m), int(e), qmin, qmax, L.ptr(out), ld,
                                    B, Cn, T, L.ptr(lens), L.stream_ptr()))
    return out


_ws_cache = {}


def _workspace(device, nbytes):
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 8192), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def minmax_f32(x, identity=None, lens=None):
    """Returns a device tensor [min, max] of x (or identity + x)."""
    B, Cn, T = x.shape
    x = x.contiguous()
    if identity is not None:
        identity = identity.contiguous()
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    ws = _workspace(x.device, int(L.lib().qasr_minmax_workspace_bytes()))
    L.check(L.lib().qasr_minmax_f32(L.ptr(x), L.ptr(identity), T, B, Cn, T, L.ptr(lens), L.ptr(out),
                                    L.ptr(ws), L.stream_ptr()))
    return out


def quantile2_f32(x, q_lo, q_hi, identity=None, lens=None):
    B, Cn, T = x.shape
    x = x.contiguous()
    if identity is not None:
        identity = identity.contiguous()
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    ws = _workspace(x.device, int(L.lib().qasr_quantile_workspace_bytes()))
    L.check(L.lib().qasr_quantile2_f32(L.ptr(x), L.ptr(identity), T, B, Cn, T, L.ptr(lens),
                                       C.c_float(float(q_lo)), C.c_float(float(q_hi)), L.ptr(out), L.ptr(ws),
                                       L.stream_ptr()))
    return out


def ctc_argmax(logits, T):
    B, V, ld = logits.shape
    labels = torch.empty((B, T), dtype=torch.int32, device=logits.device)
    L.check(L.lib().qasr_ctc_argmax(L.ptr(logits), ld, B, V, T, L.ptr(labels), L.stream_ptr()))
    return labels


def log_softmax_tv(logits, T):
    B, V, ld = logits.shape
    out = torch.empty((B, T, V), dtype=torch.float32, device=logits.device)
    L.check(L.lib().qasr_log_softmax_tv(L.ptr(logits), ld, B, V, T, L.ptr(out), L.stream_ptr()))
    return out


def ctc_collapse(labels, blank_id, lens=None):
    """Greedy CTC collapse of (B, T) int32 labels -> (tokens (B, T) int32 padded with -1, counts (B,))."""
    labels = labels.to(torch.int32).contiguous()
    B, T = labels.shape
    tokens = torch.empty_like(labels)
    counts = torch.empty(B, dtype=torch.int32, device=labels.device)
    if lens is not None:
        lens = lens.to(torch.int32).contiguous()
    L.check(L.lib().qasr_ctc_collapse(L.ptr(labels), B, T, L.ptr(lens), blank_id, L.ptr(tokens), L.ptr(counts),
                                      L.stream_ptr()))
    return tokens, counts
