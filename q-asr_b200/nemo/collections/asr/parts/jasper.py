"""Drop-in building blocks of the quantised Jasper / QuartzNet encoder: MaskedConv1d and
JasperBlock with the reference's constructor arguments, attribute names (mconv, res, res_act,
mout, convs_before_bn, dense_residual) and methods (bn_folding, set_quant_bit, set_quant_mode)
-- reference nemo/collections/asr/parts/jasper.py:116-212, 293-692 -- so that state_dicts,
`quantize_model` traversal and the `(tensor, scale)` threading are interchangeable.

Scope: what Q-ASR's quantised path exercises (batch norm, separable or dense convs, additive
residuals incl. Jasper's dense residual).  SqueezeExcite / GroupShuffle / group-instance-layer
norm are FP32-only in the reference (jasper.py:399 asserts quant_mode == 'none') and are not
part of this path; asking for them raises NotImplementedError.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from nemo.quantization.utils.quant_modules import *  # noqa: F401,F403
from nemo.quantization.utils.quant_modules import QuantAct, QuantConv1d


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


jasper_activations = {"hardtanh": nn.Hardtanh, "relu": nn.ReLU, "selu": nn.SELU, "swish": Swish}

_INITS = {
    "xavier_uniform": lambda w: nn.init.xavier_uniform_(w, gain=1.0),
    "xavier_normal": lambda w: nn.init.xavier_normal_(w, gain=1.0),
    "kaiming_uniform": lambda w: nn.init.kaiming_uniform_(w, nonlinearity="relu"),
    "kaiming_normal": lambda w: nn.init.kaiming_normal_(w, nonlinearity="relu"),
}


def init_weights(m, mode: Optional[str] = "xavier_uniform"):
    """Reference jasper.py:28-50.  As there, QuantConv1d.weight is a clone taken at set_param
    time and is NOT touched (it keeps PyTorch's default init); only plain Conv1d / Linear /
    BatchNorm1d modules are (re)initialised."""
    if isinstance(m, MaskedConv1d):
        init_weights(m.conv, mode)
    if isinstance(m, (nn.Conv1d, nn.Linear)):
        if mode is not None:
            if mode not in _INITS:
                raise ValueError("Unknown Initialization mode: {0}".format(mode))
            _INITS[mode](m.weight)
    elif isinstance(m, nn.BatchNorm1d):
        if m.track_running_stats:
            m.running_mean.zero_()
            m.running_var.fill_(1)
            m.num_batches_tracked.zero_()
        if m.affine:
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


def compute_new_kernel_size(kernel_size, kernel_width):
    new_kernel_size = max(int(kernel_size * kernel_width), 1)
    return new_kernel_size + 1 if new_kernel_size % 2 == 0 else new_kernel_size


def get_same_padding(kernel_size, stride, dilation):
    if stride > 1 and dilation > 1:
        raise ValueError("Only stride OR dilation may be greater than 1")
    if dilation > 1:
        return (dilation * kernel_size) // 2 - 1
    return kernel_size // 2


class MaskedConv1d(nn.Module):
    """length mask -> QuantAct -> QuantConv1d (reference jasper.py:116-212).  The activation
    quantiser of an `asymmetric` (post-ReLU) input gets one extra bit, i.e. an unsigned range."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 heads=-1, bias=False, use_mask=True, quant_mode="none", quant_bit=8, asymmetric=False):
        super().__init__()
        self.quant_mode = quant_mode
        self.asymmetric = asymmetric
        if not (heads == -1 or groups == in_channels):
            raise ValueError("Only use heads for depthwise convolutions")
        if heads != -1:
            raise NotImplementedError("heads != -1 is outside the quantised path")
        self.real_out_channels = out_channels
        conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias)
        act_quant_bit = quant_bit + 1 if asymmetric else quant_bit
        self.act = QuantAct(act_quant_bit, quant_mode=quant_mode, per_channel=False)
        self.conv = QuantConv1d(quant_bit, bias_bit=32, quant_mode=quant_mode, per_channel=True)
        self.conv.set_param(conv)
        self.use_mask = use_mask
        self.heads = heads

    def get_seq_len(self, lens):
        c = self.conv
        return (lens + 2 * c.padding[0] - c.dilation[0] * (c.kernel_size[0] - 1) - 1) // c.stride[0] + 1

    def forward(self, x, lens, scaling_factor=None):
        # The reference's `assert not (float(x.min()) < -1e-5 and self.asymmetric)` (jasper.py:176)
        # costs a host sync per call; the invariant (asymmetric inputs are post-ReLU) holds by
        # construction of the block and is not re-checked on the hot path.
        if self.conv.quant_mode == "symmetric" and self.act.activation_bit > (9 if self.asymmetric else 8):
            # codes travel as 8-bit tensors: signed inputs fit 8 bits, the unsigned (act_bit + 1) inputs 9
            raise NotImplementedError("act_bit %d: the integer kernels carry activation codes in 8 bits (act_bit <= 8)"
                                      % (self.act.activation_bit - (1 if self.asymmetric else 0)))
        fused_mask = False
        if self.use_mask:
            lens = lens.to(dtype=torch.long)
            # The reference zeroes x[b, :, t >= lens[b]] with masked_fill (jasper.py:177-181) and hands the copy to
            # QuantAct.  On the quantised CUDA path the range and requantisation kernels take the lengths instead
            # (masked positions count as 0 there: same statistics, same outputs, one pass over x less).
            fused_mask = (x.is_cuda and x.dim() == 3 and self.act.quant_mode == "symmetric" and not self.act.per_channel)
            if fused_mask:
                self.act._mask_lens = lens.to(device=x.device, dtype=torch.int32).contiguous()
            else:
                mask = torch.arange(x.size(2), device=x.device).unsqueeze(0) >= lens.to(x.device).unsqueeze(1)
                x = x.masked_fill(mask.unsqueeze(1), 0)
            lens = self.get_seq_len(lens)
        try:
            x, x_scaling_factor = self.act(x, scaling_factor)
        finally:
            if fused_mask:
                self.act._mask_lens = None
        if self.conv.quant_mode == "none":
            out, out_scaling_factor = self.conv(x, x_scaling_factor)
        else:
            out, out_scaling_factor = self.conv(x, x_scaling_factor, input_unsigned=self.asymmetric)
        return out, lens, out_scaling_factor

    def bn_folding(self, bn):
        self.conv.bn_folding(bn)

    def set_quant_bit(self, quant_bit, mode="all"):
        if mode in ("all", "act"):
            self.act.activation_bit = quant_bit + 1 if self.asymmetric else quant_bit
        if mode in ("all", "weight"):
            self.conv.weight_bit = quant_bit

    def set_quant_mode(self, quant_mode):
        self.quant_mode = quant_mode
        self.conv.quant_mode = quant_mode
        self.act.quant_mode = quant_mode


class JasperBlock(nn.Module):
    """repeat x (conv [+ pointwise conv] + BN + activation + dropout) with an optional additive
    (dense) residual through 1x1 conv + BN branches and `res_act` (reference jasper.py:293-692)."""

    def __init__(self, inplanes, planes, repeat=3, kernel_size=11, kernel_size_factor=1, stride=1, dilation=1,
                 padding="same", dropout=0.2, activation=None, residual=True, groups=1, separable=False,
                 heads=-1, normalization="batch", norm_groups=1, residual_mode="add", residual_panes=[],
                 conv_mask=False, se=False, se_reduction_ratio=16, se_context_window=None,
                 se_interpolation_mode="nearest", stride_last=False, quant_mode="none", quant_bit=8,
                 layer_num=-1):
        super().__init__()
        if padding != "same":
            raise ValueError("currently only 'same' padding is supported")
        if se or groups != 1 or heads != -1 or normalization != "batch":
            raise NotImplementedError("SqueezeExcite / grouped convs / non-batch norms are FP32-only in the "
                                      "reference and outside the quantised path")
        if not conv_mask and quant_mode != "none":
            raise AssertionError("Quantization mode only supports convolution with mask currently.")
        kernel_size_factor = float(kernel_size_factor)
        if type(kernel_size) in (list, tuple):
            kernel_size = [compute_new_kernel_size(k, kernel_size_factor) for k in kernel_size]
        else:
            kernel_size = [compute_new_kernel_size(kernel_size, kernel_size_factor)]
        stride = list(stride) if type(stride) in (list, tuple) else [stride]
        dilation = list(dilation) if type(dilation) in (list, tuple) else [dilation]
        padding_val = get_same_padding(kernel_size[0], stride[0], dilation[0])
        self.conv_mask = conv_mask
        self.separable = separable
        self.residual_mode = residual_mode
        self.se = se
        self.quant_mode = quant_mode
        self.layer_num = layer_num
        self.convs_before_bn = []

        def conv_bn(cin, cout, k, s, d, p, sep, first):
            return self._get_conv_bn_layer(cin, cout, kernel_size=k, stride=s, dilation=d, padding=p,
                                           separable=sep, quant_mode=quant_mode, quant_bit=quant_bit,
                                           is_first_layer=first)

        layers = nn.ModuleList()
        cin = inplanes
        for i in range(repeat - 1):
            s = [1] if stride_last else stride
            layers.extend(conv_bn(cin, planes, kernel_size, s, dilation, padding_val, separable,
                                  layer_num == 0 and i == 0))
            layers.extend(self._get_act_dropout_layer(drop_prob=dropout, activation=activation))
            cin = planes
        layers.extend(conv_bn(cin, planes, kernel_size, stride, dilation, padding_val, separable,
                              layer_num == 0 and repeat == 1))
        self.mconv = layers

        res_panes = residual_panes.copy()
        self.dense_residual = residual
        if residual:
            res_list = nn.ModuleList()
            res_stride = stride if residual_mode == "stride_add" else [1]
            if len(residual_panes) == 0:
                res_panes = [inplanes]
                self.dense_residual = False
            for ip in res_panes:
                res_list.append(nn.ModuleList(conv_bn(ip, planes, [1], res_stride, [1], 0, False, layer_num == 0)))
            self.res = res_list
        else:
            self.res = None
        self.res_act = QuantAct(quant_bit, quant_mode=quant_mode, per_channel=False)
        self.mout = nn.Sequential(*self._get_act_dropout_layer(drop_prob=dropout, activation=activation))

    # -- construction helpers -------------------------------------------------------------
    def _get_conv(self, in_channels, out_channels, kernel_size=11, stride=1, dilation=1, padding=0, bias=False,
                  groups=1, heads=-1, separable=False, quant_mode="none", quant_bit=8, asymmetric=False):
        if not self.conv_mask:
            assert quant_mode == "none", "Quantization mode only supports convolution with mask currently."
            return nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, dilation=dilation,
                             padding=padding, bias=bias, groups=groups)
        one = lambda v: v[0] if type(v) in (list, tuple) else v  # noqa: E731
        return MaskedConv1d(in_channels, out_channels, one(kernel_size), stride=one(stride),
                            dilation=one(dilation), padding=padding, bias=bias, groups=groups, heads=heads,
                            use_mask=True, quant_mode=quant_mode, quant_bit=quant_bit, asymmetric=asymmetric)

    def _get_conv_bn_layer(self, in_channels, out_channels, kernel_size=11, stride=1, dilation=1, padding=0,
                           bias=False, groups=1, heads=-1, separable=False, normalization="batch",
                           norm_groups=1, quant_mode="none", quant_bit=8, is_first_layer=False):
        if separable:
            layers = [
                self._get_conv(in_channels, in_channels, kernel_size, stride=stride, dilation=dilation,
                               padding=padding, bias=bias, groups=in_channels, quant_mode=quant_mode,
                               quant_bit=quant_bit, asymmetric=not is_first_layer),
                self._get_conv(in_channels, out_channels, kernel_size=1, stride=1, dilation=1, padding=0,
                               bias=bias, groups=groups, quant_mode=quant_mode, quant_bit=quant_bit,
                               asymmetric=False),
            ]
        else:
            layers = [self._get_conv(in_channels, out_channels, kernel_size, stride=stride, dilation=dilation,
                                     padding=padding, bias=bias, groups=groups, quant_mode=quant_mode,
                                     quant_bit=quant_bit, asymmetric=not is_first_layer)]
        layers.append(nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.1))
        self.convs_before_bn.append((layers[-2], layers[-1]))
        return layers

    def _get_act_dropout_layer(self, drop_prob=0.2, activation=None):
        if activation is None:
            activation = nn.Hardtanh(min_val=0.0, max_val=20.0)
        return [activation, nn.Dropout(p=drop_prob)]

    # -- mode switches --------------------------------------------------------------------
    def _masked_convs(self):
        for l in self.mconv:
            if isinstance(l, MaskedConv1d):
                yield l
        if self.res is not None:
            for branch in self.res:
                for l in branch:
                    if isinstance(l, MaskedConv1d):
                        yield l

    def bn_folding(self):
        """Attach every BatchNorm1d to the MaskedConv1d in front of it and drop it from the layer
        lists (reference jasper.py:451-474)."""
        def fold(seq):
            out = nn.ModuleList()
            for l in seq:
                if isinstance(l, nn.BatchNorm1d):
                    assert isinstance(out[-1], MaskedConv1d)
                    out[-1].bn_folding(l)
                else:
                    out.append(l)
            return out
        if self.mconv is not None:
            self.mconv = fold(self.mconv)
        if self.res is not None:
            self.res = nn.ModuleList(fold(branch) for branch in self.res)

    def set_quant_bit(self, quant_bit, mode="all"):
        for l in self._masked_convs():
            l.set_quant_bit(quant_bit, mode)
        self.res_act.activation_bit = quant_bit  # as the reference: regardless of `mode` (jasper.py:486)

    def set_quant_mode(self, quant_mode):
        self.quant_mode = quant_mode
        for l in self._masked_convs():
            l.set_quant_mode(quant_mode)
        self.res_act.quant_mode = quant_mode

    # -- forward --------------------------------------------------------------------------
    def forward(self, input_: Tuple[List[Tuple[Tensor, Optional[Tensor]]], Optional[Tensor]]):
        xs, lens_orig = input_ if len(input_) == 2 else (input_[0], None)
        out, out_sf = xs[-1]
        lens = lens_orig
        for l in self.mconv:
            if isinstance(l, MaskedConv1d):
                out, lens, out_sf = l(out, lens, out_sf)
            else:
                out = l(out)
        if self.res is not None:
            if self.quant_mode == "symmetric":
                assert self.residual_mode in ("add", "stride_add")
            for i, branch in enumerate(self.res):
                res_out, res_sf = xs[i]
                for l in branch:
                    if isinstance(l, MaskedConv1d):
                        res_out, _, res_sf = l(res_out, lens_orig, res_sf)
                    else:
                        res_out = l(res_out)
                if self.residual_mode in ("add", "stride_add"):
                    out, out_sf = self.res_act(out, out_sf, res_out, res_sf)
                else:
                    out = torch.max(out, res_out)
        out = self.mout(out)
        if self.res is not None and self.dense_residual:
            return xs + [(out, out_sf)], lens
        return [(out, out_sf)], lens
