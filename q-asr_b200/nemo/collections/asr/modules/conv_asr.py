"""Drop-in ConvASREncoder / ConvASRDecoder of the quantised path (reference
nemo/collections/asr/modules/conv_asr.py:47-220, 223-319) as plain torch modules: the NeMo
NeuralModule / typecheck / omegaconf substrate is not part of the hot path.

Two execution levels share these classes:
  * module level (always available): every JasperBlock / QuantAct / QuantConv1d call is its own
    kernel sequence and hands float32 (code * scale) tensors on, exactly like the reference;
  * fused level: when every quant module is fixed, static (not dynamic) and on CUDA,
    `forward` runs the compiled int8-resident plan of qasr.plan (one CUDA graph) and only the
    final encoder tensor is materialised as float32.  Set `use_fused_plan = False` to force the
    module-level path.
"""
from typing import Optional

import torch
import torch.nn as nn

from nemo.collections.asr.parts.jasper import JasperBlock, MaskedConv1d, init_weights, jasper_activations
from nemo.quantization.utils.quant_modules import *  # noqa: F401,F403
from nemo.quantization.utils.quant_modules import QuantAct, QuantConv1d

__all__ = ["ConvASRDecoder", "ConvASREncoder"]


class ConvASREncoder(nn.Module):
    def __init__(self, jasper, activation: str, feat_in: int, normalization_mode: str = "batch",
                 residual_mode: str = "add", norm_groups: int = -1, conv_mask: bool = True,
                 frame_splicing: int = 1, init_mode: Optional[str] = "xavier_uniform",
                 quant_mode: Optional[str] = "none", quant_bit: Optional[int] = 8):
        super().__init__()
        jasper = [dict(c) for c in jasper]
        self._activation_name = activation
        self._check_activation(quant_mode)
        act = jasper_activations[activation]()
        feat_in = feat_in * frame_splicing
        self._feat_in = feat_in
        self.quant_mode = quant_mode
        self.convs_before_bn = []
        self.dense_residual = False
        self.use_fused_plan = True
        self._plan_cache = {}
        residual_panes = []
        encoder_layers = []
        for i, lcfg in enumerate(jasper):
            dense_res = []
            if lcfg.get("residual_dense", False):
                residual_panes.append(feat_in)
                dense_res = residual_panes
                self.dense_residual = True
            block = JasperBlock(
                feat_in, lcfg["filters"], repeat=lcfg["repeat"], kernel_size=lcfg["kernel"],
                stride=lcfg["stride"], dilation=lcfg["dilation"], dropout=lcfg.get("dropout", 0.0),
                residual=lcfg["residual"], groups=lcfg.get("groups", 1), separable=lcfg.get("separable", False),
                heads=lcfg.get("heads", -1), residual_mode=lcfg.get("residual_mode", residual_mode),
                normalization=normalization_mode, norm_groups=norm_groups, activation=act,
                residual_panes=dense_res, conv_mask=conv_mask, se=lcfg.get("se", False),
                se_reduction_ratio=lcfg.get("se_reduction_ratio", 8),
                se_context_window=lcfg.get("se_context_size", -1),
                se_interpolation_mode=lcfg.get("se_interpolation_mode", "nearest"),
                kernel_size_factor=lcfg.get("kernel_size_factor", 1.0),
                stride_last=lcfg.get("stride_last", False), quant_mode=quant_mode, quant_bit=quant_bit,
                layer_num=i)
            encoder_layers.append(block)
            self.convs_before_bn += block.convs_before_bn
            feat_in = lcfg["filters"]
        self._feat_out = feat_in
        self.encoder_layers = encoder_layers
        self.encoder = nn.Sequential(*encoder_layers)
        self.apply(lambda x: init_weights(x, mode=init_mode))

    # -- fused plan -----------------------------------------------------------------------
    def _fusable(self, audio_signal):
        if not (self.use_fused_plan and self.quant_mode == "symmetric" and audio_signal.is_cuda and not self.training):
            return False
        # every flag below is a stamped attribute (quant_modules._bump): the verdict holds until the state epoch moves
        from nemo.quantization.utils import quant_modules as _qm
        epoch = _qm._STATE_COUNTER[0]
        hit = getattr(self, "_fusable_verdict", None)
        if hit is not None and hit[0] == epoch:
            return hit[1]
        ok = True
        for m in self.modules():
            if type(m) is QuantAct and (m.running_stat or m.dynamic or m.per_channel):
                ok = False
                break
            if isinstance(m, QuantConv1d) and (m.quant_mode != "symmetric" or (m.bn is not None and not m.fix_bn)):
                ok = False
                break
        object.__setattr__(self, "_fusable_verdict", (epoch, ok))
        return ok

    def invalidate_plan(self):
        self._plan_cache = {}

    def forward(self, audio_signal, length=None, audio_signal_scaling_factor=None):
        if length is not None and audio_signal_scaling_factor is None and self._fusable(audio_signal):
            from qasr import plan as _plan
            try:
                p = _plan.get_plan(self, None, audio_signal.shape, audio_signal.device, self._plan_cache)
            except NotImplementedError:
                p = None   # a structure the plan compiler does not cover (remembered): module-level path below
            if p is not None:
                return p.run_encoder(audio_signal, length)
        s_input = [(audio_signal, audio_signal_scaling_factor)]
        for layer in self.encoder_layers:
            s_input, length = layer((s_input, length))
        out, out_sf = s_input[-1]
        if length is None:
            assert self.quant_mode != "symmetric"
            return out
        return out, length, out_sf

    def bn_folding(self):
        for l in self.encoder_layers:
            l.bn_folding()
        self.invalidate_plan()

    def set_quant_bit(self, quant_bit, mode="all"):
        assert mode in ["all", "weight", "act"]
        for l in self.encoder_layers:
            l.set_quant_bit(quant_bit, mode)
        self.invalidate_plan()

    def _check_activation(self, quant_mode):
        """Every depthwise / dense / residual conv after the first quantises its input with an UNSIGNED range
        (act_bit + 1, jasper.py:159-164), which is only right behind a non-negative activation: the reference
        asserts it on every call (`assert not (x.min() < -1e-5 and self.asymmetric)`, jasper.py:176, one host
        sync each).  Here the condition is checked once, where it is decided."""
        if quant_mode == "symmetric" and self._activation_name != "relu":
            raise NotImplementedError("quantised execution needs a non-negative block activation (relu): '%s' would "
                                      "feed negative values to the unsigned activation quantisers "
                                      "(reference assertion jasper.py:176)" % self._activation_name)

    def set_quant_mode(self, quant_mode):
        self._check_activation(quant_mode)
        self.quant_mode = quant_mode
        for l in self.encoder_layers:
            l.set_quant_mode(self.quant_mode)
        self.invalidate_plan()


class ConvASRDecoder(nn.Module):
    """QuantAct -> 1x1 QuantConv1d (with bias, no BN) -> log_softmax (reference :223-319)."""

    def __init__(self, feat_in, num_classes, init_mode="xavier_uniform", vocabulary=None, quant_mode="none",
                 quant_bit=8):
        super().__init__()
        self.quant_mode = quant_mode
        if vocabulary is not None and num_classes != len(vocabulary):
            raise ValueError("If vocabulary is specified, it's length should be equal to the num_classes. "
                             f"Instead got: num_classes={num_classes} and len(vocabulary)={len(vocabulary)}")
        self.__vocabulary = vocabulary
        self._feat_in = feat_in
        self._num_classes = num_classes + 1  # blank
        self.act = QuantAct(quant_bit, quant_mode=self.quant_mode, per_channel=False)
        conv = nn.Conv1d(self._feat_in, self._num_classes, kernel_size=1, bias=True)
        qconv = QuantConv1d(quant_bit, bias_bit=32, quant_mode=self.quant_mode, per_channel=True)
        qconv.set_param(conv)
        self.decoder_layers = nn.Sequential(qconv)
        self.apply(lambda x: init_weights(x, mode=init_mode))

    def forward(self, encoder_output, encoder_output_scaling_factor=None):
        output, output_sf = self.act(encoder_output, encoder_output_scaling_factor)
        for l in self.decoder_layers:
            if l.quant_mode == "none":
                output, _ = l(output, output_sf)
            else:
                output, _ = l(output, output_sf, input_unsigned=False)
        self.last_logits = output
        if output.is_cuda:
            from qasr import ops as _ops
            return _ops.log_softmax_tv(output.contiguous(), output.shape[-1])
        return torch.nn.functional.log_softmax(output.transpose(1, 2), dim=-1)

    def set_quant_bit(self, quant_bit, mode="all"):
        assert mode in ["all", "weight", "act"]
        if mode in ["all", "act"]:
            self.act.activation_bit = quant_bit
        if mode in ["all", "weight"]:
            for l in self.decoder_layers:
                l.weight_bit = quant_bit

    def set_quant_mode(self, quant_mode):
        self.quant_mode = quant_mode
        self.act.quant_mode = quant_mode
        for l in self.decoder_layers:
            l.quant_mode = quant_mode

    @property
    def vocabulary(self):
        return self.__vocabulary

    @property
    def num_classes_with_blank(self):
        return self._num_classes
