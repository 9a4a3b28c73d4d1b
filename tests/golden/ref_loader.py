"""Load the UNMODIFIED reference quant path from /root/reference on CPU.

Only the golden-vector generator (tests/golden/gen_golden.py) and ad-hoc
oracle-pinning runs in the build container use this file.  Nothing in the
`-m gpu` tests, smoke() or bench.py imports it: /root/reference does not exist
on the GPU box.

Recipe:
  * NEMO_PACKAGE_BUILDING=1 skips nemo/__init__.py's heavy imports,
  * torch.Tensor.cuda is patched to identity (the reference hard-codes .cuda()
    at quant_utils.py:72,146-147 and quant_modules.py:122-123,161,165),
  * nemo.collections.asr.parts.{activations,jasper} are loaded by file path with
    stub parent packages, because nemo/collections/asr/__init__.py pulls
    pytorch_lightning.
The encoder / decoder classes (conv_asr.py) need omegaconf + nemo.core, so the
~40 lines of ConvASREncoder.__init__/forward (conv_asr.py:111-206) and
ConvASRDecoder (conv_asr.py:247-275) are restated here on top of the
reference's own JasperBlock / QuantAct / QuantConv1d classes.
"""
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("QASR_REFERENCE_ROOT", "/root/reference")


def load_reference():
    """Returns a namespace with the reference's quant_utils, quant_modules,
    quantize_model and jasper modules."""
    if "_qasr_ref_ns" in sys.modules:
        return sys.modules["_qasr_ref_ns"]
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present: %s" % REF_ROOT)
    os.environ["NEMO_PACKAGE_BUILDING"] = "1"
    for k in list(sys.modules):
        if k == "nemo" or k.startswith("nemo."):
            raise RuntimeError("a different `nemo` is already imported (%s); run the "
                               "golden generator in its own process" % k)
    sys.path.insert(0, REF_ROOT)
    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU shim
    import nemo.quantization.utils.quant_utils as qu
    import nemo.quantization.utils.quant_modules as qmod
    import nemo.quantization.utils.quantize_model as qm

    for name in ("nemo.collections", "nemo.collections.asr", "nemo.collections.asr.parts"):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = []
            sys.modules[name] = pkg

    def _load(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    _load("nemo.collections.asr.parts.activations", "nemo/collections/asr/parts/activations.py")
    jasper = _load("nemo.collections.asr.parts.jasper", "nemo/collections/asr/parts/jasper.py")

    ns = types.ModuleType("_qasr_ref_ns")
    ns.quant_utils, ns.quant_modules, ns.quantize_model, ns.jasper = qu, qmod, qm, jasper
    sys.modules["_qasr_ref_ns"] = ns
    return ns


class RefEncoder(torch.nn.Module):
    """conv_asr.py:111-206 restated over the reference's JasperBlock."""

    def __init__(self, ref, blocks, feat_in=64, quant_mode="symmetric", quant_bit=8):
        super().__init__()
        J = ref.jasper
        activation = J.jasper_activations["relu"]()
        residual_panes = []
        layers = []
        for i, lcfg in enumerate(blocks):
            dense_res = []
            if lcfg.get("residual_dense", False):
                residual_panes.append(feat_in)
                dense_res = residual_panes
            layers.append(J.JasperBlock(
                feat_in, lcfg["filters"], repeat=lcfg["repeat"], kernel_size=lcfg["kernel"],
                stride=lcfg["stride"], dilation=lcfg["dilation"], dropout=lcfg.get("dropout", 0.0),
                residual=lcfg["residual"], groups=1, separable=lcfg.get("separable", False),
                heads=-1, residual_mode="add", normalization="batch", norm_groups=-1,
                activation=activation, residual_panes=dense_res, conv_mask=True, se=False,
                quant_mode=quant_mode, quant_bit=quant_bit, layer_num=i))
            feat_in = lcfg["filters"]
        self.encoder_layers = layers
        self.encoder = torch.nn.Sequential(*layers)
        self.apply(lambda x: J.init_weights(x, mode="xavier_uniform"))

    def forward(self, audio_signal, length):
        s_input = [(audio_signal, None)]
        for layer in self.encoder_layers:
            s_input, length = layer((s_input, length))
        return s_input[-1][0], length, s_input[-1][1]

    def bn_folding(self):
        for l in self.encoder_layers:
            l.bn_folding()

    def set_quant_bit(self, bit, mode="all"):
        for l in self.encoder_layers:
            l.set_quant_bit(bit, mode)

    def set_quant_mode(self, mode):
        for l in self.encoder_layers:
            l.set_quant_mode(mode)


class RefDecoder(torch.nn.Module):
    """conv_asr.py:247-275 restated over the reference's QuantAct / QuantConv1d."""

    def __init__(self, ref, feat_in, num_classes=29, quant_mode="symmetric", quant_bit=8):
        super().__init__()
        Q = ref.quant_modules
        self.act = Q.QuantAct(quant_bit, quant_mode=quant_mode, per_channel=False)
        conv = torch.nn.Conv1d(feat_in, num_classes, kernel_size=1, bias=True)
        qconv = Q.QuantConv1d(quant_bit, bias_bit=32, quant_mode=quant_mode, per_channel=True)
        qconv.set_param(conv)
        self.decoder_layers = torch.nn.Sequential(qconv)
        self.apply(lambda x: ref.jasper.init_weights(x, mode="xavier_uniform"))

    def forward(self, encoder_output, encoder_output_scaling_factor=None):
        output, sf = self.act(encoder_output, encoder_output_scaling_factor)
        for l in self.decoder_layers:
            output, out_sf = l(output, sf)
        self.last_logits, self.last_logit_sf = output, out_sf
        return torch.nn.functional.log_softmax(output.transpose(1, 2), dim=-1)

    def set_quant_bit(self, bit, mode="all"):
        if mode in ("all", "act"):
            self.act.activation_bit = bit
        if mode in ("all", "weight"):
            for l in self.decoder_layers:
                l.weight_bit = bit


class RefModel(torch.nn.Module):
    def __init__(self, ref, blocks, quant_bit=8):
        super().__init__()
        self.encoder = RefEncoder(ref, blocks, quant_bit=quant_bit)
        self.decoder = RefDecoder(ref, blocks[-1]["filters"], quant_bit=quant_bit)

    def set_quant_bit(self, bit, mode="all"):
        self.encoder.set_quant_bit(bit, mode)
        self.decoder.set_quant_bit(bit, mode)
