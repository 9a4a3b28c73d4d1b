"""EncDecCTCModel of the quantised path (reference nemo/collections/asr/models/ctc_models.py:
constructor :88-120 forces quant_mode 'symmetric', forward :384-406, set_quant_bit/mode :141-147)
without the PyTorch-Lightning / NeMo-core substrate: preprocessor (optional) -> encoder ->
decoder -> greedy argmax."""
import torch
import torch.nn as nn

from nemo.collections.asr.modules.conv_asr import ConvASRDecoder, ConvASREncoder
from qasr import topology as _topology

__all__ = ["EncDecCTCModel"]

EN_VOCABULARY = [" ", "a", "b", "c", "d", "e", "f", "g", "h", "i", "j", "k", "l", "m", "n", "o", "p", "q", "r",
                 "s", "t", "u", "v", "w", "x", "y", "z", "'"]


class EncDecCTCModel(nn.Module):
    def __init__(self, jasper, feat_in=64, vocabulary=None, preprocessor=None, quant_bit=8):
        super().__init__()
        vocabulary = EN_VOCABULARY if vocabulary is None else vocabulary
        self.preprocessor = preprocessor
        self.spec_augmentation = None
        # the reference overrides the configs to quant_mode 'symmetric' (ctc_models.py:103-107)
        self.encoder = ConvASREncoder(jasper=jasper, activation="relu", feat_in=feat_in, conv_mask=True,
                                      quant_mode="symmetric", quant_bit=quant_bit)
        self.decoder = ConvASRDecoder(feat_in=jasper[-1]["filters"], num_classes=len(vocabulary),
                                      vocabulary=vocabulary, quant_mode="symmetric", quant_bit=quant_bit)
        self._plan_cache = {}
        self.use_fused_plan = True

    @classmethod
    def from_topology(cls, name, **kw):
        return cls(_topology.TOPOLOGIES[name](), **kw)

    def set_quant_bit(self, quant_bit, mode="all"):
        self.encoder.set_quant_bit(quant_bit, mode)
        self.decoder.set_quant_bit(quant_bit, mode)
        self._plan_cache = {}

    def set_quant_mode(self, quant_mode):
        self.encoder.set_quant_mode(quant_mode)
        self.decoder.set_quant_mode(quant_mode)
        self._plan_cache = {}

    def _fusable(self, x):
        return (self.use_fused_plan and self.decoder.quant_mode == "symmetric" and self.encoder._fusable(x)
                and not (self.decoder.act.running_stat or self.decoder.act.dynamic or self.decoder.act.per_channel))

    def plan_for(self, shape, device):
        from qasr import plan as _plan
        return _plan.get_plan(self.encoder, self.decoder, shape, device, self._plan_cache)

    def forward(self, input_signal=None, input_signal_length=None, processed_signal=None,
                processed_signal_length=None):
        has_input = input_signal is not None and input_signal_length is not None
        has_processed = processed_signal is not None and processed_signal_length is not None
        if (has_input ^ has_processed) is False:
            raise ValueError(f"{self} Arguments ``input_signal`` and ``input_signal_length`` are mutually "
                             "exclusive with ``processed_signal`` and ``processed_signal_len`` arguments.")
        if not has_processed:
            if self.preprocessor is None:
                raise ValueError("raw audio given but the model has no preprocessor")
            processed_signal, processed_signal_length = self.preprocessor(
                input_signal=input_signal, length=input_signal_length)
        if self._fusable(processed_signal):
            p = self.plan_for(processed_signal.shape, processed_signal.device)
            p.run(processed_signal, processed_signal_length)
            return p.log_probs(), p.out_lens().to(torch.long), p.greedy().to(torch.long)
        encoded, encoded_len, encoded_sf = self.encoder(audio_signal=processed_signal,
                                                        length=processed_signal_length)
        log_probs = self.decoder(encoder_output=encoded, encoder_output_scaling_factor=encoded_sf)
        return log_probs, encoded_len, log_probs.argmax(dim=-1, keepdim=False)
