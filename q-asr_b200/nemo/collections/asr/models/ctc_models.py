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

    @torch.no_grad()
    def transcribe_stream(self, batches, device=None):
        """Serving loop: the evaluation loop of examples/asr/quantization/inference.py:127-149 (for batch in the
        loader: .cuda() -> forward -> greedy argmax -> .cpu()) with the copies taken off the critical path.
        `batches` yields (processed_signal (B, C, T) float32, length (B,) int) HOST tensors, pinned for asynchronous
        copies; greedy label tensors (B, T') int64 come back on the host, in order.  The host->device copy of batch
        i+1 runs on a copy stream into the other of two staging buffers while the kernels of batch i run, and the
        labels of batch i are read back asynchronously and handed out while batch i+1 is in flight.  Falls back to
        forward() per batch when the model is not in the fused static state."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        compute = torch.cuda.current_stream(device)
        copier = torch.cuda.Stream(device)
        stage, copied, consumed = [None, None], [None, None], [None, None]
        host_out, out_ready = [None, None], [None, None]
        it = iter(batches)

        def prefetch(j, batch):
            x, n = batch
            if stage[j] is None or stage[j][0].shape != x.shape:
                stage[j] = (torch.empty(x.shape, dtype=torch.float32, device=device),
                            torch.empty(n.shape, dtype=torch.int32, device=device))
            with torch.cuda.stream(copier):
                if consumed[j] is not None:
                    copier.wait_event(consumed[j])       # the kernels that read this staging buffer are done with it
                stage[j][0].copy_(x, non_blocking=True)
                stage[j][1].copy_(n.to(torch.int32), non_blocking=True)
                copied[j] = torch.cuda.Event()
                copied[j].record(copier)

        nxt = next(it, None)
        if nxt is None:
            return
        prefetch(0, nxt)
        i, pending = 0, None
        while nxt is not None:
            j = i & 1
            cur, nxt = nxt, next(it, None)
            if nxt is not None:
                prefetch(j ^ 1, nxt)
            compute.wait_event(copied[j])
            x, n = stage[j]
            if self._fusable(x):
                p = self.plan_for(x.shape, device)
                p.run(x, n)                               # device copy into the plan's input + one graph replay
                consumed[j] = torch.cuda.Event()
                consumed[j].record(compute)
                labels = p.greedy()
            else:
                labels = self.forward(processed_signal=x, processed_signal_length=n)[2].to(torch.int32)
                consumed[j] = torch.cuda.Event()
                consumed[j].record(compute)
            if host_out[j] is None or host_out[j].shape != labels.shape:
                host_out[j] = torch.empty(labels.shape, dtype=torch.int32).pin_memory()
            host_out[j].copy_(labels, non_blocking=True)   # in stream order: before the next replay overwrites it
            out_ready[j] = torch.cuda.Event()
            out_ready[j].record(compute)
            if pending is not None:
                out_ready[pending].synchronize()
                yield host_out[pending].to(torch.long)
            pending = j
            i += 1
        out_ready[pending].synchronize()
        yield host_out[pending].to(torch.long)

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
