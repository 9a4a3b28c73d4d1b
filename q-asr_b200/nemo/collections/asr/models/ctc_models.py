"""EncDecCTCModel of the quantised path (reference nemo/collections/asr/models/ctc_models.py:
constructor :88-120 forces quant_mode 'symmetric', forward :384-406, set_quant_bit/mode :141-147)
without the PyTorch-Lightning / NeMo-core substrate: preprocessor (optional) -> encoder ->
decoder -> greedy argmax."""
import inspect
import os
import tarfile
import tempfile

import torch
import torch.nn as nn

from nemo.collections.asr.modules.conv_asr import ConvASRDecoder, ConvASREncoder
from qasr import topology as _topology

# names inside a .nemo archive (reference nemo/core/classes/modelPT.py:40-41)
_MODEL_CONFIG_YAML = "model_config.yaml"
_MODEL_WEIGHTS = "model_weights.ckpt"

__all__ = ["EncDecCTCModel"]

EN_VOCABULARY = [" ", "a", "b", "c", "d", "e", "f", "g", "h", "i", "j", "k", "l", "m", "n", "o", "p", "q", "r",
                 "s", "t", "u", "v", "w", "x", "y", "z", "'"]


class EncDecCTCModel(nn.Module):
    def __init__(self, jasper, feat_in=64, vocabulary=None, preprocessor=None, quant_bit=8):
        super().__init__()
        vocabulary = EN_VOCABULARY if vocabulary is None else list(vocabulary)
        self.preprocessor = preprocessor
        self._cfg = dict(jasper=[dict(b) for b in jasper], feat_in=int(feat_in), vocabulary=list(vocabulary),
                         preprocessor=getattr(preprocessor, "_cfg", None))
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

    # ---- .nemo checkpoints (reference nemo/core/classes/modelPT.py:254-296 save, :343-405 restore) -------------
    def to_config_dict(self):
        """The `model:` subtree the reference stores as model_config.yaml (examples/asr/conf/quartznet_15x5.yaml:3-240)."""
        c = self._cfg
        cfg = dict(
            sample_rate=16000, labels=list(c["vocabulary"]),
            encoder=dict(_target_="nemo.collections.asr.modules.ConvASREncoder", feat_in=c["feat_in"], activation="relu",
                         conv_mask=True, jasper=[dict(b) for b in c["jasper"]]),
            decoder=dict(_target_="nemo.collections.asr.modules.ConvASRDecoder", feat_in=c["jasper"][-1]["filters"],
                         num_classes=len(c["vocabulary"]), vocabulary=list(c["vocabulary"])))
        if c["preprocessor"] is not None:
            cfg["preprocessor"] = dict(_target_="nemo.collections.asr.modules.AudioToMelSpectrogramPreprocessor",
                                       **c["preprocessor"])
        # bn_folding() removes the BatchNorm entries from every block's module list (jasper.py:451-474), which renumbers
        # the state_dict keys behind them: a checkpoint of a folded model only loads into a folded module tree
        from nemo.quantization.utils.quant_modules import QuantConv1d
        if any(isinstance(m, QuantConv1d) and m.bn is not None for m in self.encoder.modules()):
            cfg["quantization"] = dict(bn_folded=True)
        return cfg

    @classmethod
    def from_config_dict(cls, config, **kw):
        """Builds the model from a reference-style config (the `model:` subtree, or a top-level file holding it).
        Encoder / decoder / preprocessor entries follow the reference's yaml; both the `_target_` and the older
        `cls` + `params` spellings are accepted.  Entries the quantised path has no use for (datasets, optimiser,
        spec_augment) are ignored, exactly like the reference's inference script never touches them."""
        cfg = config["model"] if "model" in config and "encoder" not in config else config

        def params(node):
            node = dict(node)
            if "params" in node and isinstance(node["params"], dict):
                node = dict(node["params"])
            return {k: v for k, v in node.items() if k not in ("_target_", "cls")}

        enc, dec = params(cfg["encoder"]), params(cfg["decoder"])
        for key in ("activation", "conv_mask"):
            enc.pop(key, None)   # fixed by the quantised path (relu, masked convs: ctc_models.py:103-107)
        vocabulary = dec.get("vocabulary") or cfg.get("labels")
        if vocabulary is None:
            raise ValueError("config has neither decoder.vocabulary nor labels")
        if "num_classes" in dec and dec["num_classes"] != len(vocabulary):
            raise ValueError("If vocabulary is specified, it's length should be equal to the num_classes. "
                             f"Instead got: num_classes={dec['num_classes']} and len(vocabulary)={len(vocabulary)}")
        preprocessor = None
        if cfg.get("preprocessor") is not None:
            from nemo.collections.asr.modules.audio_preprocessing import AudioToMelSpectrogramPreprocessor
            pp = params(cfg["preprocessor"])
            known = inspect.signature(AudioToMelSpectrogramPreprocessor.__init__).parameters
            pp = {k: v for k, v in pp.items() if k in known}
            preprocessor = AudioToMelSpectrogramPreprocessor(**pp)
            preprocessor._cfg = pp
        return cls([dict(b) for b in enc["jasper"]], feat_in=int(enc.get("feat_in", 64)), vocabulary=list(vocabulary),
                   preprocessor=preprocessor, **kw)

    def save_to(self, save_path, plan_shape=None):
        """Writes a .nemo archive: tar.gz of model_config.yaml + model_weights.ckpt (modelPT.py:254-296).
        plan_shape=(B, C, T): the compiled int8 plan of the calibrated, fixed model for that input shape goes into the
        same archive (plan_manifest.json + plan.npz: packed weights, integer biases, dyadic tables, epilogue choices per
        launch); `EncDecCTCModel.load_plan(path)` runs it without building the module tree.  restore_from ignores the
        two extra members, exactly as the reference ignores unknown files."""
        import io
        import yaml
        extra = []
        if plan_shape is not None:
            extra = self.plan_for(tuple(plan_shape), next(self.parameters()).device).members()
        with tempfile.TemporaryDirectory() as tmpdir:
            with open(os.path.join(tmpdir, _MODEL_CONFIG_YAML), "w") as f:
                yaml.safe_dump(self.to_config_dict(), f, sort_keys=False)
            torch.save(self.state_dict(), os.path.join(tmpdir, _MODEL_WEIGHTS))
            with tarfile.open(save_path, "w:gz") as tar:
                tar.add(os.path.join(tmpdir, _MODEL_CONFIG_YAML), arcname="./" + _MODEL_CONFIG_YAML)
                tar.add(os.path.join(tmpdir, _MODEL_WEIGHTS), arcname="./" + _MODEL_WEIGHTS)
                for name, data in extra:
                    info = tarfile.TarInfo("./" + name)
                    info.size = len(data)
                    tar.addfile(info, io.BytesIO(data))

    @staticmethod
    def load_plan(path, device=None):
        """The compiled plan stored in a .nemo archive (or a stand-alone plan archive): a qasr.plan.FusedPlan whose
        run(x, lens) / greedy() / log_probs() need no model.  Raises on a libqasr version mismatch."""
        from qasr import plan as _plan
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        return _plan.FusedPlan.load(path, device)

    @classmethod
    def restore_from(cls, restore_path, override_config_path=None, map_location=None, strict=False, **kw):
        """Restores weights and configuration from a .nemo archive (modelPT.py:343-405): model_config.yaml (or the
        override yaml, whose `model:` subtree is taken) builds the module tree, model_weights.ckpt is loaded with
        strict=False exactly like the reference does whatever `strict` says (:401), and the model moves to
        `map_location` (CUDA when available)."""
        import yaml
        if not os.path.exists(restore_path):
            raise FileNotFoundError(f"Can't find {restore_path}")
        if map_location is None:
            map_location = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        with tempfile.TemporaryDirectory() as tmpdir:
            with tarfile.open(restore_path, "r:*") as tar:
                members = [m for m in tar.getmembers()
                           if os.path.basename(m.name) in (_MODEL_CONFIG_YAML, _MODEL_WEIGHTS) and m.isfile()]
                for m in members:   # flat extraction of the two known files only (no path traversal)
                    m.name = os.path.basename(m.name)
                    tar.extract(m, tmpdir, filter="data")
            config_yaml = override_config_path or os.path.join(tmpdir, _MODEL_CONFIG_YAML)
            with open(config_yaml) as f:
                conf = yaml.safe_load(f)
            instance = cls.from_config_dict(conf, **kw)
            model_cfg = conf["model"] if "model" in conf and "encoder" not in conf else conf
            if (model_cfg.get("quantization") or {}).get("bn_folded"):
                instance.encoder.bn_folding()      # same module numbering as the checkpoint (see to_config_dict)
            state = torch.load(os.path.join(tmpdir, _MODEL_WEIGHTS), map_location="cpu")
            instance.load_state_dict(state, strict=False)
        return instance.to(map_location)

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
        """The compiled int8 plan for inputs of `shape` (frames rounded up to a bucket, qasr.plan.bucket_frames;
        at most qasr.plan.MAX_PLANS are kept, least recently used first out)."""
        from qasr import plan as _plan
        return _plan.get_plan(self.encoder, self.decoder, shape, device, self._plan_cache)

    # ---- dynamic quantisation (quant_modules.py:149-167) on the fused device-side plan (qasr.dynplan) ----------------
    def _dynamic_fusable(self, x):
        from nemo.quantization.utils.quant_modules import QuantAct, QuantConv1d
        if not (self.use_fused_plan and x.is_cuda and not self.training and self.decoder.quant_mode == "symmetric"
                and self.encoder.quant_mode == "symmetric"):
            return False
        # stamped attributes only (quant_modules._bump): the verdict holds until the state epoch moves
        from nemo.quantization.utils import quant_modules as _qm
        epoch = _qm._STATE_COUNTER[0]
        hit = getattr(self, "_dynamic_verdict", None)
        if hit is not None and hit[0] == epoch:
            return hit[1]
        ok = True
        for m in self.modules():
            if type(m) is QuantAct and (m.running_stat or not m.dynamic or m.percentile or m.per_channel):
                ok = False
                break
            if isinstance(m, QuantConv1d) and (m.quant_mode != "symmetric" or (m.bn is not None and not m.fix_bn)):
                ok = False
                break
        object.__setattr__(self, "_dynamic_verdict", (epoch, ok))
        return ok

    def _dynamic_plan_or_none(self, x):
        """The device-side dynamic plan for this exact input shape (a dynamic range depends on every frame of the batch,
        so there are no frame buckets), None when the model is not in the fixed dynamic state or its structure is not
        covered (module-level path then).  Under qasr.dist data parallelism the plan exchanges every range between the
        ranks through peer memory; all ranks must run the same batches in lockstep."""
        if not self._dynamic_fusable(x):
            return None
        from qasr import dist as _dist
        from qasr import dynplan as _dyn
        from qasr import plan as _plan
        dp_on = _dist.is_enabled()
        key = ("dynamic", tuple(x.shape), str(x.device), dp_on)
        sig = _plan._signature(self.encoder, self.decoder)
        hit = self._plan_cache.get(key)
        if hit is not None and hit[0] == sig:
            self._plan_cache[key] = self._plan_cache.pop(key)
            return None if isinstance(hit[1], _plan.PlanUnsupported) else hit[1]
        self._plan_cache.pop(key, None)
        try:
            dp = None
            if dp_on:
                dp = _dist.peer_slots(x.device)
            p = _dyn.DynamicPlan(self.encoder, self.decoder, x.shape, x.device, dp=dp)
        except NotImplementedError as exc:
            self._plan_cache[key] = (sig, _plan.PlanUnsupported(str(exc)))
            return None
        self._plan_cache[key] = (sig, p)
        while len(self._plan_cache) > _plan.MAX_PLANS:
            self._plan_cache.pop(next(iter(self._plan_cache)))
        return p

    def forward_dynamic(self, processed_signal, processed_signal_length):
        """Dynamic-mode forward on the fused plan: (float32 logits (B, V, T'), greedy labels (B, T') int64), or None
        when the plan does not apply (the caller runs the module-level path)."""
        p = self._dynamic_plan_or_none(processed_signal)
        if p is None:
            return None
        p.run(processed_signal, processed_signal_length)
        if p.overflowed():      # more near-tied channels than the refinement covers: practically never
            return None
        return p.logits[:, :, :p.cur_T_out], p.greedy().to(torch.long)

    def _plan_or_none(self, x):
        """The fused plan when the model is in the fixed static state AND its structure can be compiled; None sends
        the caller down the module-level path (configurations the plan compiler does not cover, e.g. strided
        residual branches: it remembers the refusal, so the symbolic walk is not retried per call)."""
        if not self._fusable(x):
            return None
        try:
            return self.plan_for(x.shape, x.device)
        except NotImplementedError:
            return None

    @torch.no_grad()
    def transcribe_stream(self, batches, device=None, collapse=False):
        """Serving loop: the evaluation loop of examples/asr/quantization/inference.py:127-149 (for batch in the
        loader: .cuda() -> forward -> greedy argmax -> .cpu()) with the copies taken off the critical path.
        `batches` yields (processed_signal (B, C, T) float32, length (B,) int) HOST tensors, pinned for asynchronous
        copies, or (input_signal (B, samples) float32 audio, length (B,) int samples) when the model has a
        preprocessor (BASELINE config 5's chain: the mel featurizer runs on the device in front of the encoder);
        greedy label tensors (B, T') int64 come back on the host, in order.  The host->device copy of batch
        i+1 runs on a copy stream into the other of two staging buffers while the kernels of batch i run, and the
        labels of batch i are read back asynchronously and handed out while batch i+1 is in flight.  Falls back to
        forward() per batch when the model is not in the fused static state.
        collapse=True additionally runs the greedy CTC collapse (reference metrics/wer.py:117-136: merge repeats, drop
        blanks, over the whole row like the reference) on the device (qasr_ctc_collapse) and yields
        (tokens (B, T') int64 padded with -1, counts (B,) int64) instead of the frame labels."""
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
                consumed[j] = None
                # the block may be one the compute stream's queued kernels still use (caching allocator):
                # order the first copy into it after everything already queued there
                copier.wait_stream(compute)
            with torch.cuda.stream(copier):
                if consumed[j] is not None:
                    copier.wait_event(consumed[j])       # the kernels that read this staging buffer are done with it
                stage[j][0].copy_(x, non_blocking=True)
                stage[j][1].copy_(n.to(torch.int32), non_blocking=True)
                copied[j] = torch.cuda.Event()
                copied[j].record(copier)

        def _result(host):
            host = host.to(torch.long)
            return (host[:, :-1], host[:, -1]) if collapse else host

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
            raw_audio = x.dim() == 2
            if raw_audio:
                if self.preprocessor is None:
                    raise ValueError("raw audio given but the model has no preprocessor")
                x, n = self.preprocessor(input_signal=x, length=n)
                consumed[j] = torch.cuda.Event()       # the featurizer has read the staging buffer
                consumed[j].record(compute)
                n = n.to(torch.int32)
            p = self._plan_or_none(x)
            if p is not None:
                p.run(x, n)                               # device copy into the plan's input + one graph replay
                if not raw_audio:
                    consumed[j] = torch.cuda.Event()
                    consumed[j].record(compute)
                labels = p.greedy()
            else:
                labels = self.forward(processed_signal=x, processed_signal_length=n)[2].to(torch.int32)
                if not raw_audio:
                    consumed[j] = torch.cuda.Event()
                    consumed[j].record(compute)
            if collapse:
                from qasr import ops as _ops
                tokens, counts = _ops.ctc_collapse(labels, self.decoder.num_classes_with_blank - 1)
                labels = torch.cat([tokens, counts.view(-1, 1)], dim=1)   # one read-back: [tokens | count]
            if host_out[j] is None or host_out[j].shape != labels.shape:
                host_out[j] = torch.empty(labels.shape, dtype=torch.int32).pin_memory()
            host_out[j].copy_(labels, non_blocking=True)   # in stream order: before the next replay overwrites it
            out_ready[j] = torch.cuda.Event()
            out_ready[j].record(compute)
            if pending is not None:
                out_ready[pending].synchronize()
                yield _result(host_out[pending])
            pending = j
            i += 1
        out_ready[pending].synchronize()
        yield _result(host_out[pending])

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
        p = self._plan_or_none(processed_signal)
        if p is not None:
            p.run(processed_signal, processed_signal_length)
            return p.log_probs(), p.out_lens().to(torch.long), p.greedy().to(torch.long)
        p = self._dynamic_plan_or_none(processed_signal)
        if p is not None:
            p.run(processed_signal, processed_signal_length)
            if not p.overflowed():
                self.decoder.last_logits = p.logits[:, :, :p.cur_T_out]
                self.decoder.act.act_scaling_factor = p.decoder_act_scale.clone()   # as QuantAct.forward leaves it (:177)
                self.decoder.decoder_layers[0]._prepare()                           # conv_scaling_factor / weight_integer
                return p.log_probs(), p.out_lens().to(torch.long), p.greedy().to(torch.long)
        encoded, encoded_len, encoded_sf = self.encoder(audio_signal=processed_signal,
                                                        length=processed_signal_length)
        log_probs = self.decoder(encoder_output=encoded, encoder_output_scaling_factor=encoded_sf)
        return log_probs, encoded_len, log_probs.argmax(dim=-1, keepdim=False)
