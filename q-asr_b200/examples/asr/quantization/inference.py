"""Entry point of the quantised inference path, with the reference's flags
(reference examples/asr/quantization/inference.py:46-159): load model -> set_quant_bit ->
set_percentile -> bn_folding -> calibrate on synthetic batches -> evaluate (-> set_dynamic) -> test loop
with greedy CTC decoding and WER.

Differences forced by the environment (no network, no NeMo core): `--asr_model` is a local `.nemo` archive
(EncDecCTCModel.restore_from, like the reference's :72-74) or, instead of an NGC name, a topology name
(`quartznet15x5`, `jasper10x5dr`, ...) optionally followed by `:checkpoint.pt` (a torch state_dict with the
reference's key names, loaded with strict=False exactly like ModelPT.restore_from does), and `--dataset` is
either `synthetic:<batches>:<frames>` or a torch file holding a list of
(features (B,64,T) float32, lengths (B,) int64, transcripts list[str]) batches.
"""
import os
import pickle
import sys
import time
from argparse import ArgumentParser

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..", "..")))

import nemo.quantization.utils.quantize_model as qm  # noqa: E402
from nemo.collections.asr.models.ctc_models import EncDecCTCModel  # noqa: E402
from qasr.params import reestimate_bn_, synthetic_features  # noqa: E402

if not torch.cuda.is_available():
    raise Exception("Current implementation only supports GPU")


def hypotheses_from_tokens(tokens, counts, vocabulary):
    """Strings of the greedy CTC hypotheses.  The collapse itself (reference metrics/wer.py:117-136: merge
    repeats, drop blanks) ran on the device (qasr_ctc_collapse, through EncDecCTCModel.transcribe_stream);
    `tokens` (B, T') holds each utterance's label ids compacted to the front, `counts` (B,) how many."""
    return ["".join(vocabulary[c] for c in row[:n]) for row, n in zip(tokens.tolist(), counts.tolist())]


def word_error_rate(hypotheses, references):
    """Reference metrics/wer.py:25-58: total word edit distance / total reference words."""
    scores = words = 0
    for h, r in zip(hypotheses, references):
        h_list, r_list = h.split(), r.split()
        words += len(r_list)
        prev = list(range(len(r_list) + 1))
        for i, hw in enumerate(h_list, 1):
            cur = [i] + [0] * len(r_list)
            for j, rw in enumerate(r_list, 1):
                cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (hw != rw))
            prev = cur
        scores += prev[-1]
    return scores / words if words else float("inf")


def load_dataset(spec, batch_size):
    if spec.startswith("synthetic"):
        parts = spec.split(":")
        n, frames = int(parts[1]) if len(parts) > 1 else 4, int(parts[2]) if len(parts) > 2 else 1600
        return [synthetic_features(batch_size, frames, seed=1000 + i, ragged=True) + (None,) for i in range(n)]
    return torch.load(spec)


def main():
    parser = ArgumentParser()
    parser.add_argument("--asr_model", type=str, default="quartznet15x5", required=True)
    parser.add_argument("--dataset", type=str, required=True, help="path to evaluation data")
    parser.add_argument("--batch_size", type=int, default=8)
    parser.add_argument("--normalize_text", default=True, type=bool)
    parser.add_argument("--shuffle", action="store_true", help="Shuffle test data.")
    parser.add_argument("--load", type=str, default=None, help="load path for the synthetic data")
    parser.add_argument("--percentile", type=float, default=None)
    parser.add_argument("--weight_bit", type=int, default=8)
    parser.add_argument("--act_bit", type=int, default=8)
    parser.add_argument("--dynamic", action="store_true")
    parser.add_argument("--no_quant", action="store_true")
    parser.add_argument("--eval_early_stop", type=int, default=None)
    parser.add_argument("--calib_early_stop", type=int, default=None)
    args = parser.parse_args()
    torch.set_grad_enabled(False)

    torch.manual_seed(0)
    if args.asr_model.endswith(".nemo"):   # reference inference.py:72-74
        print("Using local ASR model from %s" % args.asr_model)
        asr_model = EncDecCTCModel.restore_from(restore_path=args.asr_model)
    else:
        name, _, ckpt = args.asr_model.partition(":")
        asr_model = EncDecCTCModel.from_topology(name.lower().replace("-", ""))
        if ckpt:
            asr_model.load_state_dict(torch.load(ckpt, map_location="cpu"), strict=False)
        else:
            reestimate_bn_(asr_model.encoder)  # random init needs healthy BN statistics
    asr_model = asr_model.cuda()
    test_data = load_dataset(args.dataset, args.batch_size)

    distilled_data = None
    if args.load is not None:
        print("Data loaded from %s" % args.load)
        with open(args.load, "rb") as f:
            distilled_data = pickle.load(f)
        synthetic_batch_size, _, synthetic_seqlen = distilled_data[0].shape
    else:
        assert args.dynamic or args.no_quant, \
            "synthetic data must be loaded unless running with the dynamic quantization mode"

    asr_model.eval()
    asr_model.set_quant_bit(args.weight_bit, mode="weight")
    asr_model.set_quant_bit(args.act_bit, mode="act")
    if args.percentile is not None:
        qm.set_percentile(asr_model, args.percentile)
    if args.no_quant:
        asr_model.set_quant_mode("none")
    else:
        asr_model.encoder.bn_folding()

    if not args.dynamic and not args.no_quant:
        print("Calibrating...")
        qm.calibrate(asr_model)
        length = torch.tensor([synthetic_seqlen] * synthetic_batch_size).cuda()
        for batch_idx, inputs in enumerate(distilled_data):
            if args.calib_early_stop is not None and batch_idx == args.calib_early_stop:
                break
            encoded, encoded_len, encoded_sf = asr_model.encoder(audio_signal=inputs.cuda().float(), length=length)
            asr_model.decoder(encoder_output=encoded, encoder_output_scaling_factor=encoded_sf)

    print("Evaluating...")
    qm.evaluate(asr_model)
    qm.set_dynamic(asr_model, args.dynamic)
    vocabulary = asr_model.decoder.vocabulary
    hypotheses, references = [], []
    audio_s, t0 = 0.0, time.perf_counter()
    # same loop as the reference (inference.py:127-149), with the host<->device copies pipelined around the kernels
    meta = []

    def host_batches():
        for i, (feats, lens, transcripts) in enumerate(test_data):
            if i == args.eval_early_stop:
                break
            meta.append((transcripts, float(lens.sum()) / 100.0))
            yield feats.float().pin_memory(), lens.to(torch.int32).pin_memory()

    for k, (tokens, counts) in enumerate(asr_model.transcribe_stream(host_batches(), collapse=True)):
        hypotheses += hypotheses_from_tokens(tokens, counts, vocabulary)
        transcripts, secs = meta[k]
        if transcripts is not None:
            references += list(transcripts)
        audio_s += secs
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("decoded %d utterances, %.1f audio-seconds in %.3f s (%.0f audio-s/s incl. host)" %
          (len(hypotheses), audio_s, dt, audio_s / dt))
    if references:
        print("WER:", word_error_rate(hypotheses=hypotheses, references=references))
    return dict(hypotheses=hypotheses, model=asr_model, test_data=test_data, args=args)


if __name__ == "__main__":
    main()
