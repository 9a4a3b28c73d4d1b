"""Entry point of the zero-shot calibration-data synthesis, with the reference's flags
(reference examples/asr/quantization/synthesize.py:48-104): load the FP32 teacher, switch quantisation off,
optimise `num_batch` random batches against the BatchNorm statistics, dump them as the pickle that inference.py
reads with --load (`List[FloatTensor(batch_size, 64, seqlen)]`, :94-104).

`--asr_model` is a local .nemo archive or a topology name (no NGC download offline); `--dataset` is accepted for
command-line compatibility and unused, as in the reference (its test loader is set up and never iterated).
"""
import os
import pickle
import sys
from argparse import ArgumentParser

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..", "..")))

from nemo.collections.asr.models.ctc_models import EncDecCTCModel  # noqa: E402
from nemo.quantization.utils.distill_data import get_synthetic_data  # noqa: E402
from qasr.params import reestimate_bn_  # noqa: E402


def main():
    parser = ArgumentParser()
    parser.add_argument("--asr_model", type=str, default="QuartzNet15x5Base-En", required=True)
    parser.add_argument("--dataset", type=str, required=False, default=None, help="path to evaluation data (unused)")
    parser.add_argument("--num_batch", type=int, default=50, help="number of batches of the synthetic data")
    parser.add_argument("--batch_size", type=int, default=8, help="batch size of the synthetic data")
    parser.add_argument("--seqlen", type=int, default=500, help="sequence length of the synthetic data")
    parser.add_argument("--train_iter", type=int, default=200, help="training iterations for the synthetic data generation")
    parser.add_argument("--dump_path", type=str, default=None, help="path to dump the synthetic data")
    parser.add_argument("--dump_prefix", type=str, default="syn", help="prefix for the filename of the dumped synthetic data")
    parser.add_argument("--lr", type=float, default=0.01, help="Learning rate for the synthetic data generation")
    parser.add_argument("--seed", type=int, default=None, help="seed of the initial noise (extension: reproducible runs)")
    args = parser.parse_args()

    torch.set_grad_enabled(False)
    if args.asr_model.endswith(".nemo"):
        print("Using local ASR model from %s" % args.asr_model)
        teacher_model = EncDecCTCModel.restore_from(restore_path=args.asr_model)
    else:
        teacher_model = EncDecCTCModel.from_topology(args.asr_model.lower().replace("-", ""))
        reestimate_bn_(teacher_model.encoder)
    if torch.cuda.is_available():
        teacher_model = teacher_model.cuda()
    teacher_model.set_quant_mode("none")     # disable quantization mode for the teacher model (:80)
    print("Num batches: %d, Batch size: %d, Training iterations: %d, Learning rate: %.3f "
          % (args.num_batch, args.batch_size, args.train_iter, args.lr))
    print("Synthesizing...")
    gen = None if args.seed is None else torch.Generator().manual_seed(args.seed)
    synthetic_data = get_synthetic_data(teacher_model.encoder, teacher_model.decoder, batch_size=args.batch_size, dim=64,
                                        seqlen=args.seqlen, num_batch=args.num_batch, train_iter=args.train_iter,
                                        lr=args.lr, generator=gen)
    file_name = "%s_nb%d_iter%d_lr%.3f.pkl" % (args.dump_prefix, args.num_batch, args.train_iter, args.lr)
    if args.dump_path is not None:
        os.makedirs(args.dump_path, exist_ok=True)
        file_name = os.path.join(args.dump_path, file_name)
    print("Synthetic data dumped as ", file_name)
    with open(file_name, "wb") as f:
        pickle.dump([x.cpu() for x in synthetic_data], f)
    return file_name


if __name__ == "__main__":
    main()
