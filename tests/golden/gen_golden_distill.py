"""Golden vectors of the zero-shot data synthesis from the UNMODIFIED reference (run in the build container):

    python tests/golden/gen_golden_distill.py     # writes tests/golden/distill_tinyqn.npz

The reference's get_synthetic_data (nemo/quantization/utils/distill_data.py:71-162) runs on CPU with three shims that do
not touch its arithmetic: torch.Tensor.cuda = identity (ref_loader), the noise source `_get_random_data` replaced by a
fixed list of batches (the reference draws them in 32 DataLoader workers, not reproducible), and the `verbose` keyword
of ReduceLROnPlateau dropped (removed from torch since the reference's 1.7.1).
"""
import importlib.util
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [HERE, os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle")]

from ref_loader import REF_ROOT, RefModel, load_reference  # noqa: E402
from qasr.params import reestimate_bn_  # noqa: E402
from qasr import topology  # noqa: E402
import specio  # noqa: E402

ref = load_reference()
spec = importlib.util.spec_from_file_location("ref_distill_data",
                                              os.path.join(REF_ROOT, "nemo/quantization/utils/distill_data.py"))
D = importlib.util.module_from_spec(spec)
spec.loader.exec_module(D)

_plateau = torch.optim.lr_scheduler.ReduceLROnPlateau
D.optim.lr_scheduler.ReduceLROnPlateau = lambda opt, **kw: _plateau(opt, **{k: v for k, v in kw.items() if k != "verbose"})

if __name__ == "__main__":
    name = "distill_tinyqn"
    torch.manual_seed(zlib.crc32(name.encode()) % 10000)
    tq = dict(filters=32, kernel=9, repeat=2, n_res_blocks=2, last=48)
    m = RefModel(ref, topology.TOPOLOGIES["tiny_quartznet"](**tq))
    m.eval()
    reestimate_bn_(m.encoder, frames=96)
    m.encoder.set_quant_mode("none")
    for l in m.decoder.decoder_layers:
        l.quant_mode = "none"
    m.decoder.act.quant_mode = "none"
    m.encoder.convs_before_bn = sum((list(b.convs_before_bn) for b in m.encoder.encoder_layers), [])
    state = {k: v.numpy().copy() for k, v in m.state_dict().items()
             if not k.endswith("conv.conv.weight") and "num_batches" not in k}
    g = torch.Generator().manual_seed(7)
    start = [torch.rand(4, 64, 80, generator=g) * 0.6 - 0.3 for _ in range(2)]
    D._get_random_data = lambda batch_size, dim, seqlen: [s.clone() for s in start]
    out = D.get_synthetic_data(m.encoder, m.decoder, batch_size=4, dim=64, seqlen=80, train_iter=6, num_batch=2, lr=0.01)
    case = dict(topology="tiny_quartznet", topology_kwargs=tq, state_dict=state, start=[s.numpy() for s in start],
                refined=[o.numpy() for o in out], train_iter=6, lr=0.01)
    path = os.path.join(HERE, name + ".npz")
    specio.save(path, case)
    print(name, os.path.getsize(path) // 1024, "KB; moved by", float((out[0] - start[0]).abs().max()))
