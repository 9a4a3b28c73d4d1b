"""BASELINE config 4 on one GPU: QuartzNet15x5 W4/A8, dynamic quantisation (every QuantAct takes its range from the
batch it sees, reference quant_modules.py:149-167), through the module-level path (one kernel group per module call,
float code*scale tensors between modules like the reference), timed with CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200")]
import torch
from nemo.collections.asr.models.ctc_models import EncDecCTCModel
import nemo.quantization.utils.quantize_model as qm
from qasr.params import reestimate_bn_, synthetic_features

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = 1600
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = EncDecCTCModel.from_topology("quartznet15x5")
model.eval()
reestimate_bn_(model.encoder)
model.set_quant_bit(4, "weight")
model.set_quant_bit(8, "act")
model.encoder.bn_folding()
model = model.to(dev)
qm.evaluate(model)
qm.set_dynamic(model, True)
x, lens = synthetic_features(B, T, seed=1000)
x, lens = x.to(dev), lens.to(dev)
with torch.no_grad():
    for _ in range(2):
        model(processed_signal=x, processed_signal_length=lens)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        model(processed_signal=x, processed_signal_length=lens)
        b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
print("dynamic W4A8 module-level path: batch %d x %d frames: %.1f ms per forward = %.0f audio-s/s" % (B, T, ms, B * T / 100 / (ms / 1e3)))
