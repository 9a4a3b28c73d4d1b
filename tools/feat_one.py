"""Two calls of the mel featurizer on 32 x 20 s of audio (BASELINE config 5's front end): target of ncu launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "q-asr_b200")]
import torch
from nemo.collections.asr.modules.audio_preprocessing import AudioToMelSpectrogramPreprocessor
dev = torch.device("cuda:0")
pre = AudioToMelSpectrogramPreprocessor(dither=0.0).to(dev)
a = 0.1 * torch.randn(32, 320000, device=dev)
n = torch.full((32,), 320000, dtype=torch.long, device=dev)
for _ in range(2):
    f, l = pre(input_signal=a, length=n)
torch.cuda.synchronize()
print(f.shape)
