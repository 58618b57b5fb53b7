"""one mel_spectrogram launch at the cfg5 sweep's largest point (for `ncu -k regex:mel_kernel`)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hifigan_b200 as H
b, t = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 262144)
y = torch.rand(b, t, device="cuda") * 1.9 - 0.95
for _ in range(3):
    m = H.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
torch.cuda.synchronize()
print(tuple(m.shape), float(m.mean()))
