"""mel_spectrogram at the 65 536-frame point: CUDA-event time of the forward kernel (HG_MEL_V1=1 selects the round-1
kernel for A/B).  Tool; prints one JSON line."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
r = bench.measure_mel(torch.device("cuda"), 10)
r["kernel"] = "mel_kernel (round 1)" if os.environ.get("HG_MEL_V1") else "mel_kernel2"
print(json.dumps(r))
