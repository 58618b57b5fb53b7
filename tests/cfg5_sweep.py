"""BASELINE.json configs[4]: V3 Generator inference plus the mel_spectrogram throughput sweep
(n_fft 1024, hop 256, 80 mels, sr 22050, fmin 0, fmax 8000; batch 1..256 x T in {8192, 262144} samples, fp32).

Measurement tool (not a pytest module).  CUDA-event times after warm-up; every input set is larger than L2 or the
L2 is flushed between iterations by writing a 256 MB buffer.  Beside each of our numbers: torchaudio's
MelSpectrogram path (what the reference's meldataset.py:59-71 runs) on the same GPU when torchaudio is importable.
Writes gpurun_out/cfg5_sweep.jsonl.

    python tests/cfg5_sweep.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hifigan_b200 as H                      # noqa: E402
from hifigan_b200.configs import load_config  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "cfg5_sweep.jsonl")
HBM_GBS = 6543.7
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        HBM_GBS = json.load(f)["hbm_gbs"]
except (OSError, KeyError, ValueError):
    pass


def emit(rec):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


def timed(fn, flush, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)           # 256 MB read + write: evicts the 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    return ts[len(ts) // 2]


def mel_sweep():
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    ta = None
    try:
        import torchaudio
        ta = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256,
                                                  f_min=0, f_max=8000, n_mels=80, center=False).cuda()
    except Exception as e:  # noqa: BLE001
        emit({"note": f"torchaudio unavailable: {e}"[:200]})
    for t in (8192, 262144):
        for b in (1, 2, 4, 8, 16, 32, 64, 128, 256):
            y = (torch.rand(b, t, device="cuda") * 1.9 - 0.95)
            ms = timed(lambda: H.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000), flush)
            frames = b * (t // 256)
            rec = {"case": "mel_spectrogram", "batch": b, "t": t, "frames": frames, "ms": ms,
                   "frames_per_s": frames / ms * 1e3, "algorithmic_GBps": frames * 1344 / ms * 1e-6,
                   "frac_of_hbm_peak": frames * 1344 / ms * 1e-6 / HBM_GBS}
            if ta is not None:
                def ref():
                    p = int((1024 - 256) / 2)
                    yy = torch.nn.functional.pad(y.unsqueeze(1), (p, p), mode="reflect").squeeze(1)
                    return torch.log(torch.clamp(ta(yy), min=1e-5))
                rec["torchaudio_ms"] = timed(ref, flush)
            emit(rec)
            H.meldataset.flush_range_warnings(block=True)


def v3_sweep():
    h = load_config("v3")
    torch.manual_seed(1234)
    G = H.Generator(h).cuda().eval()
    G.remove_weight_norm()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    with torch.no_grad():
        for b, frames in ((1, 256), (1, 1024), (8, 1024), (64, 1024), (256, 1024)):
            x = torch.randn(b, 80, frames, device="cuda")
            ms = timed(lambda: G(x), flush if b * frames < 65536 else None, warm=4, iters=8)
            samples = b * frames * 256
            emit({"case": "V3 Generator forward", "batch": b, "frames": frames, "ms": ms,
                  "samples_per_s": samples / ms * 1e3, "xrt_22050": samples / ms * 1e3 / 22050,
                  "tflops": samples * 175648 / ms * 1e-9})


if __name__ == "__main__":
    if os.path.exists(OUT):
        os.remove(OUT)
    emit({"gpu": torch.cuda.get_device_name(0), "hbm_gbs_peak": HBM_GBS})
    mel_sweep()
    v3_sweep()
