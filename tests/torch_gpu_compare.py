"""Library baseline on the SAME GPU: the oracle's torch ops (cuDNN / cuBLAS / cuFFT through ATen) on cuda:0.

Measurement tool, not a pytest module and not part of the product path ("beside it PyTorch-eager
reference on the same GPU (fp32, TF32-allowed, and bf16-autocast)").  The oracle is executed here as the thing
being compared AGAINST, which tests/ may do.  Writes one JSON line per case to gpurun_out/torch_gpu_compare.jsonl.

    python tests/torch_gpu_compare.py [infer_batch=64] [frames=1024] [train_batch=16]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hifigan_oracle as O      # noqa: E402
from oracle import train_oracle as TO       # noqa: E402
import hifigan_b200 as H                    # noqa: E402  (parameter containers only)
from hifigan_b200.configs import load_config  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "torch_gpu_compare.jsonl")


def emit(rec):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


def timed(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def infer(batch, frames):
    h = load_config("v1")
    torch.manual_seed(1234)
    G = H.Generator(h)
    G.remove_weight_norm()
    sd = {k: v.detach().cuda() for k, v in G.state_dict().items()}
    torch.manual_seed(0)
    x = torch.randn(batch, 80, frames, device="cuda")
    samples = batch * frames * 256
    for name, tf32, autocast in (("fp32", False, False), ("tf32", True, False), ("bf16-autocast", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                return O.generator_forward(sd, h, x)
        try:
            ms = timed(run)
            emit({"case": "torch-cuda V1 Generator forward", "mode": name, "batch": batch, "frames": frames,
                  "ms_per_step": ms, "samples_per_s": samples / ms * 1e3, "xrt_22050": samples / ms * 1e3 / 22050})
        except Exception as e:  # noqa: BLE001
            emit({"case": "torch-cuda V1 Generator forward", "mode": name, "error": f"{type(e).__name__}: {e}"[:300]})
        torch.cuda.empty_cache()


def train(batch):
    h = load_config("v1")
    ya = O.synthetic_audio(batch, 8192, seed=3)
    x = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000).cuda()
    y_mel = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None).cuda()
    ya = ya.cuda()
    torch.set_default_device("cuda")   # the oracle builds its window / filterbank on the default device
    for name, tf32, autocast in (("tf32", True, False), ("bf16-autocast", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.manual_seed(1234)
        with torch.device("cpu"):
            G, mpd, msd = H.Generator(h), H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator()
        sds = [TO.leaf_params({k: v.detach().clone().cuda() for k, v in m.state_dict().items()}) for m in (G, mpd, msd)]
        optims = TO.make_optimizers(*sds, h)
        if autocast and not getattr(O, "_mel_fp32", False):
            # torch.fft has no bf16 path: keep the STFT in fp32 under autocast, as a user of the reference would
            mel_plain = O.mel_spectrogram

            def mel_fp32(a, *args, **kw):
                with torch.autocast("cuda", enabled=False):
                    return mel_plain(a.float(), *args, **kw)
            O.mel_spectrogram = mel_fp32
            O._mel_fp32 = True

        def run():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                TO.train_step(*sds, h, x, ya.unsqueeze(1), y_mel, optims=optims)
        try:
            t0 = time.perf_counter()
            ms = timed(run, warm=3, iters=5)
            emit({"case": "torch-cuda V1 training step (autograd + torch AdamW)", "mode": name, "batch": batch,
                  "ms_per_step": ms, "segments_per_s": batch / ms * 1e3, "wall_s": time.perf_counter() - t0})
        except Exception as e:  # noqa: BLE001
            emit({"case": "torch-cuda V1 training step", "mode": name, "error": f"{type(e).__name__}: {e}"[:300]})
        del sds, optims
        torch.cuda.empty_cache()


if __name__ == "__main__":
    ib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    fr = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    tb = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    emit({"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__})
    infer(ib, fr)
    train(tb)
