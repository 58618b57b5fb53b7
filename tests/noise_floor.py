"""The reference's own bf16 noise floor on a B200 (SURVEY §8d "to be re-calibrated against the
reference's own bf16-autocast noise floor on the box").

Runs the oracle's torch ops (= the reference's composition of cuDNN / cuBLAS / cuFFT calls) on cuda:0 twice — true fp32
(TF32 off) and `torch.autocast(bf16)`, the way a user of the reference would train in bf16 — on the SAME inputs and
weights as tests/test_gpu_named_configs.py, and records how far the two are apart with the metrics the parity tests
use.  The product's tolerances (GRAD_COS_MIN, GRAD_REL_MAX, WAVE_SNR_MIN_DB ...) are then read against these numbers.
Measurement tool, not a pytest module.  Writes gpurun_out/noise_floor.json (copied to profiles/r02_noise_floor.json).

    python tests/noise_floor.py
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hifigan_oracle as O      # noqa: E402
from oracle import train_oracle as TO       # noqa: E402
import hifigan_b200 as H                    # noqa: E402


def _snr(y, ref):
    return 10 * torch.log10((ref - ref.mean()).pow(2).sum() / (y - ref).pow(2).sum()).item()


def main():
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    h = H.AttrDict(O.config("v1"))
    torch.manual_seed(1234)
    G, mpd, msd = H.Generator(h), H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator()
    sds = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in (G, mpd, msd)]
    rec = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    mel_plain = O.mel_spectrogram

    def mel_fp32(a, *args, **kw):          # torch.fft has no bf16 path: the STFT stays fp32 under autocast
        with torch.autocast("cuda", enabled=False):
            return mel_plain(a.float(), *args, **kw)
    # ---- configs[1] inputs (8 of the 64 items: the metric is per-sample, the batch only repeats it)
    ya = O.synthetic_audio(8, 1024 * 256, seed=0).cuda()
    ya3 = O.synthetic_audio(16, 8192, seed=3).cuda()
    torch.set_default_device("cuda")           # the oracle builds its window / filterbank on the default device
    x = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    sd_g = {k: v.cuda() for k, v in sds[0].items()}
    with torch.no_grad():
        ref = O.generator_forward(sd_g, h, x)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = O.generator_forward(sd_g, h, x).float()
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
        ytf = O.generator_forward(sd_g, h, x)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    rec["cfg2_forward"] = {"bf16_autocast": {"max_abs": (y16 - ref).abs().max().item(), "snr_db": _snr(y16, ref)},
                           "tf32": {"max_abs": (ytf - ref).abs().max().item(), "snr_db": _snr(ytf, ref)},
                           "ref_std": ref.std().item()}
    print(json.dumps(rec["cfg2_forward"]), flush=True)
    del ref, y16, ytf
    torch.cuda.empty_cache()
    # ---- configs[2]: one training step at batch 16
    ya = ya3
    x = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    runs = {}
    for mode in ("fp32", "bf16_autocast", "tf32"):
        ref_sds = [TO.leaf_params({k: v.cuda() for k, v in sd.items()}) for sd in sds]
        tf = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf
        O.mel_spectrogram = mel_fp32 if mode == "bf16_autocast" else mel_plain
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
            losses, gg, gp, gs, _, dy = TO.train_step(*ref_sds, h, x, ya.unsqueeze(1), y_mel, update=False)
        runs[mode] = (losses, {**{"g." + k: v for k, v in gg.items()}, **{"mpd." + k: v for k, v in gp.items()},
                               **{"msd." + k: v for k, v in gs.items()}}, dy)
    O.mel_spectrogram = mel_plain
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    l0, g0, dy0 = runs["fp32"]
    rec["cfg3_step_b16"] = {}
    for mode in ("bf16_autocast", "tf32"):
        l1, g1, dy1 = runs[mode]
        rows = []
        for k, ref in g0.items():
            a, b = g1[k].flatten().float(), ref.flatten().float()
            rows.append((F.cosine_similarity(a, b, dim=0).item(), ((a - b).norm() / (b.norm() + 1e-20)).item(), k))
        rows.sort()
        cos = torch.tensor([r[0] for r in rows])
        rel = torch.tensor([r[1] for r in rows])
        bias = [r for r in rows if r[2].endswith(".bias")]
        rec["cfg3_step_b16"][mode] = {
            "losses_rel": {k: abs(l1[k] - l0[k]) / abs(l0[k]) for k in l0},
            "dy_cosine": F.cosine_similarity(dy1.flatten().float(), dy0.flatten().float(), dim=0).item(),
            "n_tensors": len(rows), "min_cosine": rows[0][0], "worst10": rows[:10],
            "cosine_p01_p10_p50": [cos.quantile(q).item() for q in (0.01, 0.1, 0.5)],
            "rel_l2_max_p99_p90_p50": [rel.max().item()] + [rel.quantile(q).item() for q in (0.99, 0.9, 0.5)],
            "n_below_0.999": int((cos < 0.999).sum()), "n_below_0.995": int((cos < 0.995).sum()),
            "min_bias_cosine": bias[0][0] if bias else None}
        print(mode, json.dumps({k: v for k, v in rec["cfg3_step_b16"][mode].items() if k != "worst10"}), flush=True)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "noise_floor.json"), "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
