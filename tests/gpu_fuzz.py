"""Randomised shape sweep of the CUDA path against the oracle (tool, not a pytest module).

    python tests/gpu_fuzz.py [seed=0] [cases=12]

Draws batch sizes, frame counts, audio lengths and segment counts the fixed tests do not hit (lengths that are not
multiples of the 128-row tile, of a period, of the hop; batch sizes 1..5) and reports the worst error per family.
Exit code 1 on a tolerance failure.  Tolerances are those of the fixed tests (tests/test_gpu_parity.py,
tests/test_gpu_train.py).
"""
import os
import random
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hifigan_b200 as H                     # noqa: E402
from oracle import hifigan_oracle as O      # noqa: E402
from oracle import train_oracle as TO       # noqa: E402

fails = []


def check(ok, what):
    if not ok:
        fails.append(what)
        print("FAIL", what, flush=True)


def generators(rng, n):
    worst = 0.0
    for ver in ("v1", "v3", "tiny"):
        h = H.AttrDict(O.config(ver))
        torch.manual_seed(1234)
        G = H.Generator(h).cuda().eval()
        sd = {k: v.detach().cpu() for k, v in G.state_dict().items()}
        up = 1
        for u in h.upsample_rates:
            up *= u
        for _ in range(n):
            b, frames = rng.randint(1, 5), rng.choice([1, 2, 3, 5, 17, 31, 33, 64, 97, 127, 129, 200])
            x = torch.randn(b, h.num_mels if "num_mels" in h else 80, frames)
            with torch.no_grad():
                y = G(x.cuda()).cpu()
                ref = O.generator_forward(sd, h, x)
            err = (y - ref).abs().max().item()
            worst = max(worst, err)
            check(y.shape == ref.shape == (b, 1, frames * up) and err < 2e-3, f"generator {ver} b={b} frames={frames} err={err:.2e}")
    print(f"generators: worst max-abs {worst:.2e}", flush=True)


def discriminators(rng, n):
    torch.manual_seed(1234)
    mpd, msd = H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator()
    sd_p = {k: v.detach().clone() for k, v in mpd.state_dict().items()}
    sd_s = {k: v.detach().clone() for k, v in msd.state_dict().items()}
    mpd, msd = mpd.cuda().eval(), msd.cuda().eval()
    worst = 0.0
    for _ in range(n):
        b, t = rng.randint(1, 4), rng.choice([2048, 4099, 6001, 8191, 8192, 8193, 11111, 16384, 22051])
        y = O.synthetic_audio(b, t, seed=rng.randint(0, 999)).unsqueeze(1)
        y_hat = O.synthetic_audio(b, t, seed=rng.randint(0, 999)).unsqueeze(1)
        with torch.no_grad():
            ref_p, ref_s = O.mpd_forward(sd_p, y, y_hat), O.msd_forward(sd_s, y, y_hat, train=False)
            got_p, got_s = mpd(y.cuda(), y_hat.cuda()), msd(y.cuda(), y_hat.cuda())
        for name, ref, got in (("mpd", ref_p, got_p), ("msd", ref_s, got_s)):
            for lr, lg in zip(ref[0] + ref[1], got[0] + got[1]):
                e = (lg.cpu() - lr).abs().max().item() / max(1.0, lr.abs().max().item())
                worst = max(worst, e)
                check(lr.shape == lg.shape and e < 3e-2, f"{name} logits b={b} t={t} err={e:.2e}")
            for fl_ref, fl_got in zip(ref[2] + ref[3], got[2] + got[3]):
                for fr, fg in zip(fl_ref, fl_got):
                    e = (fg.cpu() - fr).abs().max().item() / (fr.abs().max().item() + 1e-6)
                    worst = max(worst, e)
                    check(fr.shape == fg.shape and e < 4e-2, f"{name} fmap {tuple(fr.shape)} b={b} t={t} err={e:.2e}")
    print(f"discriminators: worst relative error {worst:.2e}", flush=True)


def mels(rng, n):
    worst = 0.0
    for _ in range(n):
        b, t = rng.randint(1, 6), rng.choice([1000, 4097, 8192, 12345, 30001])
        n_fft, hop, win = rng.choice([(1024, 256, 1024), (1024, 256, 800), (1024, 128, 1024), (512, 128, 512), (400, 160, 400)])
        fmax = rng.choice([8000, None])
        y = O.synthetic_audio(b, t, seed=rng.randint(0, 999))
        ref = O.mel_spectrogram(y.double(), n_fft, 80, 22050, hop, win, 0, fmax)
        got = H.mel_spectrogram(y.cuda(), n_fft, 80, 22050, hop, win, 0, fmax).cpu().double()
        peak = ref.max(dim=1, keepdim=True).values
        loud = ref > peak - 13.8
        e = (got - ref).abs()[loud].max().item()
        worst = max(worst, e)
        check(got.shape == ref.shape and e < 2e-4, f"mel n_fft={n_fft} hop={hop} win={win} fmax={fmax} b={b} t={t} err={e:.2e}")
    print(f"mel: worst log-domain error (bins within 60 dB of the frame peak) {worst:.2e}", flush=True)


def train_steps(rng, n):
    from hifigan_b200.train import TrainStep
    worst_loss, worst_cos = 0.0, 1.0
    for i in range(n):
        ver = rng.choice(["v1", "v3"])
        b = rng.choice([1, 2, 3, 5])
        h = H.AttrDict(O.config(ver))
        torch.manual_seed(rng.randint(0, 9999))
        G, mpd, msd = H.Generator(h), H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator()
        sds = [TO.leaf_params({k: v.detach().clone() for k, v in m.state_dict().items()}) for m in (G, mpd, msd)]
        ts = TrainStep(G, mpd, msd, h, "cuda")
        ya = O.synthetic_audio(b, 8192, seed=rng.randint(0, 999))
        x = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
        y_mel = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
        losses, gg, gp, gs, _, _ = TO.train_step(*sds, h, x, ya.unsqueeze(1), y_mel, update=False)
        out = ts.step(x.cuda(), ya.cuda().unsqueeze(1), y_mel.cuda(), update=False)
        for k, ref in losses.items():
            if k in out:
                e = abs(out[k].item() - ref) / (abs(ref) + 1e-9)
                worst_loss = max(worst_loss, e)
                check(e < 2e-2, f"train {ver} b={b} {k} got={out[k].item():.5f} ref={ref:.5f}")
        # update=False: the D gradients of the D step and the G gradients of the G step are both still in place
        for net, grads in ((G, gg), (mpd, gp), (msd, gs)):
            for k, p in net.named_parameters():
                got, ref = p.grad.cpu().flatten(), grads[k].flatten()
                cos = F.cosine_similarity(got, ref, dim=0).item()
                worst_cos = min(worst_cos, cos)
                check(cos >= 0.99, f"train {ver} b={b} grad {k} cosine={cos:.4f}")
        del ts
        torch.cuda.empty_cache()
    print(f"training step: worst loss rel err {worst_loss:.2e}, worst gradient cosine {worst_cos:.4f}", flush=True)


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rng = random.Random(seed)
    torch.manual_seed(seed)
    torch.set_num_threads(os.cpu_count() or 1)
    generators(rng, max(2, cases // 3))
    discriminators(rng, max(2, cases // 3))
    mels(rng, cases)
    train_steps(rng, max(2, cases // 4))
    print("failures:", len(fails))
    sys.exit(1 if fails else 0)
