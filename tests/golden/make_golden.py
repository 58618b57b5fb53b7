"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

It imports the reference's own src/models.py and src/meldataset.py unmodified, with the two import-time
dependencies that contribute no arithmetic stubbed out (matplotlib: utils.py:3-9; librosa: meldataset.py:7,9),
runs them on CPU in fp32 (torch 2.11.0 / torchaudio 2.11.0) with fixed seeds and stores inputs + outputs.
The reference has no tests or golden vectors of its own, so these files are what pins the
oracle (tests/test_oracle_cpu.py) and, through it, the CUDA kernels (tests/test_gpu_*.py).
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def load_reference(path="/root/reference/src"):
    for name in ["matplotlib", "matplotlib.pylab", "matplotlib.colors", "librosa", "librosa.util",
                 "librosa.filters"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    mpl = sys.modules["matplotlib"]
    mpl.use = lambda *a, **k: None
    mpl.pylab, mpl.colors = sys.modules["matplotlib.pylab"], sys.modules["matplotlib.colors"]
    col = sys.modules["matplotlib.colors"]
    col.BASE_COLORS, col.TABLEAU_COLORS, col.CSS4_COLORS = {}, {}, {}
    sys.modules["librosa.util"].normalize = lambda x, **k: x
    sys.modules["librosa.filters"].mel = lambda *a, **k: None
    sys.path.insert(0, path)
    import env
    import meldataset
    import models
    return models, meldataset, env


def sd_stats(sd):
    """Order-independent fingerprint of a state_dict (for same-seed construction checks)."""
    keys = sorted(sd)
    return np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum()), sd[k].numel()]
                     for k in keys], dtype=np.float64), np.array(keys)


def main():
    from oracle import hifigan_oracle as O
    models, meldataset, env = load_reference()
    torch.set_num_threads(8)

    # 1. small generators: full state_dict + input + output
    for ver in ("tiny", "tiny2"):
        h = env.AttrDict(O.config(ver))
        torch.manual_seed(1234)
        G = models.Generator(h).eval()
        sd = {k: v.detach().numpy() for k, v in G.state_dict().items()}
        torch.manual_seed(0)
        x = torch.randn(2, 80, 24)
        with torch.no_grad():
            y = G(x)
        np.savez_compressed(os.path.join(HERE, f"gen_{ver}.npz"), x=x.numpy(), y=y.numpy(),
                            **{"sd/" + k: v for k, v in sd.items()})
        print(ver, y.shape, float(y.abs().max()))

    # 2. full-size generators: seeded construction, outputs + state_dict fingerprint only
    for ver in ("v1", "v2", "v3"):
        h = env.AttrDict(O.config(ver))
        torch.manual_seed(1234)
        G = models.Generator(h).eval()
        stats, keys = sd_stats(G.state_dict())
        torch.manual_seed(0)
        x = torch.randn(1, 80, 32)
        with torch.no_grad():
            y = G(x)
            # a second weight set that drives tanh into its non-linear range (SURVEY §8d "Weights")
            sd3 = {k: (v * 3 if k.endswith("weight_g") else v) for k, v in G.state_dict().items()}
            G.load_state_dict(sd3)
            y3 = G(x)
        np.savez_compressed(os.path.join(HERE, f"gen_{ver}_seed1234.npz"), x=x.numpy(), y=y.numpy(),
                            y_g3=y3.numpy(), sd_stats=stats, sd_keys=keys)
        print(ver, y.shape, float(y.abs().max()), float(y3.abs().max()))

    # 3. mel_spectrogram: seeded audio-like input, silence, full-scale sine, out-of-range sample
    y = O.synthetic_audio(4, 8192, seed=1)
    t = torch.arange(8192, dtype=torch.float32)
    special = torch.stack([torch.zeros(8192), torch.sin(2 * np.pi * 440.0 * t / 22050),
                           1.5 * torch.sin(2 * np.pi * 1000.0 * t / 22050)])
    odd = O.synthetic_audio(2, 12345, seed=7)  # length not a multiple of hop
    out = {"y": y.numpy(), "special": special.numpy(), "odd": odd.numpy()}
    for name, inp in (("y", y), ("special", special), ("odd", odd)):
        for fmax in (8000, None):
            m = meldataset.mel_spectrogram(inp, 1024, 80, 22050, 256, 1024, 0, fmax)
            out[f"mel_{name}_fmax{fmax}"] = m.numpy()
    np.savez_compressed(os.path.join(HERE, "mel.npz"), **out)
    print("mel", out["mel_y_fmax8000"].shape, out["mel_odd_fmaxNone"].shape)

    # 4. discriminators: seeded construction (G first), logits + fmap fingerprints
    torch.manual_seed(1234)
    _ = models.Generator(env.AttrDict(O.config("v1")))
    mpd = models.MultiPeriodDiscriminator().eval()
    # train mode: the spectral-norm scale then runs its power iterations (eval mode at random init divides
    # by an unconverged sigma and overflows to ~1e24, which pins nothing)
    msd = models.MultiScaleDiscriminator().train()
    ya = O.synthetic_audio(2, 8192, seed=3).unsqueeze(1)
    yb = O.synthetic_audio(2, 8192, seed=4).unsqueeze(1)
    d = {"y": ya.numpy(), "y_hat": yb.numpy()}
    with torch.no_grad():
        for name, D in (("mpd", mpd), ("msd", msd)):
            stats, keys = sd_stats(D.state_dict())  # before the forward: u/v buffers move in train mode
            rs, gs, fr, fg = D(ya, yb)
            d[f"{name}_sd_stats"], d[f"{name}_sd_keys"] = stats, keys
            for i, (r, g) in enumerate(zip(rs, gs)):
                d[f"{name}_logits_r{i}"], d[f"{name}_logits_g{i}"] = r.numpy(), g.numpy()
            d[f"{name}_fmap_abs_mean_r"] = np.array([[float(f.abs().mean()) for f in fl] + [0.0] * (8 - len(fl))
                                                     for fl in fr])
            d[f"{name}_fmap_abs_mean_g"] = np.array([[float(f.abs().mean()) for f in fl] + [0.0] * (8 - len(fl))
                                                     for fl in fg])
            d[f"{name}_feature_loss"] = float(models.feature_loss(fr, fg))
            dl, rl, gl = models.discriminator_loss(rs, gs)
            d[f"{name}_disc_loss"] = float(dl)
            d[f"{name}_disc_r_losses"], d[f"{name}_disc_g_losses"] = np.array(rl), np.array(gl)
            gl_total, gl_parts = models.generator_loss(gs)
            d[f"{name}_gen_loss"] = float(gl_total)
    np.savez_compressed(os.path.join(HERE, "disc_seed1234.npz"), **d)
    print("disc", d["mpd_feature_loss"], d["msd_feature_loss"], d["mpd_disc_loss"], d["msd_gen_loss"])

    # 6. forward half of the UPSTREAM training step, reference modules, B = 2, no optimizer update:
    #    G forward, mel of the generated audio, D-step losses, G-step losses (msd is called twice, so the
    #    spectral-norm scale runs 4 power iterations, as in a real step)
    torch.manual_seed(1234)
    hh = env.AttrDict(O.config("v1"))
    G = models.Generator(hh).train()
    mpd = models.MultiPeriodDiscriminator().train()
    msd = models.MultiScaleDiscriminator().train()
    ya = O.synthetic_audio(2, 8192, seed=5)
    with torch.no_grad():
        x = meldataset.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
        y_mel = meldataset.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
        y = ya.unsqueeze(1)
        y_g_hat = G(x)
        y_g_hat_mel = meldataset.mel_spectrogram(y_g_hat.squeeze(1), 1024, 80, 22050, 256, 1024, 0, None)
        y_df_r, y_df_g, _, _ = mpd(y, y_g_hat)
        loss_disc_f, _, _ = models.discriminator_loss(y_df_r, y_df_g)
        y_ds_r, y_ds_g, _, _ = msd(y, y_g_hat)
        loss_disc_s, _, _ = models.discriminator_loss(y_ds_r, y_ds_g)
        loss_mel = torch.nn.functional.l1_loss(y_mel, y_g_hat_mel) * 45
        _, y_df_g, fmap_f_r, fmap_f_g = mpd(y, y_g_hat)
        _, y_ds_g, fmap_s_r, fmap_s_g = msd(y, y_g_hat)
        t = {"audio": ya.numpy(), "y_g_hat": y_g_hat.numpy(),
             "loss_disc_f": float(loss_disc_f), "loss_disc_s": float(loss_disc_s), "loss_mel": float(loss_mel),
             "loss_fm_f": float(models.feature_loss(fmap_f_r, fmap_f_g)),
             "loss_fm_s": float(models.feature_loss(fmap_s_r, fmap_s_g)),
             "loss_gen_f": float(models.generator_loss(y_df_g)[0]),
             "loss_gen_s": float(models.generator_loss(y_ds_g)[0])}
    np.savez_compressed(os.path.join(HERE, "train_fwd_seed1234.npz"), **t)
    print("train fwd", {k: v for k, v in t.items() if k.startswith("loss")})

    # 5. per-layer known answers for the conv primitives at odd sizes (torch ops the reference calls)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 6, 37, generator=g)
    w = torch.randn(4, 6, 5, generator=g)
    b = torch.randn(4, generator=g)
    wt = torch.randn(6, 3, 8, generator=g)
    bt = torch.randn(3, generator=g)
    wg = torch.randn(4, 3, 7, generator=g)
    k = {"x": x.numpy(), "w": w.numpy(), "b": b.numpy(), "wt": wt.numpy(), "bt": bt.numpy(), "wg": wg.numpy()}
    import torch.nn.functional as F
    k["conv_d3"] = F.conv1d(x, w, b, dilation=3, padding=6).numpy()
    k["convtr_s4"] = F.conv_transpose1d(x, wt, bt, stride=4, padding=2).numpy()
    k["conv_g2_s2"] = F.conv1d(x, wg, None, stride=2, padding=3, groups=2).numpy()
    k["avgpool"] = F.avg_pool1d(x, 4, 2, padding=2).numpy()
    np.savez_compressed(os.path.join(HERE, "conv_kat.npz"), **k)
    print("done")


if __name__ == "__main__":
    main()
