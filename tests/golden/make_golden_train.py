"""Golden fixture for the full UPSTREAM training step, generated FROM THE REFERENCE'S OWN MODULES.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_train.py

The fork deleted train.py; the step composed here is UPSTREAM's loop body over the reference's
unmodified src/models.py / src/meldataset.py classes, differentiated by torch autograd and updated by
torch.optim.AdamW on CPU in fp32.  Two consecutive steps on the same batch (B = 2 segments of 8192 samples,
V1 config, torch.manual_seed(1234) construction order G, MPD, MSD):

  step 1: losses, dL_gen/dy_g_hat, the L2 norm of every parameter gradient (D grads from the D step, G grads from
          the G step) and a few small gradients in full
  step 2: losses only (they depend on both AdamW updates of step 1)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402

FULL = ["conv_post.weight_v", "conv_post.weight_g", "conv_post.bias", "conv_pre.bias", "ups.3.bias", "ups.3.weight_g",
        "resblocks.11.convs2.2.bias", "resblocks.0.convs1.0.weight_g"]
FULL_MPD = ["discriminators.0.convs.0.weight_v", "discriminators.0.convs.0.bias", "discriminators.4.conv_post.weight_v",
            "discriminators.2.convs.1.bias"]
FULL_MSD = ["discriminators.0.convs.0.weight_orig", "discriminators.0.conv_post.bias", "discriminators.1.convs.0.weight_v",
            "discriminators.2.conv_post.weight_v", "discriminators.1.convs.2.bias"]


def main():
    from oracle import hifigan_oracle as O
    models, meldataset, env = load_reference()
    torch.set_num_threads(8)
    torch.manual_seed(1234)
    h = env.AttrDict(O.config("v1"))
    G = models.Generator(h).train()
    mpd = models.MultiPeriodDiscriminator().train()
    msd = models.MultiScaleDiscriminator().train()
    import itertools
    optim_g = torch.optim.AdamW(G.parameters(), h.learning_rate, betas=[h.adam_b1, h.adam_b2])
    optim_d = torch.optim.AdamW(itertools.chain(msd.parameters(), mpd.parameters()), h.learning_rate,
                                betas=[h.adam_b1, h.adam_b2])
    ya = O.synthetic_audio(2, 8192, seed=5)
    mel = lambda a, fmax: meldataset.mel_spectrogram(a, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size,
                                                     h.fmin, fmax)
    with torch.no_grad():
        x = mel(ya, h.fmax)
        y_mel = mel(ya, h.fmax_for_loss)
    y = ya.unsqueeze(1)
    out = {"audio": ya.numpy()}
    F = torch.nn.functional
    for step in (1, 2):
        y_g_hat = G(x)
        y_g_hat_mel = mel(y_g_hat.squeeze(1), h.fmax_for_loss)
        optim_d.zero_grad()
        y_df_r, y_df_g, _, _ = mpd(y, y_g_hat.detach())
        loss_disc_f, _, _ = models.discriminator_loss(y_df_r, y_df_g)
        y_ds_r, y_ds_g, _, _ = msd(y, y_g_hat.detach())
        loss_disc_s, _, _ = models.discriminator_loss(y_ds_r, y_ds_g)
        (loss_disc_s + loss_disc_f).backward()
        if step == 1:
            for name, net, full in (("mpd", mpd, FULL_MPD), ("msd", msd, FULL_MSD)):
                keys = [k for k, _ in net.named_parameters()]
                out[f"{name}_keys"] = np.array(keys)
                out[f"{name}_grad_norm"] = np.array([float(p.grad.norm()) for _, p in net.named_parameters()])
                for k in full:
                    out[f"{name}_grad::{k}"] = dict(net.named_parameters())[k].grad.numpy().copy()
        optim_d.step()
        optim_g.zero_grad()
        loss_mel = F.l1_loss(y_mel, y_g_hat_mel) * 45
        _, y_df_g, fmap_f_r, fmap_f_g = mpd(y, y_g_hat)
        _, y_ds_g, fmap_s_r, fmap_s_g = msd(y, y_g_hat)
        loss_fm_f, loss_fm_s = models.feature_loss(fmap_f_r, fmap_f_g), models.feature_loss(fmap_s_r, fmap_s_g)
        loss_gen_f, _ = models.generator_loss(y_df_g)
        loss_gen_s, _ = models.generator_loss(y_ds_g)
        loss_gen_all = loss_gen_s + loss_gen_f + loss_fm_s + loss_fm_f + loss_mel
        y_g_hat.retain_grad()
        loss_gen_all.backward()
        if step == 1:
            keys = [k for k, _ in G.named_parameters()]
            out["g_keys"] = np.array(keys)
            out["g_grad_norm"] = np.array([float(p.grad.norm()) for _, p in G.named_parameters()])
            for k in FULL:
                out[f"g_grad::{k}"] = dict(G.named_parameters())[k].grad.numpy().copy()
            out["dy_g_hat"] = y_g_hat.grad.numpy().copy()
            out["y_g_hat"] = y_g_hat.detach().numpy().copy()
        optim_g.step()
        vals = {"loss_disc_f": loss_disc_f, "loss_disc_s": loss_disc_s, "loss_mel": loss_mel, "loss_fm_f": loss_fm_f,
                "loss_fm_s": loss_fm_s, "loss_gen_f": loss_gen_f, "loss_gen_s": loss_gen_s}
        for k, v in vals.items():
            out[f"step{step}_{k}"] = float(v)
        print("step", step, {k: round(float(v), 5) for k, v in vals.items()})
    np.savez_compressed(os.path.join(HERE, "train_step_seed1234.npz"), **out)
    print("written", sum(v.nbytes for v in out.values() if hasattr(v, "nbytes")), "bytes")


if __name__ == "__main__":
    main()
