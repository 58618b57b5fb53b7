"""Forward half of the UPSTREAM training step (jik876/hifi-gan train.py; the fork deleted the file but ships every
function it calls — SURVEY.md §3.3), composed from the CUDA-backed modules of this package.

Only the forward graph and the loss values exist so far: the backward kernels (dgrad / wgrad of every conv, mel
backward, weight-norm backward), AdamW and the data-parallel gradient all-reduce are not built.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib
from .meldataset import mel_spectrogram
from .models import _device_mean, discriminator_loss, feature_loss, generator_loss


def step_losses(generator, mpd, msd, x: torch.Tensor, y: torch.Tensor, y_mel: torch.Tensor, h) -> Dict[str, torch.Tensor]:
    """One training step's loss values, call for call as the reference computes them (no optimizer update):

        y_g_hat = generator(x);  y_g_hat_mel = mel_spectrogram(y_g_hat.squeeze(1), ..., h.fmax_for_loss)
        D step:  mpd(y, y_g_hat.detach()), msd(...)  ->  loss_disc_f, loss_disc_s
        G step:  45 * L1(y_mel, y_g_hat_mel), mpd / msd again  ->  feature and generator losses

    x [B,80,F] mel, y [B,1,T] audio, y_mel [B,80,F] loss-mel (all on the GPU).  msd is evaluated twice, so a
    train-mode spectral-norm scale advances 4 power iterations per step exactly like the reference."""
    with torch.no_grad():
        y_g_hat = generator(x).clone()
        y_g_hat_mel = mel_spectrogram(y_g_hat.squeeze(1), h.n_fft, h.num_mels, h.sampling_rate, h.hop_size,
                                      h.win_size, h.fmin, h.fmax_for_loss)
        y_df_r, y_df_g, _, _ = mpd(y, y_g_hat)
        loss_disc_f, _, _ = discriminator_loss(y_df_r, y_df_g)
        y_ds_r, y_ds_g, _, _ = msd(y, y_g_hat)
        loss_disc_s, _, _ = discriminator_loss(y_ds_r, y_ds_g)
        loss_mel = _device_mean(y_mel, y_g_hat_mel, 0, 0.0) * 45
        _, y_df_g, fmap_f_r, fmap_f_g = mpd(y, y_g_hat)
        _, y_ds_g, fmap_s_r, fmap_s_g = msd(y, y_g_hat)
        out = {"y_g_hat": y_g_hat, "loss_disc_f": loss_disc_f, "loss_disc_s": loss_disc_s, "loss_mel": loss_mel,
               "loss_fm_f": feature_loss(fmap_f_r, fmap_f_g), "loss_fm_s": feature_loss(fmap_s_r, fmap_s_g),
               "loss_gen_f": generator_loss(y_df_g)[0], "loss_gen_s": generator_loss(y_ds_g)[0]}
        out["loss_disc_all"] = out["loss_disc_s"] + out["loss_disc_f"]
        out["loss_gen_all"] = (out["loss_gen_s"] + out["loss_gen_f"] + out["loss_fm_s"] + out["loss_fm_f"]
                               + out["loss_mel"])
    return out
