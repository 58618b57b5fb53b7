"""CPU oracle for the UPSTREAM training step (jik876/hifi-gan train.py restated in SURVEY.md §3.3) — TEST
INFRASTRUCTURE, not product code.  Same rules as hifigan_oracle.py: only tests/, smoke() and bench.py's CPU
baseline legs may import it.

The reference fork deleted train.py but ships every function the step calls (src/models.py:75-282,
src/meldataset.py:56-85); the step below is their composition, differentiated by torch autograd on CPU in fp32
and updated by torch.optim.AdamW exactly as UPSTREAM does:

    optim_g = AdamW(G, lr, betas=[b1, b2]);  optim_d = AdamW(chain(msd, mpd), lr, betas=[b1, b2])
    y_g_hat = G(x);  D step on y_g_hat.detach();  optim_d.step();  G step through the updated D;  optim_g.step()

Pinning: tests/golden/make_golden_train.py runs the REFERENCE's own modules through this same sequence in the
build container and stores losses, per-parameter gradient norms and a few full gradients
(tests/golden/train_step_seed1234.npz); tests/test_oracle_cpu.py checks this restatement against them.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import hifigan_oracle as O

Tensor = torch.Tensor

_PARAM_SUFFIXES = (".weight_g", ".weight_v", ".weight_orig", ".weight", ".bias")


def _is_param(key: str) -> bool:
    # spectral-norm u / v are buffers (weight_v of a spectral-norm layer has a sibling weight_orig)
    return key.endswith(_PARAM_SUFFIXES)


def leaf_params(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Turn the parameters of a state_dict into autograd leaves (buffers stay plain tensors)."""
    out = {}
    for k, v in sd.items():
        is_sn_buffer = k.endswith(".weight_v") and (k[:-len(".weight_v")] + ".weight_orig") in sd
        is_sn_buffer = is_sn_buffer or k.endswith(".weight_u")
        if _is_param(k) and not is_sn_buffer:
            out[k] = v.detach().clone().float().requires_grad_(True)
        else:
            out[k] = v.detach().clone().float()
    return out


def make_optimizers(sd_g, sd_mpd, sd_msd, h):
    """UPSTREAM: AdamW(generator.parameters()), AdamW(chain(msd.parameters(), mpd.parameters())); keep them across steps."""
    pg = [v for v in sd_g.values() if v.requires_grad]
    pd = [v for v in list(sd_msd.values()) + list(sd_mpd.values()) if v.requires_grad]
    return (torch.optim.AdamW(pg, h.learning_rate, betas=[h.adam_b1, h.adam_b2]),
            torch.optim.AdamW(pd, h.learning_rate, betas=[h.adam_b1, h.adam_b2]))


def train_step(sd_g, sd_mpd, sd_msd, h, x: Tensor, y: Tensor, y_mel: Tensor, update: bool = True, optims=None):
    """One UPSTREAM step.  Returns (losses, grads_g, grads_mpd, grads_msd, y_g_hat, dL_gen/dy_g_hat); the
    state_dicts passed in must come from `leaf_params` and are updated in place when `update`.  Pass the same
    `optims = make_optimizers(...)` to consecutive steps so the AdamW moments carry over."""
    optim_g, optim_d = optims if optims is not None else make_optimizers(sd_g, sd_mpd, sd_msd, h)
    mel = lambda a: O.mel_spectrogram(a, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin,
                                      h.fmax_for_loss)
    y_g_hat = O.generator_forward(sd_g, h, x)
    y_g_hat_mel = mel(y_g_hat.squeeze(1))
    losses = {}
    # discriminator step
    optim_d.zero_grad()
    y_df_r, y_df_g, _, _ = O.mpd_forward(sd_mpd, y, y_g_hat.detach())
    loss_disc_f, _, _ = O.discriminator_loss(y_df_r, y_df_g)
    y_ds_r, y_ds_g, _, _ = O.msd_forward(sd_msd, y, y_g_hat.detach(), train=True)
    loss_disc_s, _, _ = O.discriminator_loss(y_ds_r, y_ds_g)
    loss_disc_all = loss_disc_s + loss_disc_f
    loss_disc_all.backward()
    grads_mpd = {k: v.grad.detach().clone() for k, v in sd_mpd.items() if v.requires_grad}
    grads_msd = {k: v.grad.detach().clone() for k, v in sd_msd.items() if v.requires_grad}
    if update:
        optim_d.step()
    # generator step
    optim_g.zero_grad()
    loss_mel = F.l1_loss(y_mel, y_g_hat_mel) * 45
    _, y_df_g, fmap_f_r, fmap_f_g = O.mpd_forward(sd_mpd, y, y_g_hat)
    _, y_ds_g, fmap_s_r, fmap_s_g = O.msd_forward(sd_msd, y, y_g_hat, train=True)
    loss_fm_f, loss_fm_s = O.feature_loss(fmap_f_r, fmap_f_g), O.feature_loss(fmap_s_r, fmap_s_g)
    loss_gen_f, _ = O.generator_loss(y_df_g)
    loss_gen_s, _ = O.generator_loss(y_ds_g)
    loss_gen_all = loss_gen_s + loss_gen_f + loss_fm_s + loss_fm_f + loss_mel
    y_g_hat.retain_grad()
    loss_gen_all.backward()
    grads_g = {k: v.grad.detach().clone() for k, v in sd_g.items() if v.requires_grad}
    if update:
        optim_g.step()
    losses = {"loss_disc_f": loss_disc_f, "loss_disc_s": loss_disc_s, "loss_disc_all": loss_disc_all,
              "loss_mel": loss_mel, "loss_fm_f": loss_fm_f, "loss_fm_s": loss_fm_s, "loss_gen_f": loss_gen_f,
              "loss_gen_s": loss_gen_s, "loss_gen_all": loss_gen_all}
    losses = {k: float(v) for k, v in losses.items()}
    return losses, grads_g, grads_mpd, grads_msd, y_g_hat.detach(), y_g_hat.grad.detach().clone()
