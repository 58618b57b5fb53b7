"""The module API under torch autograd (VERDICT r1 "missing" item 1): the literal UPSTREAM train.py loop body —
`loss.backward(); optim.step()` with torch.optim.AdamW through `Generator`, `mpd`, `msd`, `mel_spectrogram` and the
three loss functions (reference src/models.py:100-116, 175-188, 232-282, src/meldataset.py:56-85) —
against the training golden generated from the REFERENCE's own modules (tests/golden/train_step_seed1234.npz).
"""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

LOSS_KEYS = ("loss_disc_f", "loss_disc_s", "loss_mel", "loss_fm_f", "loss_fm_s", "loss_gen_f", "loss_gen_s")


@pytest.fixture(scope="module")
def H():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import hifigan_b200
    hifigan_b200._lib.lib()
    return hifigan_b200


def _upstream_step(H, h, generator, mpd, msd, optim_g, optim_d, x, y, y_mel):
    """UPSTREAM train.py's loop body, verbatim in structure."""
    from hifigan_b200 import mel_spectrogram, feature_loss, generator_loss, discriminator_loss
    y_g_hat = generator(x)
    y_g_hat_mel = mel_spectrogram(y_g_hat.squeeze(1), h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size,
                                  h.fmin, h.fmax_for_loss)
    optim_d.zero_grad()
    y_df_hat_r, y_df_hat_g, _, _ = mpd(y, y_g_hat.detach())
    loss_disc_f, _, _ = discriminator_loss(y_df_hat_r, y_df_hat_g)
    y_ds_hat_r, y_ds_hat_g, _, _ = msd(y, y_g_hat.detach())
    loss_disc_s, _, _ = discriminator_loss(y_ds_hat_r, y_ds_hat_g)
    loss_disc_all = loss_disc_s + loss_disc_f
    loss_disc_all.backward()
    grads_d = {n: p.grad.detach().clone() for net, nm in ((mpd, "mpd"), (msd, "msd"))
               for n, p in ((f"{nm}::{k}", v) for k, v in net.named_parameters())}
    optim_d.step()
    optim_g.zero_grad()
    loss_mel = F.l1_loss(y_mel, y_g_hat_mel) * 45
    y_df_hat_r, y_df_hat_g, fmap_f_r, fmap_f_g = mpd(y, y_g_hat)
    y_ds_hat_r, y_ds_hat_g, fmap_s_r, fmap_s_g = msd(y, y_g_hat)
    loss_fm_f = feature_loss(fmap_f_r, fmap_f_g)
    loss_fm_s = feature_loss(fmap_s_r, fmap_s_g)
    loss_gen_f, _ = generator_loss(y_df_hat_g)
    loss_gen_s, _ = generator_loss(y_ds_hat_g)
    loss_gen_all = loss_gen_s + loss_gen_f + loss_fm_s + loss_fm_f + loss_mel
    y_g_hat.retain_grad()
    loss_gen_all.backward()
    grads_g = {k: p.grad.detach().clone() for k, p in generator.named_parameters()}
    optim_g.step()
    losses = dict(loss_disc_f=loss_disc_f, loss_disc_s=loss_disc_s, loss_mel=loss_mel, loss_fm_f=loss_fm_f,
                  loss_fm_s=loss_fm_s, loss_gen_f=loss_gen_f, loss_gen_s=loss_gen_s)
    return {k: v.item() for k, v in losses.items()}, grads_g, grads_d, y_g_hat.grad.detach().clone()


def test_upstream_loop_with_torch_adamw_vs_reference_golden(H):
    """Two consecutive UPSTREAM steps with torch.optim.AdamW on the module API against the reference golden: the 7
    losses of both steps (step 2 pins both optimizer updates and the re-packing of the changed weights), dL/dy_g_hat,
    every gradient norm and the gradients stored in full.  Same tolerances as the TrainStep golden test."""
    from conftest import load_npz
    from oracle import hifigan_oracle as O
    z = load_npz("train_step_seed1234.npz")
    h = H.AttrDict(O.config("v1"))
    torch.manual_seed(1234)
    generator, mpd, msd = H.Generator(h).cuda(), H.MultiPeriodDiscriminator().cuda(), H.MultiScaleDiscriminator().cuda()
    optim_g = torch.optim.AdamW(generator.parameters(), h.learning_rate, betas=[h.adam_b1, h.adam_b2])
    optim_d = torch.optim.AdamW(itertools.chain(msd.parameters(), mpd.parameters()), h.learning_rate,
                                betas=[h.adam_b1, h.adam_b2])
    generator.train(); mpd.train(); msd.train()
    ya = torch.from_numpy(z["audio"]).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    y = ya.unsqueeze(1)
    for step in (1, 2):
        losses, gg, gd, dy = _upstream_step(H, h, generator, mpd, msd, optim_g, optim_d, x, y, y_mel)
        for k in LOSS_KEYS:
            ref = float(z[f"step{step}_{k}"])
            assert abs(losses[k] - ref) <= 2e-2 * abs(ref), (step, k, losses[k], ref)
        if step > 1:
            continue
        ref_dy = torch.from_numpy(z["dy_g_hat"]).flatten()
        assert F.cosine_similarity(dy.cpu().flatten(), ref_dy, dim=0).item() >= 0.998
        for name, grads in (("g", gg), ("mpd", {k.split("::")[1]: v for k, v in gd.items() if k.startswith("mpd::")}),
                            ("msd", {k.split("::")[1]: v for k, v in gd.items() if k.startswith("msd::")})):
            for k, n in zip([str(k) for k in z[f"{name}_keys"]], z[f"{name}_grad_norm"]):
                got = grads[k].norm().item()
                assert abs(got - n) <= 5e-2 * n + 1e-9, (name, k, got, n)
            for key in z.files:
                if not key.startswith(f"{name}_grad::"):
                    continue
                k = key.split("_grad::")[1]
                got, ref = grads[k].cpu().flatten(), torch.from_numpy(z[key]).flatten()
                assert F.cosine_similarity(got, ref, dim=0).item() >= 0.999, key
                assert (got - ref).norm() <= 5e-2 * ref.norm(), key


def test_autograd_path_matches_trainstep(H):
    """The same step through the two front doors — the autograd module API and TrainStep — gives the same losses and
    gradients (they drive the same kernels; fp32 atomics make them non-bit-identical)."""
    from oracle import hifigan_oracle as O
    from hifigan_b200.train import TrainStep
    h = H.AttrDict(O.config("v1"))
    ya = O.synthetic_audio(3, 8192, seed=17).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    torch.manual_seed(1234)
    nets_a = [H.Generator(h).cuda(), H.MultiPeriodDiscriminator().cuda(), H.MultiScaleDiscriminator().cuda()]
    torch.manual_seed(1234)
    nets_b = [H.Generator(h), H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator()]
    for n in nets_a:
        n.train()
    optim_g = torch.optim.AdamW(nets_a[0].parameters(), h.learning_rate, betas=[h.adam_b1, h.adam_b2])
    optim_d = torch.optim.AdamW(itertools.chain(nets_a[2].parameters(), nets_a[1].parameters()), h.learning_rate,
                                betas=[h.adam_b1, h.adam_b2])
    la, gg, gd, _ = _upstream_step(H, h, *nets_a, optim_g, optim_d, x, ya.unsqueeze(1), y_mel)
    ts = TrainStep(*nets_b, h, "cuda")
    out = ts.step(x, ya.unsqueeze(1), y_mel)
    for k in LOSS_KEYS:
        assert abs(la[k] - out[k].item()) <= 1e-3 * abs(la[k]) + 1e-5, (k, la[k], out[k].item())
    # the generator gradients survive TrainStep's update (flat.g keeps them until the next step)
    for k, p in nets_b[0].named_parameters():
        a, b = gg[k].flatten(), p.grad.flatten()
        assert F.cosine_similarity(a, b, dim=0).item() >= 0.9999, k
    # after the optimizer updates both parameter sets moved the same way (torch.optim.AdamW vs the fused kernel).
    # AdamW's first update is ~ lr * sign(g): an element whose gradient is inside the atomics' reordering noise
    # may take either sign, so this is a direction check over the whole network, not an element-wise one.
    torch.manual_seed(1234)
    init = torch.cat([p.detach().flatten() for p in H.Generator(h).parameters()]).cuda()
    da = torch.cat([p.detach().flatten() for p in nets_a[0].parameters()]) - init
    db = torch.cat([p.detach().flatten() for p in nets_b[0].parameters()]) - init
    assert da.abs().max().item() <= 2.5e-4 and F.cosine_similarity(da, db, dim=0).item() >= 0.99


def test_autograd_input_gradients_and_frozen_discriminator(H):
    """Gradient at the Generator's input mel and at a discriminator's input audio with frozen parameters — the
    pieces a feature-matching / perceptual-loss user of the modules needs — against autograd over the fp32 oracle."""
    from oracle import hifigan_oracle as O
    h = H.AttrDict(O.config("v1"))
    torch.manual_seed(5)
    G = H.Generator(h).cuda().train()
    sd = {k: v.detach().clone().cpu() for k, v in G.state_dict().items()}
    x = torch.randn(2, 80, 16, generator=torch.Generator().manual_seed(1))
    w = torch.randn(2, 1, 4096, generator=torch.Generator().manual_seed(2))
    xr = x.clone().requires_grad_(True)
    (O.generator_forward(sd, h, xr) * w).sum().backward()
    xc = x.cuda().requires_grad_(True)
    (G(xc) * w.cuda()).sum().backward()
    cos = F.cosine_similarity(xc.grad.cpu().flatten(), xr.grad.flatten(), dim=0).item()
    assert xc.grad.shape == x.shape and cos >= 0.995, cos
    # a period discriminator with requires_grad_(False): only the data gradient runs
    d = H.DiscriminatorP(3).cuda().train()
    sdd = {"discriminators.0." + k: v.detach().clone().cpu() for k, v in d.state_dict().items()}
    for p in d.parameters():
        p.requires_grad_(False)
    ya = O.synthetic_audio(2, 4000, seed=3).unsqueeze(1)
    yr = ya.clone().requires_grad_(True)
    logit_r, fmap_r = O.discriminator_p_forward(sdd, "discriminators.0", yr, 3)
    (logit_r.pow(2).mean() + sum(f.abs().mean() for f in fmap_r)).backward()
    yc = ya.cuda().requires_grad_(True)
    logit, fmap = d(yc)
    (logit.pow(2).mean() + sum(f.abs().mean() for f in fmap)).backward()
    assert all(p.grad is None for p in d.parameters())
    cos = F.cosine_similarity(yc.grad.cpu().flatten(), yr.grad.flatten(), dim=0).item()
    assert cos >= 0.995, cos
