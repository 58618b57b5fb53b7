"""The oracle pinned against every golden fixture generated from the reference (tests/golden/make_golden.py),
and the torch primitives it composes pinned against the plain-C restatement (oracle/conv_ref.c)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_npz
from oracle import conv_ref
from oracle import hifigan_oracle as O


def _sd(npz):
    return {k[3:]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith("sd/")}


@pytest.mark.parametrize("ver", ["tiny", "tiny2"])
def test_generator_small_vs_reference_output(ver):
    z = load_npz(f"gen_{ver}.npz")
    y = O.generator_forward(_sd(z), O.config(ver), torch.from_numpy(z["x"]))
    assert y.shape == z["y"].shape
    assert np.abs(y.numpy() - z["y"]).max() < 2e-7


@pytest.mark.parametrize("ver", ["tiny", "tiny2"])
def test_generator_fold_is_exact(ver):
    """remove_weight_norm (models.py:118-125) must not change the function."""
    z = load_npz(f"gen_{ver}.npz")
    sd = _sd(z)
    folded = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            p = k[: -len(".weight_g")]
            folded[p + ".weight"] = O.weight_of(sd, p)
        elif not k.endswith(".weight_v"):
            folded[k] = v
    x = torch.from_numpy(z["x"])
    a = O.generator_forward(sd, O.config(ver), x)
    b = O.generator_forward(folded, O.config(ver), x)
    assert (a - b).abs().max().item() == 0.0


def test_generator_fp64_agrees_with_fp32():
    z = load_npz("gen_tiny.npz")
    sd64 = {k: v.double() for k, v in _sd(z).items()}
    y = O.generator_forward(sd64, O.config("tiny"), torch.from_numpy(z["x"]).double())
    assert np.abs(y.numpy() - z["y"]).max() < 1e-6


def mel_close(got, ref, tol):
    """log-mel comparison.  The reference computes in fp32 (torchaudio): bins more than ~60 dB below a frame's
    peak (pure tones) carry the reference's own FFT rounding noise, so those are compared in the linear
    power domain relative to the frame peak; everything else in the log domain."""
    peak = ref.max(axis=1, keepdims=True)
    loud = ref > peak - np.log(1e6)
    assert np.abs(got - ref)[loud].max() < tol
    lin = np.abs(np.exp(got) - np.exp(ref)) / np.exp(peak)
    assert lin.max() < tol


@pytest.mark.parametrize("name", ["y", "special", "odd"])
@pytest.mark.parametrize("fmax", [8000, None])
def test_mel_vs_reference(name, fmax):
    z = load_npz("mel.npz")
    ref = z[f"mel_{name}_fmax{fmax}"]
    got = O.mel_spectrogram(torch.from_numpy(z[name]).double(), 1024, 80, 22050, 256, 1024, 0, fmax)
    assert got.shape == ref.shape
    mel_close(got.numpy(), ref, 1e-4)
    got32 = O.mel_spectrogram(torch.from_numpy(z[name]), 1024, 80, 22050, 256, 1024, 0, fmax)
    mel_close(got32.numpy(), ref, 2e-4)


@pytest.mark.parametrize("case", range(4))
def test_mel_other_shapes_vs_reference(case):
    """the oracle's restatement at the shapes of the reference's other callers (n_fft != 1024, 16 kHz, fmax None,
    win < n_fft, fmin > 0) against outputs of the reference's own mel_spectrogram"""
    z = load_npz("mel_other.npz")
    n_fft, nm, sr, hop, win, fmin, fmax = [int(v) for v in z["cases"][case]]
    got = O.mel_spectrogram(torch.from_numpy(z["y"]).double(), n_fft, nm, sr, hop, win, fmin, None if fmax < 0 else fmax)
    assert got.shape == z[f"mel_{case}"].shape
    mel_close(got.numpy(), z[f"mel_{case}"], 1e-4)


def test_mel_silence_hits_the_clamp():
    z = load_npz("mel.npz")
    got = O.mel_spectrogram(torch.from_numpy(z["special"][:1]).double(), 1024, 80, 22050, 256, 1024, 0, 8000)
    assert torch.allclose(got, torch.full_like(got, float(np.log(1e-5))))


def test_fbank_sparsity_matches_survey():
    fb = O.melscale_fbanks_htk(513, 0.0, 8000.0, 80, 22050)
    assert int((fb > 0).sum()) == 729
    assert int(torch.nonzero(fb.sum(1))[-1]) == 371


def test_conv_primitives_vs_c_and_golden():
    z = load_npz("conv_kat.npz")
    x, w, b = z["x"], z["w"], z["b"]
    assert np.abs(conv_ref.conv1d(x, w, b, padding=6, dilation=3) - z["conv_d3"]).max() < 1e-5
    assert np.abs(conv_ref.conv_transpose1d(x, z["wt"], z["bt"], stride=4, padding=2) - z["convtr_s4"]).max() < 1e-5
    assert np.abs(conv_ref.conv1d(x, z["wg"], None, stride=2, padding=3, groups=2) - z["conv_g2_s2"]).max() < 1e-5
    assert np.abs(conv_ref.avg_pool1d(x, 4, 2, 2) - z["avgpool"]).max() < 1e-6
    # and torch today still agrees with the stored answers
    tx = torch.from_numpy(x)
    assert np.abs(F.conv1d(tx, torch.from_numpy(w), torch.from_numpy(b), dilation=3, padding=6).numpy()
                  - z["conv_d3"]).max() < 1e-5


def test_avgpool_edges_include_padding():
    y = conv_ref.avg_pool1d(np.ones((1, 1, 8), np.float32), 4, 2, 2)
    assert y.shape == (1, 1, 5) and y[0, 0, 0] == 0.5 and y[0, 0, -1] == 0.5


def test_polyphase_identity_of_conv_transpose():
    """ConvTranspose1d(k, u, pad=(k-u)/2) == 3-shift conv producing u phases per input step — the packing
    hg_pack_convtr1d_weight implements (include/hifigan_b200.h)."""
    g = torch.Generator().manual_seed(3)
    for (k, u) in [(16, 8), (4, 2), (8, 4)]:
        cin, cout, t = 5, 3, 11
        pad = (k - u) // 2
        x = torch.randn(2, cin, t, generator=g, dtype=torch.float64)
        w = torch.randn(cin, cout, k, generator=g, dtype=torch.float64)
        ref = F.conv_transpose1d(x, w, None, stride=u, padding=pad)
        shifts = (-1, 0, 1)
        wc = torch.zeros(len(shifts), u * cout, cin, dtype=torch.float64)
        for si, s in enumerate(shifts):
            for p in range(u):
                j = p + pad - s * u
                if 0 <= j < k:
                    wc[si, p * cout:(p + 1) * cout] = w[:, :, j].t()
        xp = F.pad(x, (1, 1))
        out = sum(torch.einsum("nc,bct->bnt", wc[si], xp[:, :, si:si + t]) for si in range(3))
        out = out.view(2, u, cout, t).permute(0, 2, 3, 1).reshape(2, cout, t * u)
        assert (out - ref).abs().max().item() < 1e-12


def _seeded_discriminator_sds():
    import hifigan_b200 as H
    torch.manual_seed(1234)
    H.Generator(H.AttrDict(O.config("v1")))
    mpd = H.MultiPeriodDiscriminator()
    msd = H.MultiScaleDiscriminator()
    return ({k: v.detach().clone() for k, v in mpd.state_dict().items()},
            {k: v.detach().clone() for k, v in msd.state_dict().items()})


def test_discriminators_and_losses_vs_reference():
    z = load_npz("disc_seed1234.npz")
    torch.set_num_threads(8)
    mpd_sd, msd_sd = _seeded_discriminator_sds()
    y, y_hat = torch.from_numpy(z["y"]), torch.from_numpy(z["y_hat"])
    with torch.no_grad():
        for name, out in (("mpd", O.mpd_forward(mpd_sd, y, y_hat)),
                          ("msd", O.msd_forward(msd_sd, y, y_hat, train=True))):
            rs, gs, fr, fg = out
            for i, (r, g) in enumerate(zip(rs, gs)):
                ref_r, ref_g = z[f"{name}_logits_r{i}"], z[f"{name}_logits_g{i}"]
                assert r.shape == ref_r.shape
                assert np.abs(r.numpy() - ref_r).max() < 1e-4 * max(1.0, np.abs(ref_r).max())
                assert np.abs(g.numpy() - ref_g).max() < 1e-4 * max(1.0, np.abs(ref_g).max())
            fl = O.feature_loss(fr, fg).item()
            assert abs(fl - float(z[f"{name}_feature_loss"])) < 1e-4 * abs(float(z[f"{name}_feature_loss"]))
            dl, rl, gl = O.discriminator_loss(rs, gs)
            assert abs(dl.item() - float(z[f"{name}_disc_loss"])) < 1e-4 * abs(float(z[f"{name}_disc_loss"]))
            assert np.allclose(rl, z[f"{name}_disc_r_losses"], rtol=1e-4)
            assert np.allclose(gl, z[f"{name}_disc_g_losses"], rtol=1e-4)
            assert abs(O.generator_loss(gs)[0].item() - float(z[f"{name}_gen_loss"])) < 1e-4 * abs(
                float(z[f"{name}_gen_loss"]))


def test_crop_or_pad_rule():
    a = torch.arange(10.0).unsqueeze(0)
    assert O.crop_or_pad_segment(a, 4, 3).tolist() == [[3.0, 4.0, 5.0, 6.0]]
    assert O.crop_or_pad_segment(a, 12, 0).tolist() == [list(range(10)) + [0.0, 0.0]]


def test_training_step_oracle_vs_reference_golden():
    """oracle/train_oracle.py (autograd over the oracle's restated forward + torch AdamW) against the REFERENCE's
    own modules run through the same two steps (tests/golden/train_step_seed1234.npz): losses of both steps,
    dL/dy_g_hat, every parameter-gradient norm and the stored full gradients.  fp32 on both sides: 2e-3 relative
    (the two differ only in op ordering: weight-norm fold, reflect pad / view order, mel via rfft)."""
    import hifigan_b200 as H
    from oracle import hifigan_oracle as O
    from oracle import train_oracle as TO
    z = load_npz("train_step_seed1234.npz")
    h = H.AttrDict(O.config("v1"))
    torch.manual_seed(1234)
    G = H.Generator(h)
    mpd = H.MultiPeriodDiscriminator()
    msd = H.MultiScaleDiscriminator()
    sd_g, sd_p, sd_s = (TO.leaf_params(m.state_dict()) for m in (G, mpd, msd))
    ya = torch.from_numpy(z["audio"])
    x = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    optims = TO.make_optimizers(sd_g, sd_p, sd_s, h)
    for step in (1, 2):
        losses, gg, gp, gs, y_g, dy_g = TO.train_step(sd_g, sd_p, sd_s, h, x, ya.unsqueeze(1), y_mel, optims=optims)
        for k in ("loss_disc_f", "loss_disc_s", "loss_mel", "loss_fm_f", "loss_fm_s", "loss_gen_f", "loss_gen_s"):
            ref = float(z[f"step{step}_{k}"])
            assert abs(losses[k] - ref) <= 2e-3 * abs(ref), (step, k, losses[k], ref)
        if step == 1:
            ref_dy = torch.from_numpy(z["dy_g_hat"])
            assert (dy_g - ref_dy).norm() <= 2e-3 * ref_dy.norm()
            for name, grads in (("g", gg), ("mpd", gp), ("msd", gs)):
                keys = [str(k) for k in z[f"{name}_keys"]]
                norms = z[f"{name}_grad_norm"]
                for k, n in zip(keys, norms):
                    assert abs(float(grads[k].norm()) - n) <= 2e-3 * n + 1e-7, (name, k, float(grads[k].norm()), n)
            for key in z.files:
                if "_grad::" in key:
                    name, k = key.split("_grad::")
                    got = {"g": gg, "mpd": gp, "msd": gs}[name][k]
                    ref = torch.from_numpy(z[key])
                    assert (got - ref).norm() <= 2e-3 * ref.norm() + 1e-7, key
