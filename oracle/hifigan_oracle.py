"""CPU oracle for the HiFi-GAN vocoder hot path — TEST INFRASTRUCTURE, not product code.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this module (and then only as the checker / the timed CPU baseline).  The product path in
`hifi-gan_b200/` never routes through it and fails loudly when the CUDA library is missing.

What it restates
----------------
The reference (AlonKellner/hifi-gan, read-only at /root/reference) is pure Python whose arithmetic lives
in third-party libraries that are NOT vendored in the reference tree and are unpinned in its
requirements.txt:1-2:

  * torch       (Conv1d / ConvTranspose1d / Conv2d / AvgPool1d / weight_norm / spectral_norm / leaky_relu /
                 tanh / stft) — call sites src/models.py:4-5,16-31,56-59,81,86-88,96,134-140,196-204,228-229
  * torchaudio  (transforms.MelSpectrogram) — call site src/meldataset.py:59-71,81

This module restates the reference's *composition* of those ops as plain functions over a state_dict
(the reference's own key names) using `torch.nn.functional` on CPU in fp32 or fp64, and restates the
torchaudio front-end (periodic Hann, frame, rFFT, |X|^2, HTK filterbank, log-clamp) from its published
algorithm.  `oracle/conv_ref.c` additionally restates the conv primitives as direct loops in C so the
torch ops themselves are pinned independently (tests/test_oracle_cpu.py).

Pinning: the reference ships no golden vectors or tests for this path, so the oracle
is pinned against outputs of the reference itself, generated in the build container by
`tests/golden/make_golden.py` (which imports /root/reference/src with matplotlib/librosa stubbed) and
committed under tests/golden/.  tests/test_oracle_cpu.py checks this oracle against every such fixture.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # src/models.py:8

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------------
# configs (UPSTREAM jik876/hifi-gan config_v{1,2,3}.json values; the fork deleted the files but its
# Generator still consumes these attributes, src/models.py:79-96)
# --------------------------------------------------------------------------------------------------
class Cfg(dict):
    """Attribute-access dict, the role src/env.py:5-8 AttrDict plays for the reference."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


_COMMON = dict(segment_size=8192, num_mels=80, n_fft=1024, hop_size=256, win_size=1024,
               sampling_rate=22050, fmin=0, fmax=8000, fmax_for_loss=None, batch_size=16,
               learning_rate=2e-4, adam_b1=0.8, adam_b2=0.99, lr_decay=0.999, seed=1234)


def config(version: str = "v1") -> Cfg:
    if version == "v1":
        spec = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                    upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])
    elif version == "v2":
        spec = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                    upsample_initial_channel=128, resblock_kernel_sizes=[3, 7, 11],
                    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])
    elif version == "v3":
        spec = dict(resblock="2", upsample_rates=[8, 8, 4], upsample_kernel_sizes=[16, 16, 8],
                    upsample_initial_channel=256, resblock_kernel_sizes=[3, 5, 7],
                    resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]])
    elif version == "tiny":  # small same-topology config for fast CPU tests / golden fixtures
        spec = dict(resblock="1", upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4],
                    upsample_initial_channel=64, resblock_kernel_sizes=[3, 5],
                    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]])
    elif version == "tiny2":
        spec = dict(resblock="2", upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4],
                    upsample_initial_channel=64, resblock_kernel_sizes=[3, 5],
                    resblock_dilation_sizes=[[1, 2], [2, 6]])
    else:
        raise ValueError(f"unknown config {version!r}")
    c = Cfg(_COMMON)
    c.update(spec)
    return c


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """src/utils.py:78-79."""
    return int((kernel_size * dilation - dilation) / 2)


# --------------------------------------------------------------------------------------------------
# weight re-parametrisations (torch.nn.utils.weight_norm / spectral_norm, un-vendored)
# --------------------------------------------------------------------------------------------------
def fold_weight_norm(g: Tensor, v: Tensor) -> Tensor:
    """w = g * v / ||v||_2, norm over every dim but 0 (old-style weight_norm, dim=0).
    For ConvTranspose1d dim 0 is the INPUT channel."""
    dims = tuple(range(1, v.dim()))
    return v * (g / v.pow(2).sum(dim=dims, keepdim=True).sqrt())


def spectral_norm_step(w: Tensor, u: Tensor, v: Tensor, n_iter: int = 1, eps: float = 1e-12
                       ) -> Tuple[Tensor, Tensor, Tensor]:
    """One train-mode forward of torch.nn.utils.spectral_norm (dim=0, n_power_iterations=1, eps=1e-12):
    v <- normalize(W^T u); u <- normalize(W v); sigma = u . (W v).  Returns (W / sigma, u, v)."""
    wm = w.flatten(1)
    with torch.no_grad():   # torch runs the power iteration outside autograd: u, v are constants of sigma
        for _ in range(n_iter):
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=eps)
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma, u, v


def weight_of(sd: StateDict, prefix: str) -> Tensor:
    """Effective conv weight for `prefix` from a reference state_dict, whichever form it is in:
    weight_g/weight_v (weight_norm attached), weight (after remove_weight_norm,
    src/models.py:118-125) or weight_orig/weight_u (spectral_norm, eval-mode fold: W / sigma with the
    stored u, v — src/models.py:194)."""
    if prefix + ".weight_g" in sd:
        return fold_weight_norm(sd[prefix + ".weight_g"], sd[prefix + ".weight_v"])
    if prefix + ".weight_orig" in sd:
        w = sd[prefix + ".weight_orig"]
        u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
        sigma = torch.dot(u, torch.mv(w.flatten(1), v))
        return w / sigma
    return sd[prefix + ".weight"]


# --------------------------------------------------------------------------------------------------
# Generator (src/models.py:75-125)
# --------------------------------------------------------------------------------------------------
def resblock1_forward(sd: StateDict, prefix: str, x: Tensor, kernel_size: int,
                      dilation: Sequence[int]) -> Tensor:
    """src/models.py:35-42."""
    for i, d in enumerate(dilation):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, weight_of(sd, f"{prefix}.convs1.{i}"), sd[f"{prefix}.convs1.{i}.bias"],
                      dilation=d, padding=get_padding(kernel_size, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, weight_of(sd, f"{prefix}.convs2.{i}"), sd[f"{prefix}.convs2.{i}.bias"],
                      dilation=1, padding=get_padding(kernel_size, 1))
        x = xt + x
    return x


def resblock2_forward(sd: StateDict, prefix: str, x: Tensor, kernel_size: int,
                      dilation: Sequence[int]) -> Tensor:
    """src/models.py:63-68."""
    for i, d in enumerate(dilation):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, weight_of(sd, f"{prefix}.convs.{i}"), sd[f"{prefix}.convs.{i}.bias"],
                      dilation=d, padding=get_padding(kernel_size, d))
        x = xt + x
    return x


def generator_forward(sd: StateDict, h, x: Tensor, taps: Optional[dict] = None) -> Tensor:
    """src/models.py:100-116.  x [B,80,F] -> [B,1,F*prod(upsample_rates)].
    `taps`, when given, receives intermediate activations (for per-layer kernel tests)."""
    nk = len(h.resblock_kernel_sizes)
    x = F.conv1d(x, weight_of(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    if taps is not None:
        taps["conv_pre"] = x
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, weight_of(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        if taps is not None:
            taps[f"ups.{i}"] = x
        xs = None
        for j, (rk, rd) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            pfx = f"resblocks.{i * nk + j}"
            if h.resblock == "1":  # string compare, src/models.py:82
                r = resblock1_forward(sd, pfx, x, rk, rd)
            else:
                r = resblock2_forward(sd, pfx, x, rk, rd)
            xs = r if xs is None else xs + r
        x = xs / nk
        if taps is not None:
            taps[f"mrf.{i}"] = x
    x = F.leaky_relu(x)  # default slope 0.01, src/models.py:112
    x = F.conv1d(x, weight_of(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


# --------------------------------------------------------------------------------------------------
# Discriminators (src/models.py:128-248) and losses (:251-282) — eval-mode forward (no power iteration)
# --------------------------------------------------------------------------------------------------
_DP_LAYERS = [(1, 32, 3), (32, 128, 3), (128, 512, 3), (512, 1024, 3), (1024, 1024, 1)]
_DS_LAYERS = [(1, 128, 15, 1, 1, 7), (128, 128, 41, 2, 4, 20), (128, 256, 41, 2, 16, 20),
              (256, 512, 41, 4, 16, 20), (512, 1024, 41, 4, 16, 20), (1024, 1024, 41, 1, 16, 20),
              (1024, 1024, 5, 1, 1, 2)]


def discriminator_p_forward(sd: StateDict, prefix: str, x: Tensor, period: int
                            ) -> Tuple[Tensor, List[Tensor]]:
    """src/models.py:142-161."""
    fmap = []
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), "reflect")
        t = t + n_pad
    x = x.view(b, c, t // period, period)
    for l, (_, _, s) in enumerate(_DP_LAYERS):
        x = F.conv2d(x, weight_of(sd, f"{prefix}.convs.{l}"), sd[f"{prefix}.convs.{l}.bias"],
                     stride=(s, 1), padding=(2, 0))
        x = F.leaky_relu(x, LRELU_SLOPE)
        fmap.append(x)
    x = F.conv2d(x, weight_of(sd, f"{prefix}.conv_post"), sd[f"{prefix}.conv_post.bias"],
                 padding=(1, 0))
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def _sn_weight(sd: StateDict, prefix: str, train: bool) -> Tensor:
    """Spectral-norm layers in train mode run one power iteration per forward and update the u/v buffers
    in place; `sd` is mutated accordingly."""
    if train and prefix + ".weight_orig" in sd:
        w, u, v = spectral_norm_step(sd[prefix + ".weight_orig"], sd[prefix + ".weight_u"],
                                     sd[prefix + ".weight_v"])
        sd[prefix + ".weight_u"], sd[prefix + ".weight_v"] = u, v
        return w
    return weight_of(sd, prefix)


def discriminator_s_forward(sd: StateDict, prefix: str, x: Tensor, train: bool = False
                            ) -> Tuple[Tensor, List[Tensor]]:
    """src/models.py:206-216."""
    fmap = []
    for l, (_, _, k, s, g, p) in enumerate(_DS_LAYERS):
        x = F.conv1d(x, _sn_weight(sd, f"{prefix}.convs.{l}", train), sd[f"{prefix}.convs.{l}.bias"],
                     stride=s, padding=p, groups=g)
        x = F.leaky_relu(x, LRELU_SLOPE)
        fmap.append(x)
    x = F.conv1d(x, _sn_weight(sd, f"{prefix}.conv_post", train), sd[f"{prefix}.conv_post.bias"], padding=1)
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


MPD_PERIODS = (2, 3, 5, 7, 11)  # src/models.py:167-173


def mpd_forward(sd: StateDict, y: Tensor, y_hat: Tensor):
    """src/models.py:175-188."""
    y_d_rs, y_d_gs, fmap_rs, fmap_gs = [], [], [], []
    for i, p in enumerate(MPD_PERIODS):
        r, fr = discriminator_p_forward(sd, f"discriminators.{i}", y, p)
        g, fg = discriminator_p_forward(sd, f"discriminators.{i}", y_hat, p)
        y_d_rs.append(r); fmap_rs.append(fr); y_d_gs.append(g); fmap_gs.append(fg)
    return y_d_rs, y_d_gs, fmap_rs, fmap_gs


def msd_forward(sd: StateDict, y: Tensor, y_hat: Tensor, train: bool = False):
    """src/models.py:232-248 (AvgPool1d(4,2,padding=2) cumulatively between scales).  With train=True the
    spectral-norm scale (discriminators.0) runs its power iterations call by call and `sd` is updated."""
    y_d_rs, y_d_gs, fmap_rs, fmap_gs = [], [], [], []
    for i in range(3):
        if i != 0:
            y = F.avg_pool1d(y, 4, 2, padding=2)
            y_hat = F.avg_pool1d(y_hat, 4, 2, padding=2)
        r, fr = discriminator_s_forward(sd, f"discriminators.{i}", y, train)
        g, fg = discriminator_s_forward(sd, f"discriminators.{i}", y_hat, train)
        y_d_rs.append(r); fmap_rs.append(fr); y_d_gs.append(g); fmap_gs.append(fg)
    return y_d_rs, y_d_gs, fmap_rs, fmap_gs


def feature_loss(fmap_r, fmap_g) -> Tensor:
    """src/models.py:251-257."""
    loss = 0
    for dr, dg in zip(fmap_r, fmap_g):
        for rl, gl in zip(dr, dg):
            loss = loss + torch.mean(torch.abs(rl - gl))
    return loss * 2


def discriminator_loss(real_outs, gen_outs):
    """src/models.py:260-271."""
    loss, r_losses, g_losses = 0, [], []
    for dr, dg in zip(real_outs, gen_outs):
        r_loss = torch.mean((1 - dr) ** 2)
        g_loss = torch.mean(dg ** 2)
        loss = loss + (r_loss + g_loss)
        r_losses.append(r_loss.item()); g_losses.append(g_loss.item())
    return loss, r_losses, g_losses


def generator_loss(outs):
    """src/models.py:274-282."""
    loss, gen_losses = 0, []
    for dg in outs:
        l = torch.mean((1 - dg) ** 2)
        gen_losses.append(l)
        loss = loss + l
    return loss, gen_losses


# --------------------------------------------------------------------------------------------------
# mel_spectrogram (src/meldataset.py:56-85; torchaudio.transforms.MelSpectrogram restated)
# --------------------------------------------------------------------------------------------------
def hann_periodic(win: int, dtype=torch.float64) -> Tensor:
    n = torch.arange(win, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / win)).to(dtype)


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                        dtype=torch.float64) -> Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') — [n_freqs, n_mels]."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=torch.float64)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0).to(dtype)


def mel_spectrogram(y: Tensor, n_fft: int, num_mels: int, sampling_rate: int, hop_size: int,
                    win_size: int, fmin: float, fmax: Optional[float], center: bool = False) -> Tensor:
    """y [B,T] -> log-power-mel [B,num_mels,frames]; computes in y.dtype (use fp64 for a tight oracle).
    reflect pad (n_fft-hop)/2 each side (meldataset.py:78), frames of n_fft at hop (center=False),
    periodic Hann(win) zero-padded to n_fft, rFFT, re^2+im^2 (power=2.0), HTK fbank with
    f_max = sr//2 when None, log(clamp(., 1e-5)) (meldataset.py:32-33)."""
    assert not center, "the reference always calls with center=False"
    pad = int((n_fft - hop_size) / 2)
    y = F.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    win = torch.zeros(n_fft, dtype=y.dtype)
    left = (n_fft - win_size) // 2
    win[left:left + win_size] = hann_periodic(win_size, y.dtype)
    frames = y.unfold(-1, n_fft, hop_size) * win            # [B, F, n_fft]
    spec = torch.fft.rfft(frames, dim=-1)
    power = spec.real ** 2 + spec.imag ** 2                  # [B, F, n_fft/2+1]
    f_max = float(sampling_rate // 2) if fmax is None else float(fmax)
    fb = melscale_fbanks_htk(n_fft // 2 + 1, float(fmin), f_max, num_mels, sampling_rate, y.dtype)
    mel = torch.matmul(power, fb).transpose(1, 2)            # [B, num_mels, F]
    return torch.log(torch.clamp(mel, min=1e-5))


# --------------------------------------------------------------------------------------------------
# MelDataset crop / pad rule (src/meldataset.py:141-150), used to pin the batched GPU sampler
# --------------------------------------------------------------------------------------------------
def crop_or_pad_segment(audio: Tensor, segment_length: int, audio_start: int) -> Tensor:
    """audio [1,L].  L >= seg: audio[:, start:start+seg] (start drawn by random.randint(0, L-seg),
    inclusive); else right zero-pad to seg."""
    if audio.size(1) >= segment_length:
        return audio[:, audio_start:audio_start + segment_length]
    return F.pad(audio, (0, segment_length - audio.size(1)), "constant")


def finetune_crop_or_pad(mel: Tensor, audio: Tensor, segment_length: int, hop_size: int, mel_start: int):
    """Fine-tuning branch of MelDataset.__getitem__ (src/meldataset.py:155-172): mel [1,M,F] loaded from .npy,
    audio [1,L].  L >= seg: frames [mel_start, mel_start + fps) with mel_start drawn by
    random.randint(0, F - fps - 1) and the audio cropped at mel_start * hop; else both right zero-padded."""
    fps = math.ceil(segment_length / hop_size)
    if audio.size(1) >= segment_length:
        return (mel[:, :, mel_start:mel_start + fps],
                audio[:, mel_start * hop_size:(mel_start + fps) * hop_size])
    return (F.pad(mel, (0, fps - mel.size(2)), "constant"),
            F.pad(audio, (0, segment_length - audio.size(1)), "constant"))


# --------------------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench
# --------------------------------------------------------------------------------------------------
def synthetic_audio(batch: int, t: int, seed: int = 0, sr: int = 22050) -> Tensor:
    """a = 0.5*sin(2*pi*f0*t) + 0.1*N(0,1), clipped to +-0.95, f0 ~ U(80,400) per item."""
    g = torch.Generator().manual_seed(seed)
    f0 = 80.0 + 320.0 * torch.rand(batch, 1, generator=g)
    n = torch.arange(t, dtype=torch.float32).unsqueeze(0)
    a = 0.5 * torch.sin(2.0 * math.pi * f0 * n / sr) + 0.1 * torch.randn(batch, t, generator=g)
    return a.clamp_(-0.95, 0.95)
