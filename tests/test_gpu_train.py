"""Backward-pass kernels against torch autograd in fp32 on the same bf16-rounded operands.

Tolerances: the tensor-core kernels multiply exact bf16 operands and accumulate in fp32, so weight gradients (fp32
outputs) agree to accumulation-order noise (rel 2e-3 of the tensor's max); data gradients are stored in bf16
(|err| <= 2^-8 |ref| + eps).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import hifigan_b200
    hifigan_b200._lib.lib()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return hifigan_b200


def _st():
    return torch.cuda.current_stream().cuda_stream


WGRAD_CASES = [
    # b, t_in, cin, cout, k, stride, dil, pad, groups
    (2, 300, 256, 256, 11, 1, 5, 25, 1),     # wide, N = 256
    (3, 257, 128, 128, 7, 1, 3, 9, 1),       # wide, N = 128
    (2, 1000, 64, 64, 11, 1, 5, 25, 1),      # narrow SW128: 2 taps per M tile
    (2, 1000, 64, 64, 3, 1, 1, 1, 1),
    (2, 2000, 32, 32, 7, 1, 3, 9, 1),        # narrow SW64: 4 taps per M tile
    (2, 500, 32, 32, 11, 1, 1, 5, 1),
    (1, 64, 512, 2048, 3, 1, 1, 1, 1),       # polyphase ups0 shape (cout = 8 * 256)
    (2, 90, 128, 512, 7, 1, 1, 3, 1),        # conv_pre shape (cin 80 padded to 128)
    (3, 300, 32, 128, 5, 3, 1, 2, 1),        # DiscriminatorP strided layers
    (2, 911, 128, 512, 5, 3, 1, 2, 1),
    (2, 100, 512, 1024, 5, 3, 1, 2, 1),
    (4, 51, 1024, 1024, 5, 1, 1, 2, 1),
    (2, 1000, 128, 128, 41, 2, 1, 20, 4),    # DiscriminatorS grouped / strided layers
    (2, 600, 128, 256, 41, 2, 1, 20, 16),
    (1, 700, 256, 512, 41, 4, 1, 20, 16),
    (1, 515, 512, 1024, 41, 4, 1, 20, 16),
    (1, 130, 1024, 1024, 41, 1, 1, 20, 16),
    (1, 1, 64, 64, 3, 1, 1, 1, 1),           # single time step
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad_vs_torch_autograd(H, case):
    from hifigan_b200 import _lib
    from hifigan_b200.models import _DiscLayer, _round_up
    L = _lib.lib()
    b, t, cin, cout, k, s, d, pad, g = case
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(sum(case))
    rows = _round_up(t, s)
    x = torch.zeros(b, rows, cin, dtype=torch.bfloat16, device=dev)
    x[:, :t] = torch.randn(b, t, cin, generator=gen).to(dev).bfloat16()
    t_out = (t + 2 * pad - d * (k - 1) - 1) // s + 1
    rows_out = t_out + 3
    dy = torch.randn(b, rows_out, cout, generator=gen).to(dev).bfloat16()   # pitch rows hold garbage on purpose
    layer = _DiscLayer(cin, cout, k, s, pad, g)
    dwp = torch.full((k, cout, layer.cin_tile), 3.0, dtype=torch.float32, device=dev)
    _lib.check(L.hg_conv1d_wgrad(x.data_ptr(), dy.data_ptr(), b, rows, cin, t_out, rows_out, layer.groups_eff, cout,
                                 k, s, d, pad, dwp.data_ptr(), 0, _st()), "hg_conv1d_wgrad")
    # accumulate = 1 adds a second copy
    _lib.check(L.hg_conv1d_wgrad(x.data_ptr(), dy.data_ptr(), b, rows, cin, t_out, rows_out, layer.groups_eff, cout,
                                 k, s, d, pad, dwp.data_ptr(), 1, _st()), "hg_conv1d_wgrad")
    dw = torch.empty(cout, cin // g, k, dtype=torch.float32, device=dev)
    from ctypes import c_int
    order = (c_int * k)(*layer.order)
    _lib.check(L.hg_unpack_wgrad_conv(dwp.data_ptr(), cout, cin // g, k, cout, layer.cin_tile, cout // g, layer.merge,
                                      order, dw.data_ptr(), _st()), "hg_unpack_wgrad_conv")
    torch.cuda.synchronize()
    w = torch.zeros(cout, cin // g, k, device=dev, requires_grad=True)
    y = F.conv1d(x[:, :t].float().transpose(1, 2), w, None, stride=s, padding=pad, dilation=d, groups=g)
    assert y.shape[2] == t_out
    y.backward(dy[:, :t_out].float().transpose(1, 2))
    ref = 2 * w.grad
    err = (dw - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-5, (err, ref.abs().max().item())


DGRAD_CASES = [
    # b, t, cin, cout, k, dil  (stride-1 "same" convs of the Generator)
    (2, 300, 256, 256, 11, 5), (3, 257, 128, 128, 7, 3), (2, 1000, 64, 64, 3, 1), (2, 777, 32, 32, 7, 5),
    (2, 129, 128, 512, 7, 1), (8, 4224, 128, 128, 7, 3),
]


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_dgrad_stride1_vs_torch_autograd(H, case):
    """dx = (conv^T(dy) * lrelu'(x_in)) + res, with the mask taken from the stored activated input."""
    from hifigan_b200 import _lib
    L = _lib.lib()
    b, t, cin, cout, k, d = case
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(sum(case))
    pad = (k - 1) * d // 2
    w = (torch.randn(cout, cin, k, generator=gen) / (cin * k) ** 0.5).to(dev)
    wp = torch.empty(k, cout, cin, dtype=torch.bfloat16, device=dev)
    _lib.check(L.hg_pack_conv1d_weight(w.data_ptr(), 0, cout, cin, k, cin, wp.data_ptr(), _st()))
    wd = torch.empty(k, cin, cout, dtype=torch.bfloat16, device=dev)
    _lib.check(L.hg_pack_dgrad_weight(wp.data_ptr(), k, cout, cin, wd.data_ptr(), _st()))
    x_pre = torch.randn(b, t, cin, generator=gen).to(dev)
    x_act = F.leaky_relu(x_pre, 0.1).bfloat16()                 # what the forward stored
    dy = torch.randn(b, t + 2, cout, generator=gen).to(dev).bfloat16()
    res = torch.randn(b, t, cin, generator=gen).to(dev).bfloat16()
    out = torch.empty(b, t, cin, dtype=torch.bfloat16, device=dev)
    pre = torch.randn(b, t, cin, generator=gen).to(dev).bfloat16()
    bsum = torch.full((2, cin), 0.25, device=dev)       # two destinations, both ADDED to
    _lib.check(L.hg_conv1d_dgrad(dy.data_ptr(), wd.data_ptr(), b, t, t + 2, cout, t, t, 1, 0, cin, k, d,
                                 (k - 1) * d - pad, x_act.data_ptr(), 0.1, 0, 0, 0.0, res.data_ptr(), 0, 0, 0.5,
                                 out.data_ptr(), 0, 0, 1, 0, pre.data_ptr(), bsum[0].data_ptr(), bsum[1].data_ptr(), 0,
                                 0, _st()), "hg_conv1d_dgrad")
    torch.cuda.synchronize()
    wr = wp.float().permute(1, 2, 0).contiguous()
    xin = x_act.float().transpose(1, 2).requires_grad_(True)
    y = F.conv1d(xin, wr, None, dilation=d, padding=pad)
    y.backward(dy[:, :t].float().transpose(1, 2))
    mask = torch.where(x_act.float() > 0, 1.0, 0.1)
    ref = ((xin.grad.transpose(1, 2) + pre.float()) * mask + res.float()) * 0.5
    assert bool(((out.float() - ref).abs() <= 2.0 ** -7 * ref.abs() + 2e-3).all())
    # the fused bias gradient: fp32 column sums of the un-rounded output, added to every destination
    cs = ref.sum((0, 1))
    tol = 1e-3 * ref.abs().sum((0, 1)).max().item() + 1e-3
    assert (bsum[0] - 0.25 - cs).abs().max().item() <= tol and (bsum[0] - bsum[1]).abs().max().item() <= tol


STRIDED_DGRAD_CASES = [
    # b, t_in, cin, cout, k, stride, pad, groups
    (3, 300, 32, 128, 5, 3, 2, 1),
    (2, 911, 128, 512, 5, 3, 2, 1),
    (2, 100, 512, 1024, 5, 3, 2, 1),
    (2, 51, 1024, 1024, 5, 1, 2, 1),
    (2, 1000, 128, 128, 41, 2, 20, 4),
    (2, 600, 128, 256, 41, 2, 20, 16),
    (1, 700, 256, 512, 41, 4, 20, 16),
    (1, 515, 512, 1024, 41, 4, 20, 16),
    (1, 130, 1024, 1024, 41, 1, 20, 16),
]


@pytest.mark.parametrize("case", STRIDED_DGRAD_CASES)
def test_dgrad_strided_grouped_vs_torch_autograd(H, case):
    """Discriminator layers: polyphase data gradient with the feature-matching term and the leaky_relu mask."""
    from hifigan_b200 import _lib
    from hifigan_b200.models import _DiscLayer, _round_up
    from hifigan_b200.train import _DiscBwdLayer
    L = _lib.lib()
    b, t, cin, cout, k, s, pad, g = case
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(sum(case))
    rows = _round_up(t, s)
    layer = _DiscLayer(cin, cout, k, s, pad, g)
    w = (torch.randn(cout, cin // g, k, generator=gen) / (cin // g * k) ** 0.5).to(dev).bfloat16().float()
    t_out = (t + 2 * pad - k) // s + 1
    rows_out = t_out + 2
    dy = torch.randn(b, rows_out, cout, generator=gen).to(dev).bfloat16()
    act_g = torch.zeros(b, rows, cin, dtype=torch.bfloat16, device=dev)
    act_g[:, :t] = F.leaky_relu(torch.randn(b, t, cin, generator=gen), 0.1).to(dev).bfloat16()
    act_r = torch.zeros_like(act_g)
    act_r[:, :t] = F.leaky_relu(torch.randn(b, t, cin, generator=gen), 0.1).to(dev).bfloat16()
    bl = _DiscBwdLayer(layer, dev)
    bl.pack(w, layer.pack(w))
    out = torch.zeros(b, rows, cin, dtype=torch.bfloat16, device=dev)
    fm_coef = 0.37
    bsum = torch.zeros(cin, device=dev)
    bl.dgrad(L, dy, b, t_out, rows_out, rows, act_g, act_r, fm_coef, out, _st(), bias_dst=bsum.data_ptr())
    torch.cuda.synchronize()
    xin = act_g[:, :t].float().transpose(1, 2).requires_grad_(True)
    y = F.conv1d(xin, w, None, stride=s, padding=pad, groups=g)
    y.backward(dy[:, :t_out].float().transpose(1, 2))
    ag, ar = act_g[:, :t].float(), act_r[:, :t].float()
    ref = (xin.grad.transpose(1, 2) + fm_coef * torch.sign(ag - ar)) * torch.where(ag > 0, 1.0, 0.1)
    got = out[:, :t].float()
    assert bool(((got - ref).abs() <= 2.0 ** -7 * ref.abs() + 3e-3).all()), (got - ref).abs().max().item()
    if rows == t:       # fused bias gradient: the `stride` phases of a channel fold together (bias_mod = cin)
        cs = ref.sum((0, 1))
        assert (bsum - cs).abs().max().item() <= 1e-3 * ref.abs().sum((0, 1)).max().item() + 1e-3


def test_small_backward_kernels_vs_torch(H):
    from hifigan_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(5)
    # column sums (bias gradients)
    x = torch.randn(3, 50, 64, generator=gen).to(dev).bfloat16()
    out = torch.empty(64, device=dev)
    _lib.check(L.hg_colsum_bf16(x.data_ptr(), 3, 47, 50, 64, 0, out.data_ptr(), _st()))
    assert torch.allclose(out, x[:, :47].float().sum((0, 1)), atol=1e-3)
    # conv_post + tanh backward
    b, t, c, k = 2, 700, 32, 7
    xr = torch.randn(b, t, c, generator=gen).to(dev)
    xa = F.leaky_relu(xr, 0.01).bfloat16()
    w = (torch.randn(1, c, k, generator=gen) * 0.1).to(dev)
    bias = torch.randn(1, generator=gen).to(dev)
    dy = torch.randn(b, t, generator=gen).to(dev)
    xin = xa.float().transpose(1, 2).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    y = torch.tanh(F.conv1d(xin, wr, br, padding=3)).squeeze(1)
    y.backward(dy)
    dx = torch.empty(b, t, c, dtype=torch.bfloat16, device=dev)
    dpre = torch.empty(b, t, device=dev)
    dw = torch.zeros(c, k, device=dev)
    db = torch.zeros(1, device=dev)
    bs = torch.zeros(3, c, device=dev)
    _lib.check(L.hg_conv_post_tanh_bwd(xa.data_ptr(), w[0].contiguous().data_ptr(), y.detach().contiguous().data_ptr(),
                                       dy.data_ptr(), b, t, c, k, 0.01, 1.0 / 3, dx.data_ptr(), dpre.data_ptr(),
                                       dw.data_ptr(), db.data_ptr(), bs[0].data_ptr(), bs[1].data_ptr(),
                                       bs[2].data_ptr(), _st()))
    torch.cuda.synchronize()
    ref_dx = xin.grad.transpose(1, 2) * torch.where(xa.float() > 0, 1.0, 0.01) / 3
    assert bool(((dx.float() - ref_dx).abs() <= 2.0 ** -7 * ref_dx.abs() + 1e-4).all())
    assert torch.allclose(bs[0], ref_dx.sum((0, 1)), rtol=1e-4, atol=1e-4) and torch.allclose(bs[0], bs[2], rtol=1e-5, atol=1e-5)
    assert torch.allclose(dw, wr.grad[0], rtol=1e-3, atol=1e-3)
    assert torch.allclose(db, br.grad, rtol=1e-3, atol=1e-3)
    # avg-pool backward
    xin = torch.randn(3, 4097, generator=gen).to(dev).requires_grad_(True)
    yo = F.avg_pool1d(xin.unsqueeze(1), 4, 2, 2).squeeze(1)
    do = torch.randn(yo.shape, generator=gen).to(dev)
    yo.backward(do)
    din = torch.zeros(3, 4097, device=dev)
    _lib.check(L.hg_avgpool_4_2_2_bwd(do.data_ptr(), 3, 4097, din.data_ptr(), _st()))
    assert torch.allclose(din, xin.grad, atol=1e-6)
    # weight-norm backward
    v = torch.randn(40, 30, 7, generator=gen).to(dev)
    g = torch.rand(40, 1, 1, generator=gen).to(dev) + 0.5
    vv, gg = v.clone().requires_grad_(True), g.clone().requires_grad_(True)
    wn = torch._weight_norm(vv, gg, 0)
    dwn = torch.randn(40, 30, 7, generator=gen).to(dev)
    wn.backward(dwn)
    dv, dg = torch.empty_like(v), torch.empty(40, device=dev)
    _lib.check(L.hg_weight_norm_bwd(dwn.data_ptr(), v.data_ptr(), g.data_ptr(), 40, 210, 0, dv.data_ptr(),
                                    dg.data_ptr(), _st()))
    assert torch.allclose(dv, vv.grad, rtol=1e-4, atol=1e-5) and torch.allclose(dg, gg.grad.flatten(), rtol=1e-4, atol=1e-5)
    # AdamW, three steps
    p = torch.randn(1000, generator=gen).to(dev)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=2e-4, betas=(0.8, 0.99))
    m, vv2 = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(1000, generator=gen).to(dev)
        ref_p.grad = gr.clone()
        opt.step()
        _lib.check(L.hg_adamw_step(p.data_ptr(), gr.data_ptr(), m.data_ptr(), vv2.data_ptr(), 1000, 2e-4, 0.8, 0.99,
                                   1e-8, 0.01, step, 0, 0, 1.0, _st()))
    assert torch.allclose(p, ref_p.detach(), rtol=1e-5, atol=1e-6)


def test_disc_end_backward_kernels_vs_torch(H):
    from hifigan_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(9)
    # last conv (Cout = 1): data + weight gradient with the feature-matching term
    s_, h, rows, c, k = 6, 51, 52, 1024, 3
    x = torch.zeros(s_, rows, c, dtype=torch.bfloat16, device=dev)
    x[:, :h] = F.leaky_relu(torch.randn(s_, h, c, generator=gen), 0.1).to(dev).bfloat16()
    fr = torch.zeros_like(x)
    fr[:, :h] = F.leaky_relu(torch.randn(s_, h, c, generator=gen), 0.1).to(dev).bfloat16()
    w = (torch.randn(1, c, k, generator=gen) * 0.05).to(dev)
    dl = torch.randn(s_, h, generator=gen).to(dev)
    xin = x[:, :h].float().transpose(1, 2).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = torch.zeros(1, device=dev, requires_grad=True)
    F.conv1d(xin, wr, br, padding=1).squeeze(1).backward(dl)
    dx = torch.zeros_like(x)
    dw, db = torch.zeros(c, k, device=dev), torch.zeros(1, device=dev)
    pre = torch.zeros_like(x)
    pre[:, :h] = torch.randn(s_, h, c, generator=gen).to(dev).bfloat16()
    bsum = torch.zeros(c, device=dev)
    _lib.check(L.hg_disc_last_conv_bwd(x.data_ptr(), w[0].contiguous().data_ptr(), dl.data_ptr(), s_, h, rows, c, k,
                                       0.1, fr.data_ptr(), 0.25, pre.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                       db.data_ptr(), bsum.data_ptr(), _st()))
    torch.cuda.synchronize()
    xa, ra = x[:, :h].float(), fr[:, :h].float()
    ref = (xin.grad.transpose(1, 2) + 0.25 * torch.sign(xa - ra) + pre[:, :h].float()) * torch.where(xa > 0, 1.0, 0.1)
    assert bool(((dx[:, :h].float() - ref).abs() <= 2.0 ** -7 * ref.abs() + 1e-4).all())
    assert torch.equal(dx[:, h:], torch.zeros_like(dx[:, h:]))          # pitch rows stay untouched
    assert torch.allclose(bsum, ref.sum((0, 1)), rtol=1e-4, atol=1e-3)
    assert torch.allclose(dw, wr.grad[0], rtol=1e-3, atol=1e-3) and torch.allclose(db, br.grad, rtol=1e-3, atol=1e-3)
    # first conv (Cin = 1) with the period view and reflect pad
    for period, k0, s0, p0, c0, t in [(3, 5, 3, 2, 32, 1000), (1, 15, 1, 7, 128, 777), (7, 5, 3, 2, 32, 8192)]:
        b = 2
        y = torch.randn(b, t, generator=gen).to(dev)
        w0 = torch.randn(c0, k0, generator=gen).to(dev)
        yy = y.clone().requires_grad_(True)
        ww = w0.clone().requires_grad_(True)
        bb = torch.zeros(c0, device=dev, requires_grad=True)
        sig = yy.unsqueeze(1)
        if t % period:
            sig = F.pad(sig, (0, period - t % period), "reflect")
        hh = sig.shape[-1] // period
        o = F.conv2d(sig.view(b, 1, hh, period), ww.view(c0, 1, k0, 1), bb, stride=(s0, 1), padding=(p0, 0))
        h_out = o.shape[2]
        dpre = torch.randn(b * period, h_out + 1, c0, generator=gen).to(dev).bfloat16()
        o.backward(dpre[:, :h_out].float().view(b, period, h_out, c0).permute(0, 3, 2, 1))
        dw0, db0, dyy = torch.zeros(c0, k0, device=dev), torch.zeros(c0, device=dev), torch.zeros(b, t, device=dev)
        _lib.check(L.hg_disc_first_conv_bwd(y.data_ptr(), w0.data_ptr(), dpre.data_ptr(), b, t, period, k0, s0, p0, c0,
                                            h_out + 1, dw0.data_ptr(), db0.data_ptr(), dyy.data_ptr(), _st()))
        torch.cuda.synchronize()
        assert torch.allclose(dw0, ww.grad, rtol=2e-3, atol=2e-2), (dw0 - ww.grad).abs().max()
        assert torch.allclose(db0, bb.grad, rtol=2e-3, atol=2e-2)
        assert torch.allclose(dyy, yy.grad, rtol=2e-3, atol=2e-3), (dyy - yy.grad).abs().max()


# ------------------------------------------------------------------------------------------ the full step
@pytest.mark.parametrize("fmax,t,b", [(None, 8192, 3), (8000, 5000, 2), (None, 12345, 1)])
def test_mel_backward_vs_autograd(H, fmax, t, b):
    """hg_mel_bwd (forward FFT recompute + one adjoint 512-point transform per frame) against torch autograd through
    the fp64 oracle mel: silent item (clamp -> zero gradient), reflect-padded edges, a length that is not a multiple
    of the hop.  fp32 kernel: max error 2e-3 of the largest gradient, cosine > 0.99999."""
    from oracle import hifigan_oracle as O
    from hifigan_b200 import _lib
    ya = O.synthetic_audio(b, t, seed=4)
    if b > 1:
        ya[1] = 0.0
    yd = ya.double().requires_grad_(True)
    mel = O.mel_spectrogram(yd, 1024, 80, 22050, 256, 1024, 0, fmax)
    dmel = torch.randn(mel.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (mel * dmel).sum().backward()
    yc = ya.cuda()
    H.mel_spectrogram(yc, 1024, 80, 22050, 256, 1024, 0, fmax)               # creates / caches the plan
    plan = H.meldataset.torch_mels[f"{yc.device}_1024_80_22050_256_1024_0_{fmax}_False"]
    dy = torch.zeros_like(yc)
    dm = dmel.float().cuda().contiguous()
    _lib.check(_lib.lib().hg_mel_bwd(plan.handle, yc.data_ptr(), dm.data_ptr(), b, t, dy.data_ptr(), _st()), "hg_mel_bwd")
    got, ref = dy.cpu().double(), yd.grad
    for i in range(b):
        if b > 1 and i == 1:
            assert torch.all(got[i] == 0)
            continue
        assert (got[i] - ref[i]).abs().max() / ref[i].abs().max() < 2e-3
        assert F.cosine_similarity(got[i], ref[i], dim=0).item() > 0.99999


def _seeded_step(H):
    from oracle import hifigan_oracle as O
    from hifigan_b200.train import TrainStep
    h = H.AttrDict(O.config("v1"))
    torch.manual_seed(1234)
    G = H.Generator(h)
    mpd = H.MultiPeriodDiscriminator()
    msd = H.MultiScaleDiscriminator()
    sds = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in (G, mpd, msd)]
    return h, TrainStep(G, mpd, msd, h, "cuda"), (G, mpd, msd), sds


LOSS_KEYS = ("loss_disc_f", "loss_disc_s", "loss_mel", "loss_fm_f", "loss_fm_s", "loss_gen_f", "loss_gen_s")


def test_train_step_vs_reference_golden(H):
    """Two full UPSTREAM steps (G fwd, D step + AdamW, G step through the updated D + AdamW) against the
    REFERENCE's own modules (tests/golden/train_step_seed1234.npz, torch autograd + torch.optim.AdamW, CPU fp32).
    Tolerances for the declared bf16-operand / fp32-accumulate numerics: losses 2e-2 relative
    (measured 3e-4), every parameter-gradient norm 5e-2 relative (measured <= 2.9e-2), gradients stored in full:
    cosine >= 0.999 and rel-L2 <= 5e-2, dL/dy_g_hat cosine >= 0.998.  Step-2 losses pin both AdamW updates."""
    from conftest import load_npz
    z = load_npz("train_step_seed1234.npz")
    h, ts, (G, mpd, msd), _ = _seeded_step(H)
    ya = torch.from_numpy(z["audio"]).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    for step in (1, 2):
        out = ts.step(x, ya.unsqueeze(1), y_mel)
        for k in LOSS_KEYS:
            ref = float(z[f"step{step}_{k}"])
            assert abs(out[k].item() - ref) <= 2e-2 * abs(ref), (step, k, out[k].item(), ref)
        if step > 1:
            continue
        dy, ref = ts.dy_audio.cpu().flatten(), torch.from_numpy(z["dy_g_hat"]).flatten()
        assert F.cosine_similarity(dy, ref, dim=0).item() >= 0.998
        for name, net in (("g", G), ("mpd", mpd), ("msd", msd)):
            named = dict(net.named_parameters())
            for k, n in zip([str(k) for k in z[f"{name}_keys"]], z[f"{name}_grad_norm"]):
                got = named[k].grad.norm().item()
                assert abs(got - n) <= 5e-2 * n + 1e-9, (name, k, got, n)
        for key in z.files:
            if "_grad::" not in key:
                continue
            name, k = key.split("_grad::")
            got = dict({"g": G, "mpd": mpd, "msd": msd}[name].named_parameters())[k].grad.cpu().flatten()
            ref = torch.from_numpy(z[key]).flatten()
            assert F.cosine_similarity(got, ref, dim=0).item() >= 0.999, key
            assert (got - ref).norm() <= 5e-2 * ref.norm(), key


def test_train_step_every_gradient_vs_oracle(H):
    """Every one of the 388 parameter gradients of one step against autograd over the CPU oracle
    (oracle/train_oracle.py, pinned to the reference by tests/test_oracle_cpu.py), batch of 3 ragged-content
    segments: per-tensor cosine >= 0.995 and rel-L2 <= 0.1 (bf16 activation gradients; the worst tensors are the
    bias gradients of the 32-channel stage, sums over 8192 x B bf16 values)."""
    from oracle import hifigan_oracle as O
    from oracle import train_oracle as TO
    h, ts, (G, mpd, msd), sds = _seeded_step(H)
    ya = O.synthetic_audio(3, 8192, seed=11)
    ya[2, 5000:] = 0.0                                  # a zero-padded tail, as MelDataset pads short files
    x_ref = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel_ref = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    sd_g, sd_p, sd_s = (TO.leaf_params(sd) for sd in sds)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    losses, gg, gp, gs, y_g, dy_g = TO.train_step(sd_g, sd_p, sd_s, h, x_ref, ya.unsqueeze(1), y_mel_ref)
    yc = ya.cuda()
    out = ts.step(H.mel_spectrogram(yc, 1024, 80, 22050, 256, 1024, 0, 8000), yc.unsqueeze(1),
                  H.mel_spectrogram(yc, 1024, 80, 22050, 256, 1024, 0, None))
    for k in LOSS_KEYS:
        assert abs(out[k].item() - losses[k]) <= 2e-2 * abs(losses[k]), (k, out[k].item(), losses[k])
    worst = (1.0, "")
    for net, grads in ((G, gg), (mpd, gp), (msd, gs)):
        for k, p in net.named_parameters():
            got, ref = p.grad.cpu().flatten(), grads[k].flatten()
            cos = F.cosine_similarity(got, ref, dim=0).item()
            rel = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
            worst = min(worst, (cos, k))
            assert cos >= 0.995 and rel <= 0.1, (k, cos, rel)
    print("worst cosine", worst)


def test_train_trajectory_vs_oracle_on_the_same_gpu(H):
    """Eight consecutive optimizer steps (fresh batch each step, graph replay from step 3 on) against the training
    oracle run in fp32 (TF32 off) on the same GPU: the losses of every step and the parameters after the last one.
    Pins the AdamW slices, the step counter, the cached / re-packed weights and the lane schedule over more than one
    update.  Tolerances: losses 3e-2 relative (bf16 activations over 8 updates), cosine of the accumulated parameter
    change >= 0.98 per network (AdamW's first updates are lr * sign(g)-like: elements whose gradient is below the
    bf16 noise floor take either sign, so this is a direction check, not an element-wise one)."""
    from oracle import hifigan_oracle as O
    from oracle import train_oracle as TO
    h, ts, (G, mpd, msd), sds = _seeded_step(H)
    init = [{k: v.clone() for k, v in sd.items()} for sd in sds]
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        batches = [O.synthetic_audio(4, 8192, seed=100 + i) for i in range(8)]
        mels = [(O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000),
                 O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)) for ya in batches]
        torch.set_default_device("cuda")      # the oracle builds its window / filterbank on the default device
        try:
            ref_sds = [TO.leaf_params({k: v.detach().clone().cuda() for k, v in sd.items()}) for sd in sds]
            optims = TO.make_optimizers(*ref_sds, h)
            ref_losses = []
            for ya, (x, y_mel) in zip(batches, mels):
                losses = TO.train_step(*ref_sds, h, x.cuda(), ya.cuda().unsqueeze(1), y_mel.cuda(), optims=optims)[0]
                ref_losses.append(losses)
        finally:
            torch.set_default_device("cpu")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    for i, (ya, (x, y_mel)) in enumerate(zip(batches, mels)):
        out = ts.step_graphed(x.cuda(), ya.cuda().unsqueeze(1), y_mel.cuda())
        for k in LOSS_KEYS:
            got, ref = out[k].item(), ref_losses[i][k]
            assert abs(got - ref) <= 3e-2 * abs(ref) + 1e-4, (i, k, got, ref)
    assert ts.graph_active
    for net, ref_sd, sd0 in zip((G, mpd, msd), ref_sds, init):
        mine = net.state_dict()
        num = den_a = den_b = 0.0
        for k, v in ref_sd.items():
            if not v.requires_grad:
                continue
            da = (mine[k].detach().float().cpu() - sd0[k]).flatten().double()
            db = (v.detach().cpu() - sd0[k]).flatten().double()
            num += float(da @ db); den_a += float(da @ da); den_b += float(db @ db)
        cos = num / (den_a ** 0.5 * den_b ** 0.5 + 1e-30)
        print("accumulated update cosine", type(net).__name__, round(cos, 4))
        assert cos >= 0.98, (type(net).__name__, cos)      # measured 0.996 (G), 0.9995 (MPD), 0.9996 (MSD)


def test_train_step_v3_and_no_update(H):
    """ResBlock2 generator (V3) through the same step; update=False leaves every parameter untouched."""
    from oracle import hifigan_oracle as O
    from oracle import train_oracle as TO
    from hifigan_b200.train import TrainStep
    h = H.AttrDict(O.config("v3"))
    torch.manual_seed(7)
    G, mpd, msd = H.Generator(h), H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator()
    sds = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in (G, mpd, msd)]
    ts = TrainStep(G, mpd, msd, h, "cuda")
    ya = O.synthetic_audio(2, 8192, seed=2)
    sd_g, sd_p, sd_s = (TO.leaf_params(sd) for sd in sds)
    losses, gg, _, _, _, _ = TO.train_step(sd_g, sd_p, sd_s, h, O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000),
                                           ya.unsqueeze(1), O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None),
                                           update=False)
    yc = ya.cuda()
    before = ts.G.flat.p.clone(), ts.D.flat.p.clone()
    out = ts.step(H.mel_spectrogram(yc, 1024, 80, 22050, 256, 1024, 0, 8000), yc.unsqueeze(1),
                  H.mel_spectrogram(yc, 1024, 80, 22050, 256, 1024, 0, None), update=False)
    assert torch.equal(before[0], ts.G.flat.p) and torch.equal(before[1], ts.D.flat.p)
    for k in LOSS_KEYS:
        assert abs(out[k].item() - losses[k]) <= 2e-2 * abs(losses[k]), (k, out[k].item(), losses[k])
    for k, p in G.named_parameters():
        got, ref = p.grad.cpu().flatten(), gg[k].flatten()
        assert F.cosine_similarity(got, ref, dim=0).item() >= 0.995, k


def test_graphed_step_matches_eager(H):
    """TrainStep.step_graphed (CUDA-graph replay, device-side AdamW step counter) walks the same trajectory as the
    eager step: losses after 4 steps agree to 1e-3 relative (fp32 atomics make the two runs non-bit-identical)."""
    from oracle import hifigan_oracle as O
    ya = O.synthetic_audio(2, 8192, seed=21).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    res = []
    for graphed in (False, True):
        h, ts, _, _ = _seeded_step(H)
        fn = ts.step_graphed if graphed else ts.step
        for _ in range(4):
            out = fn(x, ya.unsqueeze(1), y_mel)
        torch.cuda.synchronize()
        if graphed:
            assert isinstance(ts._graphs[next(iter(ts._graphs))], list), "capture fell back to eager"
        res.append({k: out[k].item() for k in LOSS_KEYS})
    for k in LOSS_KEYS:
        assert abs(res[0][k] - res[1][k]) <= 1e-3 * abs(res[0][k]) + 1e-5, (k, res[0][k], res[1][k])


def test_graph_capture_after_module_forward_keeps_repacking(H):
    """A validation-style `generator(x)` between the eager step and the capture step refreshes the host-side pack
    cache; the captured graph must still hold the re-pack launches (it would otherwise replay the generator on the
    bf16 weights of capture time for ever).  Losses after 5 steps and the generator's parameters follow the eager
    trajectory (1e-3 relative; fp32 atomics make runs non-bit-identical)."""
    from oracle import hifigan_oracle as O
    ya = O.synthetic_audio(2, 8192, seed=33).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    res, params = [], []
    for graphed in (False, True):
        h, ts, (G, _, _), _ = _seeded_step(H)
        for i in range(5):
            out = (ts.step_graphed if graphed else ts.step)(x, ya.unsqueeze(1), y_mel)
            if graphed and i == 0:                      # what train_loop.validate does at steps == 0
                ts.G.invalidate()
                G.eval()
                with torch.no_grad():
                    G(x)
                G.train()
        torch.cuda.synchronize()
        if graphed:
            assert ts.graph_active, "capture fell back to eager"
        res.append({k: out[k].item() for k in LOSS_KEYS})
        params.append(ts.G.flat.p.clone())
    for k in LOSS_KEYS:
        assert abs(res[0][k] - res[1][k]) <= 2e-3 * abs(res[0][k]) + 1e-5, (k, res[0][k], res[1][k])
    cos = F.cosine_similarity(params[0], params[1], dim=0).item()
    assert (params[0] - params[1]).norm() <= 2e-3 * params[0].norm(), (params[0] - params[1]).norm().item()
    assert cos > 0.999999


def test_spectral_norm_kernels_vs_torch(H):
    """hg_spectral_norm_fwd / _bwd against torch.nn.utils.spectral_norm itself (train-mode forward = one power
    iteration updating u / v in place; autograd through W / sigma with u, v constant)."""
    from hifigan_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda")
    for shape in [(128, 1, 15), (1024, 64, 41), (1, 1024, 3), (256, 8, 41)]:
        torch.manual_seed(sum(shape))
        m = torch.nn.utils.spectral_norm(torch.nn.Conv1d(shape[1], shape[0], shape[2])).to(dev).train()
        w = m.weight_orig.detach().clone()
        u, v = m.weight_u.detach().clone(), m.weight_v.detach().clone()
        rows, cols = shape[0], shape[1] * shape[2]
        for it in range(2):                       # two consecutive calls: the state must carry over
            x = torch.randn(2, shape[1], 50, device=dev)
            y = m(x)                              # runs the power iteration, builds m.weight
            d = torch.randn_like(m.weight)
            m.weight_orig.grad = None
            (m.weight * d).sum().backward()
            eff = torch.empty(rows, cols, device=dev)
            sig, uc, vc = torch.empty(1, device=dev), torch.empty(rows, device=dev), torch.empty(cols, device=dev)
            ws = torch.empty(rows + cols + 4, device=dev)
            _lib.check(L.hg_spectral_norm_fwd(w.data_ptr(), u.data_ptr(), v.data_ptr(), rows, cols, 1, eff.data_ptr(),
                                              sig.data_ptr(), uc.data_ptr(), vc.data_ptr(), ws.data_ptr(), _st()))
            assert torch.allclose(eff.view_as(m.weight), m.weight.detach(), rtol=1e-4, atol=1e-6)
            assert torch.allclose(u, m.weight_u, rtol=1e-4, atol=1e-6) and torch.allclose(v, m.weight_v, rtol=1e-4, atol=1e-6)
            out = torch.full((rows, cols), 1.0, device=dev)
            _lib.check(L.hg_spectral_norm_bwd(d.contiguous().data_ptr(), eff.data_ptr(), uc.data_ptr(), vc.data_ptr(),
                                              sig.data_ptr(), rows, cols, 1, out.data_ptr(), ws.data_ptr(), _st()))
            ref = m.weight_orig.grad.reshape(rows, cols) + 1.0
            assert torch.allclose(out, ref, rtol=1e-3, atol=1e-5), (out - ref).abs().max()
        # eval mode: stored u, v, no update
        u0 = u.clone()
        _lib.check(L.hg_spectral_norm_fwd(w.data_ptr(), u.data_ptr(), v.data_ptr(), rows, cols, 0, eff.data_ptr(),
                                          sig.data_ptr(), 0, 0, ws.data_ptr(), _st()))
        m.eval()
        m(x)
        assert torch.equal(u, u0) and torch.allclose(eff.view_as(m.weight), m.weight.detach(), rtol=1e-4, atol=1e-6)


def test_train_driver_checkpoints_and_resumes(H, tmp_path):
    """hifigan_b200.train_loop (UPSTREAM train.py's command line): trains on a tiny synthetic wav
    set, writes g_* / do_* files in the reference's format at the configured cadence, validates, and a second run
    resumes from the newest pair (steps, epoch, optimizer moments) and continues."""
    import json
    from scipy.io.wavfile import write
    from hifigan_b200 import train_loop
    from oracle import hifigan_oracle as O
    cfg = dict(H.load_config("v1"))
    cfg.update(batch_size=2, seed=1234, num_gpus=0)
    (tmp_path / "config_v1.json").write_text(json.dumps(cfg))
    wavs = tmp_path / "wavs"
    wavs.mkdir()
    audio = O.synthetic_audio(5, 12000, seed=9)
    names = [f"utt{i}" for i in range(5)]
    for n, a in zip(names, audio):
        write(wavs / f"{n}.wav", 22050, (a.numpy() * 32767).astype(np.int16))
    (tmp_path / "training.txt").write_text("\n".join(f"{n}|text" for n in names[:4]))
    (tmp_path / "validation.txt").write_text(f"{names[4]}|text")
    cp = tmp_path / "cp"
    argv = ["--input_wavs_dir", str(wavs), "--input_training_file", str(tmp_path / "training.txt"),
            "--input_validation_file", str(tmp_path / "validation.txt"), "--checkpoint_path", str(cp),
            "--config", str(tmp_path / "config_v1.json"), "--stdout_interval", "1", "--checkpoint_interval", "2",
            "--validation_interval", "3", "--summary_interval", "1000"]
    r1 = train_loop.main(argv + ["--training_epochs", "2"])
    assert r1["final_steps"] == 4 and (cp / "g_00000002").exists() and (cp / "do_00000002").exists()
    assert (cp / "config.json").exists() and np.isfinite(r1["val_mel_error"]) and np.isfinite(r1["loss_gen_all"])
    do = torch.load(cp / "do_00000002", map_location="cpu")
    assert set(do) == {"mpd", "msd", "optim_g", "optim_d", "steps", "epoch"} and do["steps"] == 2
    # the optimizer state is torch.optim.AdamW's format, parameter order chain(msd, mpd) as UPSTREAM
    msd, mpd = H.MultiScaleDiscriminator(), H.MultiPeriodDiscriminator()
    opt = torch.optim.AdamW(list(msd.parameters()) + list(mpd.parameters()), 2e-4, betas=[0.8, 0.99])
    opt.load_state_dict(do["optim_d"])
    assert float(opt.state_dict()["state"][0]["step"]) == 3.0     # three D updates before the save at steps == 2
    g = torch.load(cp / "g_00000002", map_location="cpu")
    assert set(g) == {"generator"} and "conv_pre.weight_g" in g["generator"]
    r2 = train_loop.main(argv + ["--training_epochs", "3"])
    assert r2["final_steps"] == 3 + 2 * 2 and (cp / "g_00000004").exists()
    # fine-tuning mode (UPSTREAM --fine_tuning True: input mels from <input_mels_dir>/<name>.npy, no peak normalisation)
    mels = tmp_path / "mels"
    mels.mkdir()
    for n, a in zip(names, audio):
        m = O.mel_spectrogram(a.unsqueeze(0) / 32768.0 * 32767 / 32768.0, 1024, 80, 22050, 256, 1024, 0, 8000)
        np.save(mels / f"{n}.npy", m.numpy().astype(np.float32))
    cp_ft = tmp_path / "cp_ft"
    argv_ft = [x if x != str(cp) else str(cp_ft) for x in argv] + ["--fine_tuning", "True", "--input_mels_dir", str(mels)]
    r3 = train_loop.main(argv_ft + ["--training_epochs", "2"])
    assert r3["final_steps"] == 4 and (cp_ft / "g_00000002").exists()
    assert np.isfinite(r3["val_mel_error"]) and np.isfinite(r3["loss_gen_all"])


@pytest.mark.gpu
def test_train_loop_validate_matches_oracle(H):
    """train_loop.validate (UPSTREAM validation: mean over files of L1(mel(y), mel(G(mel(y)))) on whole utterances,
    fmax_for_loss) against the same number from the CPU oracle's generator_forward + mel_spectrogram."""
    from hifigan_b200 import train_loop
    from oracle import hifigan_oracle as O
    h = H.load_config("v1")
    torch.manual_seed(1234)
    G = H.Generator(h)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    utts = [u for u in O.synthetic_audio(2, 9000, seed=11)]          # 9000 samples: not a multiple of the hop
    ref = 0.0
    for y in utts:
        frames = y.numel() // h.hop_size
        yc = y[: frames * h.hop_size].reshape(1, -1)
        x = O.mel_spectrogram(yc, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax)
        with torch.no_grad():
            yg = O.generator_forward(sd, h, x.float())
        m0 = O.mel_spectrogram(yc, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax_for_loss)
        m1 = O.mel_spectrogram(yg.reshape(1, -1).to(yc.dtype), h.n_fft, h.num_mels, h.sampling_rate, h.hop_size,
                               h.win_size, h.fmin, h.fmax_for_loss)
        ref += (m0 - m1).abs().mean().item()
    ref /= len(utts)
    got = train_loop.validate(G.cuda(), utts, h, torch.device("cuda"))
    # tolerance: the generated waveform carries the bf16 forward's error (SNR >= 40 dB); in the log-mel domain of a
    # random-init generator that is well under one per cent of the L1 distance
    assert abs(got - ref) / ref < 1e-2, (got, ref)
