"""Hardware bring-up script (run under gpurun; not collected by pytest).

    python tests/gpu_bringup.py <stage> [...]

Each stage runs in its own process (see tests/run_bringup.sh) so a faulting kernel cannot poison the rest.
Prints one JSON line per check to stdout and appends it to gpurun_out/bringup.jsonl.
"""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hifigan_b200 as H  # noqa: E402
from hifigan_b200 import _lib  # noqa: E402
from oracle import hifigan_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    with open(os.path.join(OUT, "bringup.jsonl"), "a") as f:
        f.write(line + "\n")


def stage_mel():
    dev = torch.device("cuda")
    for (b, t) in [(4, 8192), (1, 22050), (3, 40000)]:
        y = O.synthetic_audio(b, t, seed=b)
        for fmax in (8000, None):
            ref = O.mel_spectrogram(y.double(), 1024, 80, 22050, 256, 1024, 0, fmax)
            got = H.mel_spectrogram(y.to(dev), 1024, 80, 22050, 256, 1024, 0, fmax).cpu().double()
            emit(stage="mel", b=b, t=t, fmax=fmax, shape=list(got.shape), max_abs=(got - ref).abs().max().item())
    H.meldataset.flush_range_warnings()


def conv_case(L, b, t, cin, cout, k, d, mode, seed=0, with_res=True):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(b, t, cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    res = torch.randn(b, t, cout, generator=g).to(dev).bfloat16() if with_res else None
    wp = torch.empty(k, cout, cin, dtype=torch.bfloat16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.hg_pack_conv1d_weight(w.data_ptr(), 0, cout, cin, k, cin, wp.data_ptr(), st))
    out_raw = torch.zeros(b, t, cout, dtype=torch.bfloat16, device=dev)
    out_act = torch.zeros(b, t, cout, dtype=torch.bfloat16, device=dev)
    pad = (k - 1) * d // 2
    _lib.check(L.hg_conv1d_fwd(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), b, t, cin, cout, k, d, pad,
                               0 if res is None else res.data_ptr(), 0, 0, 0.5, out_raw.data_ptr(),
                               out_act.data_ptr(), 0.1, 0, 1, st))
    torch.cuda.synchronize()
    wr = wp.float().permute(1, 2, 0).contiguous()  # [cout, cin, k] bf16-rounded
    ref = F.conv1d(x.float().transpose(1, 2), wr, bias, dilation=d, padding=pad).transpose(1, 2)
    if res is not None:
        ref = ref + res.float()
    ref = ref * 0.5
    e_raw = (out_raw.float() - ref).abs().max().item()
    e_act = (out_act.float() - F.leaky_relu(ref, 0.1)).abs().max().item()
    return e_raw, e_act, ref.abs().max().item()


def stage_conv(mode):
    L = _lib.lib()
    cases = [
        # b, t, cin, cout, k, d
        (1, 128, 64, 64, 1, 1),
        (1, 128, 64, 64, 3, 1),
        (2, 256, 64, 64, 3, 3),
        (2, 384, 128, 128, 7, 5),
        (1, 512, 256, 256, 11, 5),
        (3, 200, 128, 256, 3, 1),
        (2, 256, 32, 32, 3, 1),
        (2, 256, 32, 32, 11, 5),
        (2, 256, 32, 64, 7, 3),
        (1, 100, 128, 512, 7, 1),
    ]
    for c in cases:
        try:
            e_raw, e_act, mag = conv_case(L, *c, mode)
            emit(stage="conv", mode=mode, case=list(c), err_raw=e_raw, err_act=e_act, ref_max=mag,
                 ok=bool(e_raw < 0.03 * max(mag, 1.0)))
        except Exception as e:  # noqa: BLE001
            emit(stage="conv", mode=mode, case=list(c), error=str(e)[:300])
            raise


def stage_gen(version, b, frames, mode):
    h = H.AttrDict(O.config(version))
    torch.manual_seed(1234)
    G = H.Generator(h)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    torch.manual_seed(0)
    x = torch.randn(b, 80, frames)
    taps = {}
    with torch.no_grad():
        ref = O.generator_forward(sd, h, x, taps)
    G = G.cuda().eval()
    with torch.no_grad():
        y = G(x.cuda()).float().cpu()
    torch.cuda.synchronize()
    err = (y - ref).abs().max().item()
    num = (ref - ref.mean()).pow(2).sum()
    snr = 10 * torch.log10(num / (y - ref).pow(2).sum()).item()
    emit(stage="gen", version=version, b=b, frames=frames, mode=mode, max_abs=err, ref_max=ref.abs().max().item(),
         snr_db=snr)
    # after remove_weight_norm the result must not move
    G.remove_weight_norm()
    with torch.no_grad():
        y2 = G(x.cuda()).float().cpu()
    emit(stage="gen_folded", version=version, max_abs_vs_wn=(y2 - y).abs().max().item())


def stage_time(version, b, frames, mode, iters=5):
    h = H.AttrDict(O.config(version))
    torch.manual_seed(1234)
    G = H.Generator(h).cuda().eval()
    G.remove_weight_norm()
    x = torch.randn(b, 80, frames, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            G(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            G(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    samples = b * frames * 256
    emit(stage="time", version=version, b=b, frames=frames, ms=ms, samples_per_s=samples / ms * 1e3,
         xrt=samples / ms * 1e3 / 22050, tflops=samples * (2398848 if version == "v1" else 175648) / ms / 1e9)


def stage_layers(b, frames):
    """Per-layer-shape timing of hg_conv1d_fwd at the V1 stage shapes (CUDA events, 5 iterations)."""
    L = _lib.lib()
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    shapes = [(512, 512 * 0 + 2048, 3, 1, frames, "ups0 polyphase"), (256, 1024, 3, 1, frames * 8, "ups1 polyphase"),
              (128, 128, 3, 1, frames * 64, "ups2 polyphase"), (64, 64, 3, 1, frames * 128, "ups3 polyphase")]
    for c, t in ((256, frames * 8), (128, frames * 64), (64, frames * 128), (32, frames * 256)):
        for k in (3, 7, 11):
            for d in (1, 5):
                shapes.append((c, c, k, d, t, "resblock"))
    for cin, cout, k, d, t, name in shapes:
        x = torch.randn(b, t, cin, device=dev).bfloat16()
        wp = (torch.randn(k, cout, cin, device=dev) / (cin * k) ** 0.5).bfloat16()
        bias = torch.zeros(cout, device=dev)
        res = torch.randn(b, t, cout, device=dev).bfloat16()
        o1 = torch.empty(b, t, cout, dtype=torch.bfloat16, device=dev)
        o2 = torch.empty(b, t, cout, dtype=torch.bfloat16, device=dev)
        pad = (k - 1) * d // 2
        def run():
            _lib.check(L.hg_conv1d_fwd(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), b, t, cin, cout, k, d, pad,
                                       res.data_ptr(), 0, 0, 1.0, o1.data_ptr(), o2.data_ptr(), 0.1, 0, 1, st))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flops = 2.0 * b * t * cin * cout * k
        byts = 2.0 * b * t * (cin + 3 * cout)
        emit(stage="layer", name=name, cin=cin, cout=cout, k=k, d=d, t=t, b=b, ms=round(ms, 4),
             tflops=round(flops / ms / 1e9, 1), gbs=round(byts / ms / 1e6, 1))
        del x, wp, res, o1, o2


def stage_pairs(b, frames, debug=0):
    """Timing of hg_resblock_pair_fwd at the V1 narrow-stage shapes vs the two-launch path."""
    L = _lib.lib()
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    for c, t in ((64, frames * 128), (32, frames * 256)):
        for k in (3, 7, 11):
            for d in (1, 5):
                if not L.hg_resblock_pair_supported(c, k, d):
                    emit(stage="pair", c=c, k=k, d=d, supported=False)
                    continue
                x = torch.randn(b, t, c, device=dev).bfloat16()
                w1 = (torch.randn(k, c, c, device=dev) / (c * k) ** 0.5).bfloat16()
                w2 = (torch.randn(k, c, c, device=dev) / (c * k) ** 0.5).bfloat16()
                b1 = torch.zeros(c, device=dev)
                o1 = torch.empty(b, t, c, dtype=torch.bfloat16, device=dev)
                def run():
                    _lib.check(L.hg_resblock_pair_fwd(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                                      b1.data_ptr(), b, t, c, k, d, 0.1, 0, 0, 1.0, o1.data_ptr(), 0,
                                                      0.1, 0, 1, st))
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                flops = 2.0 * 2 * b * t * c * c * k
                emit(stage="pair", debug=debug, c=c, k=k, d=d, t=t, b=b, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1),
                     gbs=round(2.0 * b * t * c * 2 / ms / 1e6, 1))
                del x, o1


def stage_post(b, t):
    L = _lib.lib()
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(b, t, 32, device=dev).bfloat16()
    w = torch.randn(32, 7, device=dev) * 0.1
    bias = torch.zeros(1, device=dev)
    y = torch.empty(b, t, device=dev)
    def run():
        _lib.check(L.hg_conv_post_tanh_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), b, t, 32, 7, y.data_ptr(), st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = torch.tanh(F.conv1d(x.float().transpose(1, 2), w.unsqueeze(0), bias, padding=3))[:, 0]
    emit(stage="post", b=b, t=t, ms=round(ms, 4), gbs=round(b * t * 68 / ms / 1e6, 1), err=(y - ref).abs().max().item())


def stage_melperf():
    """cfg5 sweep: fused mel kernel throughput vs the 1344 B/frame HBM roofline, and vs torchaudio on the GPU."""
    import torchaudio
    ref = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, win_length=1024, hop_length=256,
                                               f_min=0, f_max=8000, n_mels=80, center=False).cuda()
    for t in (8192, 262144):
        for b in (1, 4, 16, 64, 256):
            if b * t > 256 * 262144 // 4:
                continue
            y = (torch.rand(b, t, device="cuda") * 1.9 - 0.95)
            def ours():
                return H.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
            def theirs():
                yp = F.pad(y.unsqueeze(1), (384, 384), mode="reflect").squeeze(1)
                return torch.log(torch.clamp(ref(yp), min=1e-5))
            res = {}
            for name, fn in (("ours", ours), ("torchaudio", theirs)):
                for _ in range(3):
                    out = fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 20
                e0.record()
                for _ in range(n):
                    out = fn()
                e1.record()
                torch.cuda.synchronize()
                res[name] = e0.elapsed_time(e1) / n
            frames = out.shape[0] * out.shape[2]
            emit(stage="melperf", b=b, t=t, frames=frames, ms=round(res["ours"], 4),
                 frames_per_s=round(frames / res["ours"] * 1e3), gbs=round(frames * 1344 / res["ours"] / 1e6, 1),
                 torchaudio_ms=round(res["torchaudio"], 4))
    H.meldataset.flush_range_warnings()


def stage_disc(b, t):
    """Timing of the discriminator forwards (both inputs), CUDA events."""
    torch.manual_seed(1234)
    mpd, msd = H.MultiPeriodDiscriminator().cuda().eval(), H.MultiScaleDiscriminator().cuda().eval()
    y = O.synthetic_audio(b, t, seed=1).unsqueeze(1).cuda()
    y_hat = O.synthetic_audio(b, t, seed=2).unsqueeze(1).cuda()
    for name, D, macs in (("mpd", mpd, 4719952768), ("msd", msd, 3930822400)):
        with torch.no_grad():
            for _ in range(2):
                D(y, y_hat)
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                D(y, y_hat)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        emit(stage="disc", which=name, b=b, t=t, ms=round(ms, 3), launches=(_lib.launch_count() - n0) // 5,
             tflops=round(2 * 2 * b * macs * (t / 8192) / ms / 1e9, 1))


if __name__ == "__main__":
    st = sys.argv[1]
    t0 = time.time()
    if st == "mel":
        stage_mel()
    elif st == "conv":
        stage_conv(int(sys.argv[2]))
    elif st == "gen":
        stage_gen(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    elif st == "post":
        stage_post(int(sys.argv[2]), int(sys.argv[3]))
    elif st == "melperf":
        stage_melperf()
    elif st == "disc":
        stage_disc(int(sys.argv[2]), int(sys.argv[3]))
    elif st == "pairs":
        stage_pairs(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    elif st == "layers":
        stage_layers(int(sys.argv[2]), int(sys.argv[3]))
    elif st == "time":
        stage_time(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    emit(stage="done", which=sys.argv[1:], seconds=time.time() - t0)
