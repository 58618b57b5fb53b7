"""Bring-up driver for the training step (not a pytest): runs TrainStep for two steps from the seeded weights and
prints, per parameter, our gradient norm vs the reference's (tests/golden/train_step_seed1234.npz), cosines for
the gradients stored in full, dL/dy_g_hat agreement, losses; optionally timing.
    python tests/gpu_bringup_train.py [parity|time B]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hifigan_b200 as H  # noqa: E402
from hifigan_b200 import _lib  # noqa: E402
from hifigan_b200.train import TrainStep  # noqa: E402

CFG = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
           upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
           resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], segment_size=8192, num_mels=80, n_fft=1024,
           hop_size=256, win_size=1024, sampling_rate=22050, fmin=0, fmax=8000, fmax_for_loss=None,
           learning_rate=2e-4, adam_b1=0.8, adam_b2=0.99)


def build():
    h = H.AttrDict(CFG)
    torch.manual_seed(1234)
    G = H.Generator(h)
    mpd = H.MultiPeriodDiscriminator()
    msd = H.MultiScaleDiscriminator()
    return h, TrainStep(G, mpd, msd, h, "cuda"), (G, mpd, msd)


def parity():
    z = np.load(os.path.join(ROOT, "tests", "golden", "train_step_seed1234.npz"))
    h, ts, (G, mpd, msd) = build()
    ya = torch.from_numpy(z["audio"]).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    for step in (1, 2):
        out = ts.step(x, ya.unsqueeze(1), y_mel)
        torch.cuda.synchronize()
        for k in ("loss_disc_f", "loss_disc_s", "loss_mel", "loss_fm_f", "loss_fm_s", "loss_gen_f", "loss_gen_s"):
            ref = float(z[f"step{step}_{k}"])
            print(f"step{step} {k:12s} ours {out[k].item():.5f} ref {ref:.5f} rel {abs(out[k].item() - ref) / abs(ref):.2e}")
        if step == 1:
            dy = ts.dy_audio.cpu()
            ref = torch.from_numpy(z["dy_g_hat"]).reshape(dy.shape)
            cos = torch.nn.functional.cosine_similarity(dy.flatten(), ref.flatten(), dim=0).item()
            print(f"dy_g_hat: norm ours {dy.norm():.4e} ref {ref.norm():.4e} cos {cos:.5f}")
            for name, net in (("g", G), ("mpd", mpd), ("msd", msd)):
                keys = [str(k) for k in z[f"{name}_keys"]]
                norms = z[f"{name}_grad_norm"]
                named = dict(net.named_parameters())
                worst = []
                for k, n in zip(keys, norms):
                    got = named[k].grad.norm().item()
                    worst.append((abs(got - n) / (n + 1e-12), k, got, n))
                worst.sort(reverse=True)
                print(f"{name}: {len(keys)} tensors; worst norm mismatches:")
                for w in worst[:8]:
                    print(f"   rel {w[0]:.3e}  {w[1]:45s} ours {w[2]:.4e} ref {w[3]:.4e}")
                med = sorted(w[0] for w in worst)[len(worst) // 2]
                print(f"   median rel {med:.3e}")
            for key in z.files:
                if "_grad::" in key:
                    name, k = key.split("_grad::")
                    net = {"g": G, "mpd": mpd, "msd": msd}[name]
                    got = dict(net.named_parameters())[k].grad.cpu().flatten()
                    ref = torch.from_numpy(z[key]).flatten()
                    cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
                    print(f"   full {key:55s} cos {cos:.5f} relL2 {(got - ref).norm() / ref.norm():.3e}")


def timing(b):
    h, ts, _ = build()
    from oracle import hifigan_oracle as O
    ya = O.synthetic_audio(b, 8192, seed=3).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    y = ya.unsqueeze(1)
    fn = ts.step if os.environ.get("HG_TRAIN_EAGER") else ts.step_graphed
    for _ in range(3):
        fn(x, y, y_mel)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record()
    iters = 5
    for _ in range(iters):
        out = fn(x, y, y_mel)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / iters
    print(json.dumps({"batch": b, "ms_per_step": ms, "segments_per_s": b / ms * 1e3, "wall_ms": (time.time() - t0) / iters * 1e3,
                      "launches_per_step": (_lib.launch_count() - n0) / iters, "graphed": fn is not ts.step,
                      "loss_gen_all": out["loss_gen_all"].item(), "loss_disc_all": out["loss_disc_all"].item()}))


def profile(b):
    """one step between cudaProfilerStart/Stop (ncu --profile-from-start off)"""
    h, ts, _ = build()
    from oracle import hifigan_oracle as O
    ya = O.synthetic_audio(b, 8192, seed=3).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    y = ya.unsqueeze(1)
    for _ in range(2):
        ts.step(x, y, y_mel)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ts.step(x, y, y_mel)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def convfloor():
    """per-launch cost of the tcgen05 conv / wgrad kernels on small problems (back-to-back launches on one stream)"""
    L = _lib.lib()
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    for (b, t, c, k, d) in [(1, 128, 64, 3, 1), (1, 128, 256, 3, 1), (16, 256, 256, 11, 5), (16, 2048, 128, 7, 3),
                            (16, 4096, 64, 11, 5), (16, 8192, 32, 11, 5), (16, 8192, 32, 3, 1)]:
        x = torch.randn(b, t, c, device=dev).bfloat16()
        w = torch.randn(k, c, c, device=dev).bfloat16()
        out = torch.empty_like(x)
        dwp = torch.zeros(k, c, c, device=dev)
        pad = (k - 1) * d // 2
        res = {}
        for name, fn in (("fwd", lambda: L.hg_conv1d_fwd(x.data_ptr(), w.data_ptr(), 0, b, t, c, c, k, d, pad, 0, 0, 0, 1.0,
                                                         out.data_ptr(), 0, 0.1, 0, 1, st)),
                         ("wgrad", lambda: L.hg_conv1d_wgrad(x.data_ptr(), out.data_ptr(), b, t, c, t, t, 1, c, k, 1, d, pad,
                                                             dwp.data_ptr(), 1, st))):
            for _ in range(5):
                _lib.check(fn())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 200 * 1e3
        flop = 2.0 * b * t * c * c * k
        print(f"b={b} t={t} c={c} k={k}: fwd {res['fwd']:.1f} us ({flop / res['fwd'] / 1e6:.0f} TFLOP/s)  "
              f"wgrad {res['wgrad']:.1f} us ({flop / res['wgrad'] / 1e6:.0f} TFLOP/s)")


def firstbwd():
    """micro-benchmark of hg_disc_first_conv_bwd in its two modes at the MSD scale-1 shape"""
    L = _lib.lib()
    dev = torch.device("cuda")
    for (b, t, period, k, s, pad, cout) in [(16, 8192, 1, 15, 1, 7, 128), (32, 8192, 1, 15, 1, 7, 128), (16, 8192, 2, 5, 3, 2, 32)]:
        y = torch.randn(b, t, device=dev)
        w = torch.randn(cout, k, device=dev)
        h_in = (t + period - 1) // period
        h_out = (h_in + 2 * pad - k) // s + 1
        dpre = torch.randn(b * period, h_out, cout, device=dev).bfloat16()
        dw, db, dy = torch.zeros(cout, k, device=dev), torch.zeros(cout, device=dev), torch.zeros(b, t, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for mode in ("dw", "dy"):
            args = (dw.data_ptr(), db.data_ptr(), 0) if mode == "dw" else (0, 0, dy.data_ptr())
            for _ in range(3):
                _lib.check(L.hg_disc_first_conv_bwd(y.data_ptr(), w.data_ptr(), dpre.data_ptr(), b, t, period, k, s, pad,
                                                    cout, h_out, *args, st))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                _lib.check(L.hg_disc_first_conv_bwd(y.data_ptr(), w.data_ptr(), dpre.data_ptr(), b, t, period, k, s, pad,
                                                    cout, h_out, *args, st))
            e1.record()
            torch.cuda.synchronize()
            print(f"first_bwd b={b} period={period} cout={cout} mode={mode}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")


def trace(b):
    """torch.profiler (CUPTI) kernel timeline of two graph replays -> gpurun_out/train_trace.json (kernels only)"""
    from torch.profiler import ProfilerActivity, profile as tprofile
    h, ts, _ = build()
    from oracle import hifigan_oracle as O
    ya = O.synthetic_audio(b, 8192, seed=3).cuda()
    x = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
    y_mel = H.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
    y = ya.unsqueeze(1)
    for _ in range(4):
        ts.step_graphed(x, y, y_mel)
    torch.cuda.synchronize()
    with tprofile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            ts.step_graphed(x, y, y_mel)
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            ev.append({"name": e.name[:80], "start": e.time_range.start, "dur": e.time_range.elapsed_us()})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    prof.export_chrome_trace(os.path.join(ROOT, "gpurun_out", "train_trace_full.json"))
    print("events", len(ev))


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if mode == "parity":
        parity()
    elif mode == "convfloor":
        convfloor()
    elif mode == "firstbwd":
        firstbwd()
    elif mode == "trace":
        trace(int(sys.argv[2]) if len(sys.argv) > 2 else 16)
    elif mode == "profile":
        profile(int(sys.argv[2]) if len(sys.argv) > 2 else 16)
    else:
        timing(int(sys.argv[2]) if len(sys.argv) > 2 else 16)
