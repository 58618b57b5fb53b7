"""Per-kernel totals of one graph-replayed training step (CUPTI via torch.profiler; durations are taken while the
step's lanes overlap, so they include time spent sharing the SMs).  Tool.   python tests/train_kernels.py [batch=16]"""
import collections, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hifigan_b200 as H
from hifigan_b200.configs import load_config
from hifigan_b200.train import TrainStep
from torch.profiler import ProfilerActivity, profile

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
h = load_config("v1")
torch.manual_seed(1234)
ts = TrainStep(H.Generator(h), H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator(), h, "cuda")
g = torch.Generator().manual_seed(3)
y = (torch.rand(b, 8192, generator=g) * 1.6 - 0.8).cuda()
x = H.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000)
ym = H.mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, None)
for _ in range(5):
    ts.step_graphed(x, y.unsqueeze(1), ym)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        ts.step_graphed(x, y.unsqueeze(1), ym)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
ev = []
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CUDA:
        continue
    n = e.name.replace("(anonymous namespace)::", "").replace("void ", "")[:60]
    a = agg[n]
    a[0] += 1; a[1] += e.time_range.elapsed_us(); a[2] = max(a[2], e.time_range.elapsed_us())
    ev.append((e.time_range.start, e.time_range.elapsed_us(), n))
tot = sum(v[1] for v in agg.values())
print(f"2 replays: {len(ev)} device activities, summed {tot / 2e3:.2f} ms per step")
for n, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{n:62s} x{c / 2:6.1f}  {t / 2e3:7.3f} ms/step  max {mx:7.1f} us")
ev.sort()
t0, t1 = ev[0][0], max(s + d for s, d, _ in ev)
print(f"span of the two replays: {(t1 - t0) / 1e3:.2f} ms")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
full = os.path.join(root, "gpurun_out", "train_trace_full.json")
prof.export_chrome_trace(full)
tr = json.load(open(full))
ks = [{"ts": e["ts"], "dur": e["dur"], "n": e["name"].replace("(anonymous namespace)::", "").replace("void ", "")[:70],
       "s": e.get("args", {}).get("stream"), "grid": e.get("args", {}).get("grid"), "pri": e.get("args", {}).get("priority")}
      for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ks.sort(key=lambda e: e["ts"])
json.dump(ks, open(os.path.join(root, "gpurun_out", "train_kernels.json"), "w"))
os.remove(full)
