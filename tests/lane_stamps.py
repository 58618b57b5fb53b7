"""Where the lanes of one graph-replayed training step finish (hg_timestamp marks inside the graph).

    HG_LANE_STAMPS=1 python tests/lane_stamps.py [per-gpu batch]                               (1 GPU)
    HG_LANE_STAMPS=1 python -m torch.distributed.run --nproc-per-node N ... tests/lane_stamps.py   (N ranks; rank 0 prints)

Prints, per sub-discriminator lane (0-4 periods 2,3,5,7,11; 5-7 scales), the time its D-step backward ends, the time
its gradient all-reduce returns (N > 1: the gap is the exposed exchange), and the end of its generator-step pass; then
the generator backward / all-reduce / end of step.  Appends to gpurun_out/lane_stamps.txt.
"""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HG_LANE_STAMPS", "1")
H = importlib.import_module("hifi-gan_b200")
train = importlib.import_module("hifi-gan_b200.train")
configs = importlib.import_module("hifi-gan_b200.configs")


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    h = configs.load_config("v1")
    torch.manual_seed(1234)
    ts = train.TrainStep(H.Generator(h), H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator(), h, dev)
    g = torch.Generator().manual_seed(100 + rank)
    y = (0.5 * torch.sin(torch.arange(8192) * 0.05) + 0.1 * torch.randn(batch, 8192, generator=g)).clamp(-0.95, 0.95).to(dev)
    x = H.mel_spectrogram(y, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax)
    ym = H.mel_spectrogram(y, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax_for_loss)
    y3 = y.unsqueeze(1)
    for _ in range(ts.warmup_calls + 3):
        ts.step_graphed(x, y3, ym)
    torch.cuda.synchronize()
    rows = []
    for _ in range(5):
        ts.step_graphed(x, y3, ym)
        torch.cuda.synchronize()
        rows.append(ts.stamps.read())
    graph_on, order = ts.graph_active, list(ts.D.order)
    n = len(ts.D.subs)
    del ts            # the captured graphs hold NCCL kernels: release them before the communicator goes away
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        med = {k: sorted(r[k] for r in rows)[len(rows) // 2] for k in rows[0]}
        out = {"world": world, "per_gpu_batch": batch, "graph": graph_on, "order": order,
               "us": {k: round(v, 1) for k, v in med.items()}}
        print(json.dumps(out))
        print(f"world {world}, batch {batch}/GPU, median of 5 replays (us since step start); G forward done {med['g_fwd_done']:.0f}")
        for i in range(n):
            ar = med.get(f"d{i}_allreduce_done")
            print(f"  lane {i}: backward done {med[f'd{i}_bwd_done']:8.0f}   all-reduce returned "
                  f"{(ar if ar is not None else float('nan')):8.0f}   generator-step pass done {med[f'd{i}_gstep_done']:8.0f}")
        print(f"  G backward {med['g_bwd_start']:.0f} -> {med['g_bwd_done']:.0f}, all-reduce returned "
              f"{med['g_allreduce_done']:.0f}, end {med['end']:.0f}")
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/lane_stamps.txt", "a") as f:
            f.write(json.dumps(out) + "\n")
    sys.stdout.flush()
    if world > 1:
        torch.distributed.barrier()
        os._exit(0)       # skip the communicator teardown: it can wait for ever on graph-captured collectives


if __name__ == "__main__":
    main()
