"""Summarise gpurun_out/train_trace.json (kernel timeline of a graph replay written by gpu_bringup_train.py trace):
per-stream busy time, phase boundaries, concurrency histogram, top kernels of the busiest lanes."""
import collections
import json
import sys


def main(path, top=3):
    ev = json.load(open(path))
    ev.sort(key=lambda e: e["ts"])
    # one step = the kernels between two generator AdamW launches (the generator's update is the last kernel of a
    # step, on the main stream; the sub-discriminators update slice by slice on their own lanes)
    main = [e for e in ev if "adamw" in e["n"]][-1]["s"]
    ad = [i for i, e in enumerate(ev) if "adamw" in e["n"] and e["s"] == main]
    step = ev[ad[0] + 1:ad[1] + 1]
    t0 = step[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in step)
    print(f"step span {(t1 - t0) / 1e3:.2f} ms, {len(step)} kernels, summed durations {sum(e['dur'] for e in step) / 1e3:.2f} ms")
    for label, name in [("G fwd end", "conv_post_tanh_kernel"), ("D bwd start", "loss_grad_kernel"), ("first adamw D", "adamw_kernel"),
                        ("mel_bwd", "mel_bwd_kernel"), ("G bwd start", "conv_post_bwd_dx")]:
        x = next((e["ts"] for e in step if name in e["n"]), None)
        if x is not None:
            print(f"  {label:14s} at {(x - t0) / 1e3:6.2f} ms")
    dad = [e for e in step if "adamw" in e["n"] and e["s"] != main]
    if dad:
        print(f"  last adamw D   at {(dad[-1]['ts'] - t0) / 1e3:6.2f} ms")
    streams = collections.defaultdict(list)
    for e in step:
        streams[e["s"]].append(e)
    order = sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1]))
    for s, l in order[:12]:
        print(f"  stream {s}: {len(l):4d} kernels busy {sum(e['dur'] for e in l) / 1e3:5.2f} ms  span {(l[0]['ts'] - t0) / 1e3:5.2f}"
              f"..{(max(e['ts'] + e['dur'] for e in l) - t0) / 1e3:5.2f}")
    pts = []
    for e in step:
        pts += [(e["ts"], 1), (e["ts"] + e["dur"], -1)]
    pts.sort()
    act, last, hist = 0, t0, collections.defaultdict(float)
    for tt, d in pts:
        hist[act] += tt - last
        last = tt
        act += d
    print("  ms with k kernels in flight:", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
    for s, l in order[:top]:
        agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
        prev = None
        for e in l:
            a = agg[e["n"][:50]]
            a[0] += 1
            a[1] += e["dur"]
            a[2] += 0 if prev is None else max(0, e["ts"] - prev)
            prev = e["ts"] + e["dur"]
        print(f"  -- stream {s}")
        for k, (c, b, g) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2]))[:12]:
            print(f"     {c:4d} x busy {b:7.1f} us  gap-before {g:7.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/train_trace.json", int(sys.argv[2]) if len(sys.argv) > 2 else 3)
