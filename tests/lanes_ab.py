"""A/B of the MRF-lane schedule (hifi-gan_b200/models.py `_GeneratorEngine._launch`): branches of a stage one after
the other on all SMs vs side by side on SM subsets.  Checks bit-identity of the waveform and prints ms/step.

    python tests/lanes_ab.py [batch] [frames] [version]        (on a B200; results appended to gpurun_out/lanes_ab.txt)
"""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
H = importlib.import_module("hifi-gan_b200")


def timed(eng, x, steps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        y = eng.forward(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, y


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    ver = sys.argv[3] if len(sys.argv) > 3 else "v1"
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    G = H.Generator(importlib.import_module("hifi-gan_b200.configs").load_config(ver)).to(dev).eval()
    G.remove_weight_norm()
    x = torch.randn(batch, 80, frames, device=dev)
    eng = G._engine(dev)
    out = {"batch": batch, "frames": frames, "version": ver}
    with torch.no_grad():
        os.environ["HG_MRF_LANES"] = "0"
        for _ in range(2):
            eng.forward(x)
        ms0, y0 = timed(eng, x)
        y0 = y0.clone()
        out["serial_ms"] = ms0
        os.environ["HG_MRF_LANES"] = "1"
        for k in range(4):
            eng.forward(x)
            st = eng.ws[(batch, frames)]["_lanes"]
            print("call", k, "split", st["split"], flush=True)
        ms1, y1 = timed(eng, x)
        out["lanes_ms"] = ms1
        out["split"] = st["split"]
        out["bit_identical"] = bool(torch.equal(y0, y1))
        for forced in sys.argv[4:]:
            os.environ["HG_MRF_LANES"] = forced
            eng.ws[(batch, frames)].pop("_lanes", None)
            eng.forward(x)
            ms2, y2 = timed(eng, x)
            out[f"forced_{forced}_ms"] = ms2
            out[f"forced_{forced}_equal"] = bool(torch.equal(y0, y2))
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/lanes_ab.txt", "a") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
