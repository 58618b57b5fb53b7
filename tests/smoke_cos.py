"""Which generator gradients are furthest from the CPU fp32 oracle at smoke()'s size (tool): python tests/smoke_cos.py [batch]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hifigan_b200 as H
from hifigan_b200.train import TrainStep
from oracle import hifigan_oracle as O
from oracle import train_oracle as TO

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
h = H.AttrDict(O.config("v1"))
torch.manual_seed(1234)
nets = (H.Generator(h), H.MultiPeriodDiscriminator(), H.MultiScaleDiscriminator())
sds = [TO.leaf_params({k: v.detach().clone() for k, v in m.state_dict().items()}) for m in nets]
ya = O.synthetic_audio(b, 8192, seed=4)
mel_in = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, 8000)
mel_loss = O.mel_spectrogram(ya, 1024, 80, 22050, 256, 1024, 0, None)
ref_losses, gg, gd_p, gd_s, _, _ = TO.train_step(*sds, h, mel_in, ya.unsqueeze(1), mel_loss)
ts = TrainStep(*nets, h, "cuda")
yc = ya.cuda()
out = ts.step(H.mel_spectrogram(yc, 1024, 80, 22050, 256, 1024, 0, 8000), yc.unsqueeze(1),
              H.mel_spectrogram(yc, 1024, 80, 22050, 256, 1024, 0, None), update=False)
rows = []
for k, p in nets[0].named_parameters():
    g, r = p.grad.cpu().flatten().double(), gg[k].flatten().double()
    cos = torch.nn.functional.cosine_similarity(g, r, dim=0).item()
    rows.append((cos, k, tuple(p.shape), (g - r).norm().item() / r.norm().item(), r.norm().item()))
rows.sort()
print(f"batch {b}: worst generator gradients vs the CPU fp32 oracle")
for cos, k, shp, rel, nrm in rows[:4]:
    print(f"  cos {cos:.5f}  rel-L2 {rel:.3e}  |ref| {nrm:.3e}  {k} {shp}")
