"""Stage-by-stage comparison of the CUDA BMGAN generator with the CPU oracle (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import petsyn
from oracle import bmgan as OB

SMALL = dict(input_conv_channel=64, output_conv_channel=64, down_channels=[64, 128, 128, 128], middle_channels=[128],
             up_channels=[128, 128, 128, 128, 64])
torch.manual_seed(777)
gen = petsyn.dense_unet_generator(**SMALL).train()
og = OB.DenseUnetGenerator(**SMALL).train()
og.load_state_dict(gen.state_dict())
g = torch.Generator().manual_seed(1)
shape = (2, 64, 96, 64)
t1 = torch.rand(shape[0], 1, *shape[1:], generator=g); z = torch.randn(shape[0], 8, generator=g)
feats = {}
def hook(name):
    def f(m, i, o): feats[name] = o.detach()
    return f
og.input_layer.register_forward_hook(hook("F0"))
og.input_layer[2].register_forward_hook(hook("in.a0"))
for i, b in enumerate(og.down_layers):
    b.register_forward_hook(hook(f"down{i}"))
    b[0].register_forward_hook(hook(f"down{i}.cat0"))
    b[3].register_forward_hook(hook(f"down{i}.y0"))
    b[4].register_forward_hook(hook(f"down{i}.cat1"))
    b[7].register_forward_hook(hook(f"down{i}.y1"))
og.middle_layers.register_forward_hook(hook("mid"))
for j, b in enumerate(og.up_layers):
    b.register_forward_hook(hook(f"up{j}"))
    b[0].register_forward_hook(hook(f"up{j}.cat0"))
    b[7].register_forward_hook(hook(f"up{j}.y1"))
yo = og(t1, z)
gen = gen.cuda()
with torch.no_grad():
    y = gen(t1.cuda(), z.cuda())
eng = list(gen._engines.values())[0]
def cmp(name, buf, sl=None):
    ref = feats[name]
    n, c, d, h, w = ref.shape
    t = eng.bufs[buf].t.float().view(n, d, h, w, -1).permute(0, 4, 1, 2, 3).cpu()
    if sl is not None: t = t[:, sl[0]:sl[1]]
    e = (t - ref).abs()
    print(f"{name:14s} vs {buf:12s} max {e.max().item():.4f} mean {e.mean().item():.5f} refmax {ref.abs().max().item():.3f}")
cmp("in.a0", "in.a0")
cmp("F0", "down0.cat0", (0, 64))
for i in range(4):
    cmp(f"down{i}.cat0", f"down{i}.cat0")
    cmp(f"down{i}.cat1", f"down{i}.cat1")
    cmp(f"down{i}.y1", f"down{i}.y1")
cmp("down3", "mid.cat0", (0, 128))
cmp("mid", "up0.cat0", (0, 128))
for j in range(5):
    cmp(f"up{j}.cat0", f"up{j}.cat0")
    cmp(f"up{j}.y1", f"up{j}.y1")
cmp("up4", "out.in")
e = (y.cpu() - yo.detach()).abs(); print("output", e.max().item(), e.mean().item())
