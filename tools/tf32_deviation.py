"""How far does TF32 move module-level gradients?  Compare (a) this package on tcgen05 and (b) stock
PyTorch/cuDNN with TF32 enabled (the reference's own GPU path) against the fp32 CPU fixture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.conftest import load_golden
from oracle import model as om
import multi_stylegan_b200.u_net_2d_discriminator as D_mod
import multi_stylegan_b200.multi_stylegan_generator as G_mod

def l2(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
def mx(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()

dev = "cuda:0"
g = load_golden("discriminator.pt")
net = D_mod.Discriminator(g["config"], no_rfp=True); net.load_state_dict(g["state_dict"]); net.to(dev)
s, p = net(g["x"].to(dev))
((s * g["ds"].to(dev)).sum() + (p * g["dp"].to(dev)).sum()).backward()
mine = {n: q.grad for n, q in net.named_parameters()}
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    sd = {k: v.to(dev).clone().requires_grad_(v.dtype.is_floating_point) for k, v in g["state_dict"].items()}
    s2, p2 = om.discriminator_forward(sd, g["x"].to(dev))
    names = sorted(g["grads"])
    grads = torch.autograd.grad((s2 * g["ds"].to(dev)).sum() + (p2 * g["dp"].to(dev)).sum(), [sd[n] for n in names])
    stock = dict(zip(names, grads))
    print("=== D: stock torch on GPU, tf32 =", tf32, " out max err", mx(s2, g["scalar"]), mx(p2, g["pixel"]))
    worst = sorted(((l2(stock[n], g["grads"][n]), mx(stock[n], g["grads"][n]), n) for n in names), reverse=True)[:5]
    for w in worst: print("   stock  l2 %.4f max %.4f %s" % w)
print("=== D: this package (tcgen05): out max err", mx(s, g["scalar"]), mx(p, g["pixel"]))
worst = sorted(((l2(mine[n], g["grads"][n]), mx(mine[n], g["grads"][n]), n) for n in mine), reverse=True)[:8]
for w in worst: print("   mine   l2 %.4f max %.4f %s" % w)
allm = torch.cat([mine[n].flatten().cpu() for n in sorted(mine)]); allr = torch.cat([g["grads"][n].flatten() for n in sorted(mine)])
print("   mine global l2", l2(allm, allr))
