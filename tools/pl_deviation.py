"""How far does TF32 move the path-length (double-backward) parameter gradients of the tiny fixture generator?
(a) this package on tcgen05, (b) stock PyTorch/cuDNN with TF32 on and off, all against the fp32 CPU fixture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.conftest import load_golden
from oracle import model as om
import multi_stylegan_b200.multi_stylegan_generator as G_mod
from multi_stylegan_b200 import loss


def l2(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def mx(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-3)).item()


def report(tag, grads, ref, pl_grad, ref_pl):
    names = [n for n in sorted(ref) if n in grads and grads[n] is not None]
    rows = sorted(((mx(grads[n], ref[n]), l2(grads[n], ref[n]), n) for n in names), reverse=True)
    allg = torch.cat([grads[n].flatten().double().cpu() for n in names])
    allr = torch.cat([ref[n].flatten().double() for n in names])
    print("=== %s: pl_grad l2 %.5f | global l2 over %d tensors %.4f | worst max-metric %.4f | worst l2 %.4f" % (
        tag, l2(pl_grad, ref_pl), len(names), l2(allg, allr), rows[0][0], max(r[1] for r in rows)))
    for r in rows[:6]:
        print("    max %.4f  l2 %.4f  %s (ref absmax %.3e)" % (r[0], r[1], r[2], ref[r[2]].abs().max()))


dev = "cuda:0"
g = load_golden("generator.pt")
noise = [t.to(dev) for t in g["noise"]]
ref = g["pl_param_grads"]

net = G_mod.Generator(g["config"], compute_dead_branch=False)
net.load_state_dict(g["state_dict"]); net.to(dev)
latent = net._latent(g["z1"].to(dev), False, None)
image = net(latent, noise=noise, input_is_latent=True)
pl_grad = torch.autograd.grad((image * g["pl_noise"].to(dev)).sum(), latent, create_graph=True)[0]
penalty, pl = loss.PathLengthRegularization()(pl_grad)
net.zero_grad(); penalty.backward()
report("this package (tcgen05 tf32)", {n: p.grad for n, p in net.named_parameters()}, ref, pl_grad, g["pl_grad"])

for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    sd = {k: v.to(dev).clone().requires_grad_(v.dtype.is_floating_point) for k, v in g["state_dict"].items()}
    lat = om.generator_latent(sd, g["z1"].to(dev), None, latent.shape[1])
    pg = om.path_length_grads(sd, lat, noise, g["pl_noise"].to(dev))
    pen, _, _ = om.path_length_penalty(pg, torch.zeros(1, device=dev))
    names = [n for n in sorted(ref)]
    grads = torch.autograd.grad(pen, [sd[n] for n in names], allow_unused=True)
    report("stock torch GPU tf32=%s" % tf32, dict(zip(names, grads)), ref, pg, g["pl_grad"])
