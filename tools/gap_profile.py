"""Where is the GPU idle inside a train step?  Gaps between consecutive kernels on the stream, attributed to the kernel
that started late (torch.profiler / CUPTI timeline)."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
from multi_stylegan_b200 import config
import multi_stylegan_b200.multi_stylegan_generator as G_mod
import multi_stylegan_b200.u_net_2d_discriminator as D_mod
from multi_stylegan_b200.model_wrapper import ModelWrapper

dev = torch.device("cuda:0")
torch.manual_seed(0)
G = G_mod.Generator(config.multi_style_gan_generator_config, compute_dead_branch=False).to(dev)
D = D_mod.Discriminator(config.u_net_2d_discriminator_config, no_rfp=True).to(dev)
hp = dict(config.generation_hyperparameters)
opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-4, lr_style=2e-6), betas=hp["betas"], fused=True)
opt_d = torch.optim.Adam(D.parameters(), lr=6e-4, betas=hp["betas"], fused=True)
mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev)
real = torch.rand(8, 2, 3, 256, 256, device=dev)
for i in range(4):
    mw.iteration = 0
    mw.train_step(real)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        mw.iteration = 0
        mw.train_step(real)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA" and e.time_range is not None]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
gaps = collections.defaultdict(lambda: [0, 0.0])
hist = collections.Counter()
end = ev[0].time_range.end
for a, b in zip(ev[:-1], ev[1:]):
    end = max(end, a.time_range.end)
    g = b.time_range.start - end
    if g > 0:
        gaps[b.name[:70]][0] += 1
        gaps[b.name[:70]][1] += g
        hist[min(int(g // 2) * 2, 40)] += 1
print("kernels %d  span %.2f ms  busy %.2f ms  idle %.2f ms (2 steps)" % (len(ev), span / 1e3, busy / 1e3, (span - busy) / 1e3))
print("gap histogram (us bucket: count):", sorted(hist.items()))
for k, (n, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%8.1f us total  %5d gaps  avg %5.1f us  before %s" % (t, n, t / n, k))
