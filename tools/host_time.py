"""Host-side issue time of one train step vs device time (is the step CPU-launch-bound?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
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
for i in range(3):
    mw.iteration = 0
    mw.train_step(real)
torch.cuda.synchronize()
N = 6
host = []
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(N):
    mw.iteration = 0
    t0 = time.perf_counter()
    mw.train_step(real)
    host.append(time.perf_counter() - t0)
b.record()
torch.cuda.synchronize()
print("device ms/step %.2f   host issue ms/step: %s" % (a.elapsed_time(b) / N, ["%.1f" % (h * 1e3) for h in host]))
# per-phase host time of one step with syncs between phases (how much GPU work each phase queues)
import cProfile, pstats, io
pr = cProfile.Profile()
pr.enable()
mw.iteration = 0
mw.train_step(real)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
