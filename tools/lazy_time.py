"""Wall time of plain vs lazy (R1 + path length) iterations, with caching-allocator statistics."""
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
seq = [15, 0, 0, 0, 15, 0, 0, 15, 0, 15]
for it0 in seq:
    mw.iteration = it0
    st = torch.cuda.memory_stats()
    a0, f0 = st.get("num_device_alloc", 0), st.get("num_device_free", 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mw.train_step(real)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st = torch.cuda.memory_stats()
    print("%s  host %.1f ms  total %.1f ms  cudaMalloc +%d cudaFree +%d  reserved %.1f GB  peak alloc %.1f GB  retries %d" % (
        "lazy " if it0 == 15 else "plain", (t1 - t0) * 1e3, (t2 - t0) * 1e3, st.get("num_device_alloc", 0) - a0,
        st.get("num_device_free", 0) - f0, st["reserved_bytes.all.current"] / 2**30,
        st["allocated_bytes.all.peak"] / 2**30, st.get("num_alloc_retries", 0)))
