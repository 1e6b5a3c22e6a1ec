"""Train-step parity (one iteration of model_wrapper._gan_training incl. lazy R1 + path length) against the
oracle restatement, and the data-parallel plumbing on world_size 2 (gloo, CPU)."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from oracle.make_golden import TINY_D, TINY_G, randomize
from tests.conftest import rel_err
from oracle.train_step import OracleTrainer

HP = None


def _hp():
    from multi_stylegan_b200 import config
    hp = dict(config.generation_hyperparameters)
    hp["lazy_discriminator_regularization"] = 2      # so the second iteration runs R1 and path length
    hp["lazy_generator_regularization"] = 2
    hp["p_mixed_noise"] = 1.0
    return hp


class FixedNoiseGenerator(torch.nn.Module):
    """Pins the per-layer noise maps and inject_index so oracle and product see the same draws."""

    def __init__(self, net, noise, inject):
        super().__init__()
        self.net, self.noise, self.inject = net, noise, inject
        self.latent_dimensions = net.latent_dimensions

    def parameters(self, recurse=True):
        return self.net.parameters(recurse)

    def forward(self, input, **kw):
        b = input[0].shape[0] if isinstance(input, list) else input.shape[0]
        return self.net(input, noise=[n[:b] for n in self.noise], inject_index=self.inject, **kw)


def build(dev, seed=0):
    import multi_stylegan_b200.multi_stylegan_generator as G_mod
    import multi_stylegan_b200.u_net_2d_discriminator as D_mod
    torch.manual_seed(seed)
    G = G_mod.Generator(TINY_G, compute_dead_branch=False)
    D = D_mod.Discriminator(TINY_D, no_rfp=True)
    randomize(G, 1), randomize(D, 2)
    return G.to(dev), D.to(dev)


def run_parity(dev, tol, n_iter=2, tf32_matmul=True):
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    hp = _hp()
    G, D = build("cpu")
    sd_g = {k: v.detach().clone() for k, v in G.state_dict().items()}
    sd_d = {k: v.detach().clone() for k, v in D.state_dict().items()}
    G, D = G.to(dev), D.to(dev)
    gen = torch.Generator().manual_seed(3)
    B = 2
    noise = [torch.randn(B, 1, 4, 4, generator=gen)] + \
            [torch.randn(B, 1, 2 ** (i // 2 + 3), 2 ** (i // 2 + 3), generator=gen) for i in range(6)]
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"])
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"])
    wrapped = FixedNoiseGenerator(G, [n.to(dev) for n in noise], 3)
    mw = ModelWrapper(wrapped, D, opt_g, opt_d, hyperparameters=hp, generator_ema=__import__("copy").deepcopy(G), device=dev,
                      allow_tf32_matmul=tf32_matmul)
    mw.generator_ema_ref = G
    oracle = OracleTrainer(sd_g, sd_d, (2e-3, 2e-5), 6e-3, hp["betas"], hp)
    # EMA must follow the real generator, not the wrapper
    import multi_stylegan_b200.misc as misc
    for it in range(n_iter):
        real = torch.rand(B, 2, 3, 32, 32, generator=gen)
        zs = [[torch.randn(B, 16, generator=gen), torch.randn(B, 16, generator=gen)] for _ in range(2)]
        z_pl = [torch.randn(1, 16, generator=gen), torch.randn(1, 16, generator=gen)]
        pl_noise = torch.randn(1, 2, 3, 32, 32, generator=gen)
        want = oracle.step(real, zs[0], zs[1], z_pl, noise, noise, [n[:1] for n in noise], 3, pl_noise)
        mw.generator_ema, keep = torch.nn.Identity(), mw.generator_ema     # EMA checked separately below
        import unittest.mock as um
        with um.patch.object(misc, "exponential_moving_average", lambda **kw: None):
            got = mw.train_step(real.to(dev), z_d=[z.to(dev) for z in zs[0]], z_g=[z.to(dev) for z in zs[1]],
                                z_pl=[z.to(dev) for z in z_pl], pl_noise=pl_noise.to(dev).requires_grad_(True))
        mw.generator_ema = keep
        misc.exponential_moving_average(model_ema=keep, model_train=G)
        assert set(got) == set(want), (sorted(got), sorted(want))
        for k in want:
            # the path-length penalty is (l - mean)^2 with mean = 0.01 l on the first lazy step: it doubles the
            # relative error of the path length l itself
            ktol = 3 * tol if k == "loss_path_length_regularization" else tol
            assert rel_err(got[k], want[k]) < ktol, (it, k, float(got[k]), float(want[k]))
    worst = 0.0
    for n, p in G.named_parameters():
        worst = max(worst, rel_err(p, oracle.sd_g[n]))
        assert rel_err(dict(keep.named_parameters())[n], oracle.ema[n]) < max(tol, 1e-5), n
    for n, p in D.named_parameters():
        worst = max(worst, rel_err(p, oracle.sd_d[n]))
    return worst


def test_train_step_host_logic_matches_oracle(oracle_backend):
    assert run_parity("cpu", 2e-4) < 2e-3


@pytest.mark.gpu
def test_train_step_cuda_core_engine_matches_oracle(built_library):
    from multi_stylegan_b200 import _C, _lib
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_SIMT
    try:
        assert run_parity("cuda:0", 5e-4, tf32_matmul=False) < 5e-3      # exact fp32 everywhere
    finally:
        _C.conv_flags = old


@pytest.mark.gpu
def test_train_step_tensor_core_engine_matches_oracle(built_library):
    # losses within the TF32 tolerance; updated weights move by lr * sign-like Adam steps, so compare loosely
    assert run_parity("cuda:0", 2e-2) < 0.2


# ---- data parallel: world_size 2 over gloo ---------------------------------------------------------------
def _dp_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from tests import backend_oracle
    from multi_stylegan_b200 import _C, dist as mdist
    for name in backend_oracle.__all__:
        setattr(_C, name, getattr(backend_oracle, name))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G, D = build("cpu", seed=rank)             # different init per rank: broadcast must fix it
        mdist.broadcast_parameters([G, D])
        gen = torch.Generator().manual_seed(5)
        x = torch.rand(world * 2, 2, 3, 32, 32, generator=gen)
        shard = x[rank * 2:(rank + 1) * 2]
        s, p = D(shard)
        (torch.nn.functional.softplus(-s).mean() + torch.nn.functional.softplus(-p).mean()).backward()
        n = mdist.all_reduce_gradients(list(D.parameters()))
        r = torch.tensor([float(rank)])
        mdist.all_reduce_mean_(r)
        torch.save({"grads": {k: v.grad.clone() for k, v in D.named_parameters()}, "n": n, "r": r,
                    "w0": next(iter(D.parameters())).detach().clone()}, os.path.join(tmp, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo(tmp_path, oracle_backend):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_dp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r), weights_only=False) for r in range(world)]
    # single-process reference: mean over shards of the per-shard gradients (MinibatchStdDev is per shard,
    # exactly as under the reference's DataParallel, u_net_2d_discriminator.py:212-214)
    G, D = build("cpu", seed=0)
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(world * 2, 2, 3, 32, 32, generator=gen)
    acc = None
    for r in range(world):
        D.zero_grad()
        s, p = D(x[r * 2:(r + 1) * 2])
        (torch.nn.functional.softplus(-s).mean() + torch.nn.functional.softplus(-p).mean()).backward()
        g = {k: v.grad.clone() / world for k, v in D.named_parameters()}
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    for o in outs:
        assert o["n"] == sum(p.numel() for p in D.parameters())
        assert abs(float(o["r"]) - 0.5) < 1e-6
        assert torch.equal(o["w0"], next(iter(D.parameters())).detach())      # broadcast from rank 0
        for k in acc:
            assert rel_err(o["grads"][k], acc[k]) < 1e-5, k
    for k in acc:
        assert torch.equal(outs[0]["grads"][k], outs[1]["grads"][k])


def _dp_structure_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import random
    import torch.distributed as dist
    from tests import backend_oracle
    from multi_stylegan_b200 import _C, dist as mdist
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    for name in backend_oracle.__all__:
        setattr(_C, name, getattr(backend_oracle, name))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hp = _hp()
        G, D = build("cpu", seed=0)
        mdist.broadcast_parameters([G, D])
        opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"])
        opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"])
        mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device="cpu")
        mw.epochs, mw.epoch = 10, 9                  # CutMix probability 0.45, wrong-order fakes on
        random.seed(100 + rank)                      # the ranks' own host streams differ
        torch.manual_seed(100 + rank)
        keys = []
        torch.set_num_threads(4)
        for it in range(2):                          # the shared draws give: CutMix, then no CutMix
            out = mw.train_step(torch.rand(2, 2, 3, 32, 32))
            keys.append(sorted(out))
        torch.save({"keys": keys, "w": [p.detach().clone() for p in D.parameters()]}, os.path.join(tmp, "s%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_data_parallel_ranks_agree_on_step_structure_gloo(tmp_path, oracle_backend):
    """Late-epoch iterations on 2 ranks whose host random streams differ: the CutMix draw decides how many optimiser steps
    (= gradient all-reduces) an iteration has, so every rank must take the same decision — otherwise this test hangs in a
    mismatched collective.  Parameters stay identical across ranks."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_dp_structure_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a, b = [torch.load(os.path.join(str(tmp_path), "s%d.pt" % r), weights_only=False) for r in range(world)]
    assert a["keys"] == b["keys"]
    assert any("loss_cut_mix_augmentation" in k for k in a["keys"]) and any("loss_cut_mix_augmentation" not in k for k in a["keys"])
    for x, y in zip(a["w"], b["w"]):
        assert torch.equal(x, y)


def test_late_epoch_branches_host_logic(oracle_backend):
    """CutMix augmentation + consistency steps, wrong-order fakes and top-k filtering (model_wrapper.py:272-277,331-376,
    392-401; probability 0 in epoch 0, so the benchmark never runs them): one iteration with all of them switched on."""
    import random
    from multi_stylegan_b200 import loss
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    hp = _hp()
    G, D = build("cpu")
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"])
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"])
    mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device="cpu")
    mw.epochs, mw.epoch = 10, 9                 # wrong-order fakes on, CutMix probability 0.45
    mw.top_k = loss.TopK(starting_iteration=0, final_iteration=1)
    random.seed(0)
    torch.manual_seed(0)
    before = {n: p.detach().clone() for n, p in D.named_parameters()}
    seen = set()
    for it in range(6):
        out = mw.train_step(torch.rand(4, 2, 3, 32, 32))
        assert all(torch.isfinite(v).all() for v in out.values())
        seen |= set(out)
    assert {"loss_cut_mix_augmentation", "loss_cut_mix_regularization", "loss_generator",
            "loss_discriminator_real"} <= seen
    assert any(not torch.equal(before[n], p.detach()) for n, p in D.named_parameters())


def test_device_inject_index_matches_host_index(oracle_backend):
    """Generator._latent with the crossover index as a device scalar (CUDA-graph form) == the reference's int form
    (multi_stylegan_generator.py:162-169), including the 'no mixing' encoding inject_index = n_latent."""
    G, _ = build("cpu")
    z = [torch.randn(3, G.latent_dimensions), torch.randn(3, G.latent_dimensions)]
    for idx in range(1, G.n_latent - 1):
        assert torch.equal(G._latent(z, False, idx), G._latent(z, False, torch.tensor(idx)))
    assert torch.equal(G._latent(z[0], False, None), G._latent(z, False, torch.tensor(G.n_latent)))


@pytest.mark.gpu
@pytest.mark.parametrize("segmented", [False, True])
def test_cuda_graph_replay_matches_eager(built_library, segmented):
    """ModelWrapper(cuda_graphs=True): captured + replayed iterations (plain and lazy variants; as one graph, and as the
    multi-rank program of graph segments cut at every gradient all-reduce) give the same losses and the same parameters
    as the eager execution of the same iterations with the same host / device random streams."""
    import random
    import numpy as np
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    dev = torch.device("cuda:0")
    hp = _hp()
    hp["p_mixed_noise"] = 0.6
    runs = []
    for graphed in (False, True):
        G, D = build(dev)
        opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"], fused=True, capturable=True)
        opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"], fused=True, capturable=True)
        mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev, cuda_graphs=True)
        mw._always_break = segmented            # the multi-rank form: graph segments around the (here no-op) all-reduces
        random.seed(7), np.random.seed(7), torch.manual_seed(7)
        gen = torch.Generator().manual_seed(11)
        hist = []
        for it in range(6):
            if not graphed:
                mw._graphs.clear()              # every iteration is a "first occurrence": eager execution
            real = torch.rand(4, 2, 3, 32, 32, generator=gen).to(dev)
            hist.append({k: v.detach().clone() for k, v in mw.train_step(real).items()})
        torch.cuda.synchronize()
        assert mw.graph_replays == (4 if graphed else 0)
        if graphed:
            assert mw.graph_launches > 0
            n_graphs = {k[2:4]: sum(isinstance(i, torch.cuda.CUDAGraph) for i in st.program) for k, st in mw._graphs.items()}
            # plain: D fwd/bwd | G fwd (overlaps the D all-reduce) | D update .. G bwd | G update + EMA; lazy: one break per
            # optimiser step (the lazy regularisers need the updated networks, nothing to overlap)
            assert n_graphs == ({(False, False): 4, (True, True): 5} if segmented else {(False, False): 1, (True, True): 1})
        runs.append((hist, [p.detach().clone() for p in list(G.parameters()) + list(D.parameters())],
                     mw.path_length_regularization.mean_path_length.clone()))
    (h0, p0, m0), (h1, p1, m1) = runs
    for a, b in zip(h0, h1):
        assert set(a) == set(b)
        for k in a:
            assert torch.allclose(a[k], b[k], rtol=2e-3, atol=1e-5), (k, a[k], b[k])
    assert "loss_path_length_regularization" in h1[-1] and "loss_discriminator_regularization" in h1[-1]
    assert torch.allclose(m0, m1, rtol=2e-3)
    for a, b in zip(p0, p1):
        assert rel_err(b, a) < 2e-3


def _graph_vs_eager(make_wrapper, schedule, n_iter=6):
    """Run the same iterations eagerly and with CUDA-graph replay (same host / device random streams) and return both
    histories.  `schedule(mw, it)` may change wrapper state before iteration `it` and returns extra train_step kwargs."""
    import random
    import numpy as np
    dev = torch.device("cuda:0")
    runs = []
    for graphed in (False, True):
        mw, nets = make_wrapper(dev)
        random.seed(7), np.random.seed(7), torch.manual_seed(7)
        gen = torch.Generator().manual_seed(11)
        hist = []
        for it in range(n_iter):
            kw = schedule(mw, it, gen)
            if not graphed:
                mw._graphs.clear()              # every iteration is a "first occurrence": eager execution
            real = torch.rand(4, 2, 3, 32, 32, generator=gen).to(dev)
            hist.append({k: v.detach().clone() for k, v in mw.train_step(real, **kw).items()})
        torch.cuda.synchronize()
        runs.append((mw, hist, [p.detach().clone() for net in nets for p in net.parameters()],
                     mw.path_length_regularization.mean_path_length.clone()))
    return runs


def _assert_same_runs(runs, rtol=2e-3, atol=1e-5, first=None, param_tol=None):
    """`first`: compare the loss values of that many leading iterations only (keys are compared for all)."""
    (_, h0, p0, m0), (_, h1, p1, m1) = runs
    for i, (a, b) in enumerate(zip(h0, h1)):
        assert set(a) == set(b), (i, sorted(a), sorted(b))
        for k in a:
            assert torch.isfinite(a[k]).all() and torch.isfinite(b[k]).all()
            if first is None or i < first:
                assert torch.allclose(a[k], b[k], rtol=rtol, atol=atol), (i, k, a[k], b[k])
    if first is None:
        assert torch.allclose(m0, m1, rtol=rtol)
    for a, b in zip(p0, p1):
        assert rel_err(b, a) < (rtol if param_tol is None else param_tol)


@pytest.mark.gpu
def test_cuda_graph_keys_cover_trap_weights_and_running_mean_survives_eager_iterations(built_library):
    """(1) The trap weighting of the pixel-wise losses switches on at trap_weight * epochs (model_wrapper.py:262-263):
    a graph captured before the switch must not be replayed after it.  (2) An eager lazy iteration (here: fixed
    latents) between graphed ones moves the path-length running mean; the next replay must start from it."""
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    hp = _hp()
    hp["p_mixed_noise"] = 0.0
    trap = torch.rand(32, 32) + 0.5

    def make(dev):
        G, D = build(dev)
        opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"], fused=True, capturable=True)
        opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"], fused=True, capturable=True)
        mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev, cuda_graphs=True,
                          trap_weights_map=trap.clone())            # a CPU map: moved to the device before capture
        mw.epochs, mw.epoch = 4, 0
        return mw, (G, D)

    def schedule(mw, it, gen):
        if it == 10:
            mw.epoch = 1                       # trap_weight (0.25) * epochs (4) <= epoch: weighting on from here
        if it == 7:                            # a lazy iteration (every 2nd) forced to run eagerly in both runs, after
            return dict(z_pl=torch.randn(2, 16, generator=gen).to("cuda:0"))   # its variant was captured (it = 3)
        return {}
    runs = _graph_vs_eager(make, schedule, n_iter=14)
    _assert_same_runs(runs)
    mw = runs[1][0]
    assert mw.graph_replays >= 4
    assert {k[5] for k in mw._graphs} == {False, True}          # separate programs before / after the switch


@pytest.mark.gpu
def test_cuda_graph_replay_with_ada_matches_eager(built_library):
    """The ADA wrapper is graph-capturable: its per-call draws stay on the host (reference order) and reach the captured
    kernels through plan tensors refreshed before every replay; p moves at the same iterations as in eager execution."""
    from multi_stylegan_b200.adaptive_discriminator_augmentation import AdaptiveDiscriminatorAugmentation
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    hp = _hp()
    hp["p_mixed_noise"] = 0.5

    def make(dev):
        G, D = build(dev)
        opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"], fused=True, capturable=True)
        opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"], fused=True, capturable=True)
        ada = AdaptiveDiscriminatorAugmentation(D, r_update=2, p_step=0.1)
        ada.p = 0.5
        mw = ModelWrapper(G, ada, opt_g, opt_d, hyperparameters=hp, device=dev, cuda_graphs=True)
        mw._d_params = lambda: list(D.parameters())
        return mw, (G, D)
    runs = _graph_vs_eager(make, lambda mw, it, gen: {}, n_iter=8)
    # The adjoint of the warp scatters with fp32 atomics whose summation order differs from launch to launch (1e-4 relative
    # on the first replayed iteration), and this 32x32 toy GAN at lr 6e-3 amplifies a perturbation about tenfold per
    # iteration (tools/ada_graph_probe.py prints both trajectories: 1e-4, 2e-3, 1e-2, 5e-2 ... from the first replay on; the
    # spread varies from run to run).  The two eager iterations and the first replayed one are compared value by value;
    # after that only finiteness, the loss keys, the early controller state and a coarse bound on the parameters (Adam
    # moves an element by at most ~lr per iteration) are checked.
    _assert_same_runs(runs, rtol=3e-2, atol=2e-3, first=3, param_tol=0.5)
    eager, graphed = runs[0][0], runs[1][0]
    assert graphed.graph_replays >= 4
    rh_e, rh_g = eager.discriminator.r_history, graphed.discriminator.r_history
    assert len(rh_e) == len(rh_g) and len(rh_g) >= 6
    assert rh_e[:3] == pytest.approx(rh_g[:3])
    assert abs(eager.discriminator.p - graphed.discriminator.p) <= 0.2 + 1e-6        # at most two controller steps apart


def run_late_epoch_parity(dev, tol):
    """One iteration with wrong-order fakes, trap-weighted pixel-wise losses, both CutMix steps and top-k filtering
    (model_wrapper.py:262-277,331-376,392-405) against the oracle trainer with the same draws injected."""
    import unittest.mock as um
    from multi_stylegan_b200 import loss, misc
    from multi_stylegan_b200 import model_wrapper as MW
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    hp = _hp()
    hp["lazy_discriminator_regularization"] = 10 ** 6
    hp["lazy_generator_regularization"] = 10 ** 6
    G, D = build("cpu")
    sd_g = {k: v.detach().clone() for k, v in G.state_dict().items()}
    sd_d = {k: v.detach().clone() for k, v in D.state_dict().items()}
    G, D = G.to(dev), D.to(dev)
    gen = torch.Generator().manual_seed(5)
    B = 4
    noise = [torch.randn(B, 1, 4, 4, generator=gen)] + \
            [torch.randn(B, 1, 2 ** (i // 2 + 3), 2 ** (i // 2 + 3), generator=gen) for i in range(6)]
    trap = torch.rand(32, 32, generator=gen) + 0.5
    perm = torch.tensor([2, 0, 2])                                  # with replacement, like misc.random_permutation
    maps = []
    for (ch, cw, low, invert) in ((9, 20, True, False), (21, 7, False, True)):
        m = torch.zeros(1, 1, 1, 32, 32)
        if low:
            m[..., ch:, cw:] = 1.0
        else:
            m[..., :ch, :cw] = 1.0
        maps.append(1.0 - m if invert else m)
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"])
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"])
    wrapped = FixedNoiseGenerator(G, [n.to(dev) for n in noise], 3)
    mw = ModelWrapper(wrapped, D, opt_g, opt_d, hyperparameters=hp, generator_ema=__import__("copy").deepcopy(G), device=dev,
                      trap_weights_map=trap.to(dev))
    mw.epochs, mw.epoch = 10, 9                                     # wrong order on (>= 0.5 * epochs), trap on, CutMix p = 0.45
    mw.top_k = loss.TopK(starting_iteration=0, final_iteration=1)   # v = 0.5
    oracle = OracleTrainer(sd_g, sd_d, (2e-3, 2e-5), 6e-3, hp["betas"], hp)
    real = torch.rand(B, 2, 3, 32, 32, generator=gen)
    zs = [[torch.randn(B, 16, generator=gen), torch.randn(B, 16, generator=gen)] for _ in range(2)]
    n_wrong = max(1, int(hp["batch_factor_wrong_order"] * B))
    want = oracle.step(real, zs[0], zs[1], None, noise, noise, None, 3, None,
                       late=dict(perm=perm, n_wrong=n_wrong, cut_mix=tuple(maps), top_k_v=0.5, trap=trap))
    map_iter = iter([m.to(dev) for m in maps])
    with um.patch.object(misc, "random_permutation", lambda n: perm), \
            um.patch.object(MW.random, "random", lambda: 0.0), \
            um.patch("multi_stylegan_b200.u_net_2d_discriminator._generate_binary_cut_mix_map",
                     lambda height, width, device="cpu": next(map_iter)), \
            um.patch.object(misc, "exponential_moving_average", lambda **kw: None):
        # the wrong-order reals make the fake batch larger than the real one: the generator's noise maps must cover it
        got = mw.train_step(real.to(dev), z_d=[z.to(dev) for z in zs[0]], z_g=[z.to(dev) for z in zs[1]])
    assert set(got) == set(want), (sorted(got), sorted(want))
    assert {"loss_cut_mix_augmentation", "loss_cut_mix_regularization"} <= set(got)
    for k in want:
        assert rel_err(got[k], want[k]) < tol, (k, float(got[k]), float(want[k]))
    worst = 0.0
    for n, p in D.named_parameters():
        worst = max(worst, rel_err(p, oracle.sd_d[n]))
    for n, p in G.named_parameters():
        worst = max(worst, rel_err(p, oracle.sd_g[n]))
    return worst


def test_late_epoch_branches_match_oracle_host_logic(oracle_backend):
    assert run_late_epoch_parity("cpu", 2e-4) < 2e-3


@pytest.mark.gpu
def test_late_epoch_branches_match_oracle_cuda_core_engine(built_library):
    from multi_stylegan_b200 import _C, _lib
    old = _C.conv_flags
    _C.conv_flags = _lib.CONV_FORCE_SIMT
    try:
        # losses to 1e-3; the updated weights only loosely: this iteration takes five optimiser steps, and Adam's first
        # steps are +-lr whatever the gradient's size, so a parameter whose gradient is rounding noise moves by several lr
        assert run_late_epoch_parity("cuda:0", 1e-3) < 0.1
    finally:
        _C.conv_flags = old


@pytest.mark.gpu
def test_late_epoch_branches_match_oracle_tensor_core_engine(built_library):
    assert run_late_epoch_parity("cuda:0", 2e-2) < 0.2
