"""GPU, >= 2 devices: data-parallel parity over NCCL.

Two ranks x b samples, gradients all-reduced over NCCL/NVLink by ModelWrapper, against ONE GPU x 2b samples with
MinibatchStdDev evaluated per b-sized group (which is what each DataParallel replica of the reference computes,
u_net_2d_discriminator.py:212-214).  The single-GPU side runs the two shards as two lock-stepped trainers in one process
whose "collective" is an in-process mean — the same arithmetic as one 2b batch with per-group statistics, and
independent of NCCL.  Two iterations (plain + lazy R1 / path length), fixed latents and noise maps, real kernels."""
import os
import threading

import pytest
import torch
import torch.multiprocessing as mp

from tests.conftest import rel_err
from tests.test_train_step import FixedNoiseGenerator, _hp, build

pytestmark = pytest.mark.gpu
B, WORLD, ITERS = 2, 2, 2


def _inputs(rank):
    gen = torch.Generator().manual_seed(1000 + rank)
    noise = [torch.randn(B, 1, 4, 4, generator=gen)] + \
            [torch.randn(B, 1, 2 ** (i // 2 + 3), 2 ** (i // 2 + 3), generator=gen) for i in range(6)]
    steps = []
    for _ in range(ITERS):
        steps.append(dict(real=torch.rand(B, 2, 3, 32, 32, generator=gen),
                          z_d=[torch.randn(B, 16, generator=gen), torch.randn(B, 16, generator=gen)],
                          z_g=[torch.randn(B, 16, generator=gen), torch.randn(B, 16, generator=gen)],
                          z_pl=[torch.randn(1, 16, generator=gen), torch.randn(1, 16, generator=gen)],
                          pl_noise=torch.randn(1, 2, 3, 32, 32, generator=gen)))
    return noise, steps


def _train(rank, dev, process_group=None, nets=None):
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    hp = _hp()
    G, D = build(dev, seed=0) if nets is None else nets
    noise, steps = _inputs(rank)
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"])
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"])
    wrapped = FixedNoiseGenerator(G, [n.to(dev) for n in noise], 3)
    mw = ModelWrapper(wrapped, D, opt_g, opt_d, hyperparameters=hp, generator_ema=__import__("copy").deepcopy(wrapped),
                      device=dev, process_group=process_group)
    losses = []
    for s in steps:
        out = mw.train_step(s["real"].to(dev), z_d=[z.to(dev) for z in s["z_d"]], z_g=[z.to(dev) for z in s["z_g"]],
                            z_pl=[z.to(dev) for z in s["z_pl"]], pl_noise=s["pl_noise"].to(dev).requires_grad_(True))
        losses.append({k: v.detach().cpu() for k, v in out.items()})
    torch.cuda.synchronize(dev)
    return {"losses": losses, "g": [p.detach().cpu() for p in G.parameters()], "d": [p.detach().cpu() for p in D.parameters()],
            "pl_mean": mw.path_length_regularization.mean_path_length.detach().cpu()}


def _nccl_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        torch.save(_train(rank, torch.device("cuda", rank)), os.path.join(tmp, "nccl%d.pt" % rank))
    finally:
        dist.destroy_process_group()


class _Lockstep:
    """In-process stand-in for the collective: every participant contributes its flat buffer, all receive the mean."""

    def __init__(self, n):
        self.n, self.barrier, self.slots, self.local = n, threading.Barrier(n), [None] * n, threading.local()

    def all_reduce_tensors(self, tensors, group=None):
        if not tensors:
            return 0
        flat = torch._utils._flatten_dense_tensors(tensors)
        torch.cuda.synchronize()
        self.slots[self.local.rank] = flat
        self.barrier.wait()
        total = sum(self.slots[1:], self.slots[0].clone()) / self.n
        self.barrier.wait()
        for t, f in zip(tensors, torch._utils._unflatten_dense_tensors(total, tensors)):
            t.copy_(f)
        return flat.numel()


@pytest.mark.timeout(900)
def test_two_ranks_over_nccl_equal_one_gpu_with_per_group_statistics(built_library, tmp_path):
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    import unittest.mock as um
    from multi_stylegan_b200 import dist as mdist
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_nccl_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    nccl = [torch.load(os.path.join(str(tmp_path), "nccl%d.pt" % r), weights_only=False) for r in range(WORLD)]

    ls = _Lockstep(WORLD)
    results, errors = [None] * WORLD, []
    # the replicas are built one after the other in this thread: torch's global CPU generator (parameter init) is shared
    # by all threads of a process
    replicas = [build(torch.device("cuda", 0), seed=0) for _ in range(WORLD)]

    def run(rank):
        try:
            ls.local.rank = rank
            results[rank] = _train(rank, torch.device("cuda", 0), nets=replicas[rank])
        except BaseException as exc:            # a dead participant must not leave the other one in the barrier
            errors.append(exc)
            ls.barrier.abort()
    class _Pending:
        def __init__(self, tensors):
            self.tensors = tensors

        def wait(self):
            return ls.all_reduce_tensors(self.tensors)
    with um.patch.object(mdist, "world_size", lambda group=None: WORLD), \
            um.patch.object(mdist, "all_reduce_tensors", ls.all_reduce_tensors), \
            um.patch.object(mdist, "all_reduce_tensors_begin", lambda tensors, group=None: _Pending(tensors) if tensors else None):
        threads = [threading.Thread(target=run, args=(r,)) for r in range(WORLD)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    assert not errors, errors
    for r in range(WORLD):
        # replicas stay identical across ranks, and equal the single-GPU emulation of the 2b batch
        for a, b in zip(nccl[r]["g"] + nccl[r]["d"], nccl[0]["g"] + nccl[0]["d"]):
            assert torch.equal(a, b)
        # Adam's first steps move every element by about +-lr whatever the gradient's magnitude, so an element whose
        # gradient is near zero can take the opposite step when the summation order differs (NCCL ring vs the emulation's
        # in-process mean; TF32 convolutions on both sides): bounded by 2 * lr per iteration, and rare.
        worst, differing, total = 0.0, 0, 0
        for a, b in zip(nccl[r]["g"] + nccl[r]["d"], results[r]["g"] + results[r]["d"]):
            diff = (a - b).abs()
            worst = max(worst, float(diff.max()))
            differing += int((diff > 1e-4).sum())
            total += diff.numel()
        assert worst <= 2.2 * 6e-3 * ITERS, worst
        assert differing / total < 0.02, (differing, total)
        assert rel_err(nccl[r]["pl_mean"], results[r]["pl_mean"]) < 1e-4
        for la, lb in zip(nccl[r]["losses"], results[r]["losses"]):
            assert set(la) == set(lb)
            for k in la:
                assert rel_err(la[k], lb[k]) < 1e-3, (k, la[k], lb[k])
    assert "loss_path_length_regularization" in nccl[0]["losses"][-1]


def _graph_worker(rank, world, port, tmp, modes):
    """Per rank: the same six iterations (plain + lazy variants) issued eagerly, replayed as ONE graph per variant with the
    NCCL all-reduces captured inside it, and replayed as graph segments with the collectives issued eagerly in between."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import random
    import numpy as np
    import torch.distributed as dist
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        hp = _hp()
        hp["p_mixed_noise"] = 0.6
        runs = {}
        for mode in ("eager",) + tuple(modes):
            os.environ["MSG_B200_NCCL_IN_GRAPH"] = "0" if mode == "segments" else "1"
            G, D = build(dev, seed=0)
            opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"], fused=True, capturable=True)
            opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"], fused=True, capturable=True)
            mw = ModelWrapper(G, D, opt_g, opt_d, hyperparameters=hp, device=dev, cuda_graphs=True)
            random.seed(7), np.random.seed(7), torch.manual_seed(7)
            gen = torch.Generator().manual_seed(11 + rank)          # every rank its own shard
            hist = []
            for it in range(6):
                if mode == "eager":
                    mw._graphs.clear()
                real = torch.rand(4, 2, 3, 32, 32, generator=gen).to(dev)
                hist.append({k: v.detach().cpu() for k, v in mw.train_step(real).items()})
            torch.cuda.synchronize(dev)
            n_graphs = {k[2:4]: sum(isinstance(i, torch.cuda.CUDAGraph) for i in st.program)
                        for k, st in mw._graphs.items() if st is not None}
            runs[mode] = dict(hist=hist, replays=mw.graph_replays, n_graphs=n_graphs,
                              params=[p.detach().cpu() for p in list(G.parameters()) + list(D.parameters())],
                              pl_mean=mw.path_length_regularization.mean_path_length.detach().cpu())
        torch.save(runs, os.path.join(tmp, "graph%d.pt" % rank))
    finally:
        os.environ.pop("MSG_B200_NCCL_IN_GRAPH", None)
        dist.destroy_process_group()


def _graph_forms(tmp_path, modes):
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_graph_worker, args=(WORLD, port, str(tmp_path), tuple(modes)), nprocs=WORLD, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "graph%d.pt" % r), weights_only=False) for r in range(WORLD)]
    want_graphs = {"in_graph": {(False, False): 1, (True, True): 1}, "segments": {(False, False): 4, (True, True): 5}}
    for r in range(WORLD):
        assert res[r]["eager"]["replays"] == 0
        for mode in modes:
            assert res[r][mode]["replays"] == 4 and res[r][mode]["n_graphs"] == want_graphs[mode]
            for a, b in zip(res[r]["eager"]["hist"], res[r][mode]["hist"]):
                assert set(a) == set(b)
                for k in a:
                    assert torch.allclose(a[k], b[k], rtol=2e-3, atol=1e-5), (mode, k, a[k], b[k])
            assert torch.allclose(res[r]["eager"]["pl_mean"], res[r][mode]["pl_mean"], rtol=2e-3)
            for a, b in zip(res[r]["eager"]["params"], res[r][mode]["params"]):
                assert rel_err(b, a) < 2e-3
            for a, b in zip(res[r][mode]["params"], res[0][mode]["params"]):
                assert torch.equal(a, b)


@pytest.mark.timeout(600)
def test_multi_rank_cuda_graph_segments_match_eager_over_nccl(built_library, tmp_path):
    """Multi-rank CUDA-graph replay, the default form: graph segments with the NCCL all-reduces issued eagerly in between;
    same losses / parameters as eager issue, replicas identical across ranks."""
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    _graph_forms(tmp_path, ("segments",))


@pytest.mark.timeout(600)
def test_nccl_all_reduce_inside_the_cuda_graph_matches_eager(built_library, tmp_path):
    """The opt-in form (MSG_B200_NCCL_IN_GRAPH=1): the all-reduces as nodes of the iteration's ONE graph
    (ModelWrapper._collectives_in_graph).  Runs only when MSG_B200_TEST_NCCL_IN_GRAPH=1: the form passed here on 2 GPUs
    (profiles/r2q_dist_gpu_tests.txt) but did not finish at the benchmark's size, and a hung collective
    must not be able to take a test run with it."""
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    if os.environ.get("MSG_B200_TEST_NCCL_IN_GRAPH") != "1":
        pytest.skip("opt-in: MSG_B200_TEST_NCCL_IN_GRAPH=1")
    _graph_forms(tmp_path, ("in_graph",))
