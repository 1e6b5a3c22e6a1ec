"""The G+D training step of the reference's trainer (multi_stylegan/model_wrapper.py:245-451), same order of
sub-steps and the same lazy-regularisation cadence, restructured for one-process-per-GPU execution:

  * `train_step(real_images)` = one iteration of `_gan_training`: D step -> lazy R1 (every 16th) ->
    [CutMix augmentation + consistency, probability 0 in epoch 0] -> G step -> lazy path length (every 16th,
    half batch) -> EMA.  Losses are returned as device scalars; nothing calls `.item()` on the hot path
    (the reference syncs 6-9 times per iteration for logging, :300-305,:327-329,:414-416,:442-444).
  * gradients are averaged across ranks with one flat all-reduce per optimiser step (dist.py); the
    path-length running mean is averaged too (the reference computes it from DataParallel's gathered batch).
  * the generator's unobservable second branch (multi_stylegan_generator.py:184,187,189) is not evaluated.
  * `cuda_graphs=True` replays an iteration as ONE CUDA graph (~2300 kernel launches of the plain iteration; one
    graph per (input shape, lazy R1, lazy path length, wrong-order) variant, captured the second time the variant
    occurs).  With several ranks the gradient all-reduces are graph breaks: the iteration becomes 2-4 graph segments in
    one memory pool with the NCCL calls issued eagerly between their replays.  Host-side random decisions stay on the host and enter the graph through device scalars: the
    style-mixing crossover index (misc.py:244-252 + multi_stylegan_generator.py:162-169) and the wrong-order frame
    permutation (:268-273).  Iterations the graph cannot express (a CutMix draw, top-k, fixed latents, the ADA
    wrapper's per-call host decisions) run eagerly on the same parameters and optimiser state.
"""
import copy
import os
import random
from types import SimpleNamespace
from typing import Any, Callable, Dict, Iterable, Optional

import numpy as np
import torch
import torch.nn as nn

from . import config as default_config
from . import dist as mdist
from . import loss, misc
from ._mode import higher_order_gradients
from .u_net_2d_discriminator import (generate_cut_mix_augmentation_data, generate_cut_mix_transformation_data)


_CAPTURE_STREAMS: Dict[Any, Any] = {}


def _capture_stream(device) -> "torch.cuda.Stream":
    """One capture stream per device for every wrapper and graph variant of the process: autograd's AccumulateGrad nodes
    remember the stream they were created on, and a capture that has to synchronise with another (non-capturing) side
    stream of an earlier capture is invalidated."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _CAPTURE_STREAMS:
        _CAPTURE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _CAPTURE_STREAMS[key]


class ModelWrapper(object):
    def __init__(self,
                 generator: nn.Module,
                 discriminator: nn.Module,
                 generator_optimizer: torch.optim.Optimizer,
                 discriminator_optimizer: torch.optim.Optimizer,
                 training_dataset: Optional[Iterable[torch.Tensor]] = None,
                 hyperparameters: Dict[str, Any] = default_config.generation_hyperparameters,
                 trap_weights_map: Optional[torch.Tensor] = None,
                 generator_loss: nn.Module = None,
                 discriminator_loss: nn.Module = None,
                 discriminator_regularization_loss: nn.Module = None,
                 cut_mix_augmentation_loss: nn.Module = None,
                 cut_mix_regularization_loss: nn.Module = None,
                 path_length_regularization: nn.Module = None,
                 generator_ema: Optional[nn.Module] = None,
                 device: str = "cuda",
                 process_group=None,
                 allow_tf32_matmul: bool = True,
                 cuda_graphs: bool = False) -> None:
        self.generator, self.discriminator = generator, discriminator
        self.generator_optimizer, self.discriminator_optimizer = generator_optimizer, discriminator_optimizer
        self.training_dataset = training_dataset
        self.hyperparameters = hyperparameters
        self.trap_weights_map = trap_weights_map
        self.generator_loss = generator_loss or loss.NonSaturatingLogisticGeneratorLoss()
        self.discriminator_loss = discriminator_loss or loss.NonSaturatingLogisticDiscriminatorLoss()
        self.discriminator_regularization_loss = discriminator_regularization_loss or loss.R1Regularization()
        self.cut_mix_augmentation_loss = cut_mix_augmentation_loss or loss.NonSaturatingLogisticDiscriminatorLossCutMix()
        self.cut_mix_regularization_loss = cut_mix_regularization_loss or nn.MSELoss(reduction="mean")
        self.path_length_regularization = path_length_regularization or loss.PathLengthRegularization()
        self.device = device
        self.process_group = process_group
        # The reference environment (PyTorch 1.8.1, requirements.txt:1) runs every CUDA matmul with TF32 enabled by
        # default; newer PyTorch turned that off, which sends the non-local block's bmm and its backward to an fp32
        # CUDA-core sgemm.  Restore the reference's setting for the library GEMMs that stay on torch.
        torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32_matmul)
        if generator_ema is None:
            generator_ema = copy.deepcopy(generator)
        self.generator_ema = generator_ema.eval()
        self.latent_dimensions = generator.latent_dimensions
        self.iteration = 0           # == the reference's progress_bar.n after update(n=1) (:255)
        self.epoch, self.epochs = 0, 1
        self.resume_training = False
        self.top_k: Callable = nn.Identity()
        # Decisions that change WHICH optimiser steps an iteration runs (the CutMix draw, :331) must agree on every rank,
        # or the ranks would issue different sequences of gradient all-reduces: with several ranks they are drawn from a
        # generator that is seeded identically everywhere (the reference is a single process and has one draw anyway).
        self._structure_rng = random.Random(0x5EED)
        self.cuda_graphs = bool(cuda_graphs)
        self._graphs: Dict[Any, Any] = {}        # variant key -> None (seen once, ran eagerly) | captured state
        self._capture = None                     # the program being captured (graph segments + eager collectives)
        self._always_break = False               # tests: cut the capture at every optimiser step on a single rank too
        self.graph_replays = 0
        self.graph_launches = 0                  # kernels of this library replayed from graphs (bench accounting)

    # ---- helpers ------------------------------------------------------------------------------------
    def _noise(self, batch: int):
        return misc.get_noise(batch_size=batch, latent_dimension=self.latent_dimensions,
                              p_mixed_noise=self.hyperparameters["p_mixed_noise"], device=self.device)

    def _d_params(self):
        return [p for p in self.discriminator.parameters()]

    def _reduce_begin(self, params, extra=()):
        """Start the cross-rank average of the gradients (and of `extra` state tensors); returns a token for
        `_reduce_end`.  The collective is asynchronous, so what the caller issues before `_reduce_end` overlaps the
        transfer.  While an iteration is being captured the collective is a graph break: the capture is cut into segments
        and the all-reduce is issued eagerly between their replays on the tensors the segments produce (static
        addresses in the shared graph pool)."""
        if not (mdist.world_size(self.process_group) > 1 or self._always_break):
            return None
        tensors = [p.grad for p in params if p.grad is not None] + list(extra)
        prog = self._capture
        if prog is None or self._collectives_in_graph():
            # (while capturing over NCCL: the flatten, the collective on NCCL's stream and the join are graph nodes)
            return ("eager", mdist.all_reduce_tensors_begin(tensors, self.process_group))
        box = {}
        self._segment_end(prog)
        prog.items.append(lambda t=tensors, g=self.process_group, b=box: b.__setitem__("p", mdist.all_reduce_tensors_begin(t, g)))
        self._segment_begin(prog)
        return ("graph", box)

    def _reduce_end(self, token, params, optimizer) -> None:
        """Wait for the averaged gradients, clip, step."""
        if token is not None:
            kind, state = token
            if kind == "eager":
                if state is not None:
                    state.wait()
            else:
                prog = self._capture
                self._segment_end(prog)
                prog.items.append(lambda b=state: b["p"].wait() if b.get("p") is not None else None)
                self._segment_begin(prog)
        torch.nn.utils.clip_grad_norm_(params, max_norm=5.)
        optimizer.step()

    def _optimize(self, params, optimizer, extra=()) -> None:
        """Cross-rank average of the gradients (and of `extra` state tensors), clip, step — without anything in between
        (one graph break instead of two while capturing)."""
        if mdist.world_size(self.process_group) > 1 or self._always_break:
            prog = self._capture
            if prog is None or self._collectives_in_graph():
                mdist.all_reduce_gradients(params, self.process_group)
                mdist.all_reduce_tensors(list(extra), self.process_group)
            else:
                self._segment_end(prog)
                tensors = [p.grad for p in params if p.grad is not None] + list(extra)
                prog.items.append(lambda t=tensors, g=self.process_group: mdist.all_reduce_tensors(t, g))
                self._segment_begin(prog)
        torch.nn.utils.clip_grad_norm_(params, max_norm=5.)
        optimizer.step()

    def _collectives_in_graph(self) -> bool:
        """Opt-in (MSG_B200_NCCL_IN_GRAPH=1): NCCL collectives are stream work like any kernel and can be captured
        (NCCL >= 2.9); the iteration then stays ONE graph — the all-reduce runs on NCCL's stream between a captured fork and
        join — instead of 2-4 segments with the collectives issued eagerly in between.  Parity of the two forms:
        tests/test_dist_gpu.py (2 GPUs, small networks).  Not the default: at the benchmark's size (203 / 212 MB buffers)
        2-GPU runs of the one-graph form did not finish, so the segments stay the product path."""
        if self._always_break or os.environ.get("MSG_B200_NCCL_IN_GRAPH", "0") != "1":
            return False
        return (mdist.world_size(self.process_group) > 1
                and torch.distributed.get_backend(self.process_group) == "nccl")

    @staticmethod
    def _segment_begin(prog) -> None:
        graph = torch.cuda.CUDAGraph()
        # thread_local: CUDA calls of other threads (the pinned-memory allocator's event queries, autograd worker
        # threads of an earlier eager pass, NCCL's watchdog) do not invalidate this thread's capture
        ctx = torch.cuda.graph(graph, pool=prog.pool, stream=prog.stream,
                               capture_error_mode=os.environ.get("MSG_B200_CAPTURE_MODE", "thread_local"))
        ctx.__enter__()
        prog.open = (graph, ctx)

    @staticmethod
    def _segment_end(prog) -> None:
        graph, ctx = prog.open
        ctx.__exit__(None, None, None)
        prog.items.append(graph)
        prog.open = None

    def _zero(self) -> None:
        self.discriminator_optimizer.zero_grad()
        self.generator_optimizer.zero_grad()

    def _trap(self):
        use = self.hyperparameters["trap_weight"] * self.epochs <= self.epoch or self.resume_training
        return self.trap_weights_map if use else None

    # ---- one iteration of _gan_training (model_wrapper.py:253-451) -----------------------------------
    def train_step(self, real_images: torch.Tensor, z_d=None, z_g=None, z_pl=None,
                   pl_noise=None) -> Dict[str, torch.Tensor]:
        """real_images [B, 2, 3, H, W] on the device.  z_* optionally fix the latent draws (parity tests).
        With `cuda_graphs` the returned scalars live in the graph's static memory: read them before the next call."""
        hp = self.hyperparameters
        self.iteration += 1
        lazy_r1 = self.iteration % hp["lazy_discriminator_regularization"] == 0
        lazy_pl = self.iteration % hp["lazy_generator_regularization"] == 0
        wrong_order = bool(self.epoch >= hp["wrong_order_start"] * self.epochs or self.resume_training)
        rng = self._structure_rng if mdist.world_size(self.process_group) > 1 else random
        cut_mix = (rng.random() <= ((0.5 / float(self.epochs)) * float(self.epoch))) \
            or (self.resume_training and rng.random() <= 0.5)
        if self.cuda_graphs and self._graphable(real_images, cut_mix, z_d, z_g, z_pl, pl_noise):
            return self._train_step_graphed(real_images, lazy_r1, lazy_pl, wrong_order)
        return self._iteration(real_images, lazy_r1, lazy_pl, wrong_order, cut_mix, z_d, z_g, z_pl, pl_noise)

    # ---- CUDA-graph replay of one iteration ----------------------------------------------------------
    def reset_cuda_graphs(self) -> None:
        """Drop every captured iteration (they are re-captured on demand).  Needed after anything that replaces the
        tensors a graph has baked in: parameters or optimiser state re-created (`module.to(...)`, a new optimiser),
        a different process group.  `load_state_dict` copies in place and needs no reset."""
        self._graphs.clear()

    def _graphable(self, real_images, cut_mix, *fixed) -> bool:
        return (real_images.is_cuda and not cut_mix and all(f is None for f in fixed)
                and isinstance(self.top_k, nn.Identity)
                and getattr(self.discriminator, "graph_capturable", True)
                and getattr(self.generator, "graph_capturable", True))

    def _draw_inject_index(self) -> int:
        """The host half of misc.get_noise + Generator._latent: a crossover index in [1, n_latent - 2] with probability
        p_mixed_noise, otherwise n_latent (= every layer takes the first latent, i.e. no mixing)."""
        n_latent = self.generator.n_latent
        p = self.hyperparameters["p_mixed_noise"]
        if (p > 0) and (random.random() < p):
            return int(np.random.randint(1, n_latent - 1))
        return n_latent

    def _train_step_graphed(self, real_images, lazy_r1, lazy_pl, wrong_order) -> Dict[str, torch.Tensor]:
        from . import _C
        # every host decision the captured program bakes in is part of the key (the trap weighting switches on at
        # trap_weight * epochs, :262-263 via _trap())
        use_trap = self._trap() is not None
        key = (tuple(real_images.shape), real_images.dtype, lazy_r1, lazy_pl, wrong_order, use_trap)
        if key not in self._graphs:
            # first occurrence: eager, which also creates optimiser state, running means and kernel attributes
            self._graphs[key] = None
            inject = torch.tensor([self._draw_inject_index() for _ in range(3 if lazy_pl else 2)] + [0] * (0 if lazy_pl else 1),
                                  dtype=torch.int64).to(real_images.device)
            return self._iteration(real_images, lazy_r1, lazy_pl, wrong_order, False, None, None, None, None,
                                   inject=inject)
        st = self._graphs[key]
        if st is None:
            for opt in (self.generator_optimizer, self.discriminator_optimizer):
                if not all(g.get("capturable", False) for g in opt.param_groups):
                    raise RuntimeError("cuda_graphs=True needs optimisers constructed with capturable=True")
            st = SimpleNamespace()
            if use_trap and self.trap_weights_map.device != real_images.device:
                self.trap_weights_map = self.trap_weights_map.to(real_images.device)     # no H2D copy inside the capture
            st.real = real_images.detach().clone()
            st.inject = torch.zeros(3, dtype=torch.int64, device=real_images.device)
            st.perm = torch.arange(real_images.shape[2], dtype=torch.int64, device=real_images.device)
            plr = self.path_length_regularization
            st.mean_path_length = plr.mean_path_length.detach().to(real_images.device).clone()
            torch.cuda.synchronize(real_images.device)
            launches0 = _C.launch_count()
            prog = SimpleNamespace(items=[], pool=torch.cuda.graph_pool_handle(),
                                   stream=_capture_stream(real_images.device), open=None)
            self._capture = prog
            ada_begin = getattr(self.discriminator, "begin_plan_capture", None)
            if ada_begin is not None:
                ada_begin(2 * real_images.shape[0], real_images.device)
            self._segment_begin(prog)
            try:
                plr.mean_path_length = st.mean_path_length
                st.out = self._iteration(st.real, lazy_r1, lazy_pl, wrong_order, False, None, None, None, None,
                                         inject=st.inject, perm=st.perm)
                if lazy_pl:      # the running mean is rebound by the loss module; keep it in the graph's static slot
                    st.mean_path_length.copy_(plr.mean_path_length.detach())
            finally:
                self._capture = None
                if prog.open is not None:
                    self._segment_end(prog)
                st.ada = self.discriminator.end_plan_capture() if ada_begin is not None else None
            # the loss module was rebound to tensors of the capture (never computed: capturing does not execute); the
            # running mean lives in the static slot the graph reads and writes
            plr.mean_path_length = st.mean_path_length
            st.program = prog.items      # CUDA graphs in capture order (one shared pool), eager collectives in between
            st.launches = _C.launch_count() - launches0
            self._graphs[key] = st
        # per-replay host decisions -> device scalars (fill kernels take the value as a launch argument: no host buffer)
        for i in range(3 if lazy_pl else 2):
            st.inject[i].fill_(self._draw_inject_index())
        if wrong_order:
            from . import misc as _misc
            for i, v in enumerate(_misc.random_permutation(real_images.shape[2]).tolist()):
                st.perm[i].fill_(int(v))
        st.real.copy_(real_images, non_blocking=True)
        plr = self.path_length_regularization
        if lazy_pl and plr.mean_path_length is not st.mean_path_length:
            # an eager lazy iteration (CutMix draw, top-k, another graph variant) moved the running mean since this
            # variant last ran: its static slot must start from the current value
            st.mean_path_length.copy_(plr.mean_path_length.detach().to(st.mean_path_length.device), non_blocking=True)
        if st.ada is not None:
            self.discriminator.refresh_plans(st.ada)     # fresh augmentation draws for every ADA call of the iteration
        for item in st.program:
            if isinstance(item, torch.cuda.CUDAGraph):
                item.replay()
            else:
                item()
        if st.ada is not None:
            self.discriminator.after_replay(st.ada)
        if lazy_pl:
            plr.mean_path_length = st.mean_path_length
        self.graph_replays += 1
        self.graph_launches += st.launches
        return dict(st.out)

    def _generate(self, batch: int, z, inject, slot: int, **kw):
        if z is None and inject is not None:
            # graph form of the mixed-noise draw: both latents always, crossover index from the device
            z2 = torch.randn(2, batch, self.latent_dimensions, dtype=torch.float32, device=self.device)
            return self.generator(input=list(z2.unbind(0)), inject_index=inject[slot], **kw)
        return self.generator(input=self._noise(batch) if z is None else z, **kw)

    def _iteration(self, real_images, lazy_r1, lazy_pl, wrong_order, cut_mix, z_d, z_g, z_pl, pl_noise,
                   inject=None, perm=None) -> Dict[str, torch.Tensor]:
        hp = self.hyperparameters
        out: Dict[str, torch.Tensor] = {}
        B = real_images.shape[0]
        g_params = [p for p in self.generator.parameters()]
        d_params = self._d_params()

        # ---------------- discriminator step (:258-305) ----------------
        self._zero()
        with torch.no_grad():
            fake_images = self._generate(B, z_d, inject, 0)
        if wrong_order:
            n = max(1, int(hp["batch_factor_wrong_order"] * B))
            order = misc.random_permutation(real_images.shape[2]) if perm is None else perm
            fake_images = torch.cat([fake_images, real_images[:n, :, order]], 0)
        pair = getattr(self.discriminator, "forward_pair", None)
        if pair is not None and fake_images.shape == real_images.shape:
            # the reference's two calls (:279-283) as one batched pass with identical results (see forward_pair)
            (real_pred, real_pred_px), (fake_pred, fake_pred_px) = pair(real_images, fake_images)
        else:
            real_pred, real_pred_px = self.discriminator(real_images, is_real=True, is_cut_mix=False)
            fake_pred, fake_pred_px = self.discriminator(fake_images, is_real=False, is_cut_mix=False)
        l_real, l_fake = self.discriminator_loss(real_pred, fake_pred)
        l_real_px, l_fake_px = self.discriminator_loss(real_pred_px, fake_pred_px, weight=self._trap())
        (l_real + l_fake + l_real_px + l_fake_px).backward()
        # With several ranks and nothing between this step and the generator step that needs the updated discriminator
        # (no lazy R1, no CutMix), the gradient all-reduce overlaps the generator step's generator forward, which does
        # not depend on the discriminator (same results, different issue order).
        overlap = not lazy_r1 and not cut_mix
        pending = self._reduce_begin(d_params) if overlap else None
        if not overlap:
            self._optimize(d_params, self.discriminator_optimizer)
        out.update(loss_discriminator_real=l_real.detach(), loss_discriminator_fake=l_fake.detach(),
                   loss_discriminator_real_pixel_wise=l_real_px.detach(),
                   loss_discriminator_fake_pixel_wise=l_fake_px.detach())

        # ---------------- lazy R1 (:307-329) ----------------
        if lazy_r1:
            self._zero()
            real_r1 = real_images.detach().requires_grad_(True)
            with higher_order_gradients():      # R1 differentiates the discriminator's input gradient (loss.py:311-316)
                rp, rp_px = self.discriminator(real_r1, is_real=False, is_cut_mix=True)
            r1 = self.discriminator_regularization_loss(rp, real_r1, rp_px)
            (hp["w_discriminator_regularization_r1"] * r1).backward()
            self._optimize(d_params, self.discriminator_optimizer)
            out["loss_discriminator_regularization"] = r1.detach()

        # ---------------- CutMix augmentation + consistency (:331-376) ----------------
        if cut_mix:
            self._zero()
            images, label = generate_cut_mix_augmentation_data(real_images, fake_images)
            _, pred = self.discriminator(images, is_cut_mix=True)
            cm_real, cm_fake = self.cut_mix_augmentation_loss(pred, label)
            (hp["w_discriminator_regularization"] * (cm_real + cm_fake)).backward()
            self._optimize(d_params, self.discriminator_optimizer)
            out["loss_cut_mix_augmentation"] = (cm_real + cm_fake).detach()
            self.discriminator_optimizer.zero_grad()
            images, label = generate_cut_mix_transformation_data(real_images.detach(), fake_images.detach(),
                                                                 real_pred_px.detach(), fake_pred_px.detach())
            _, pred = self.discriminator(images, is_cut_mix=True)
            cm_reg = self.cut_mix_regularization_loss(pred, label)
            (hp["w_discriminator_regularization"] * cm_reg).backward()
            self._optimize(d_params, self.discriminator_optimizer)
            out["loss_cut_mix_regularization"] = cm_reg.detach()

        # ---------------- generator step (:377-416) ----------------
        if overlap:
            # (the generator's gradients are still None from the zero_grad before the discriminator step, whose generator
            # pass ran under no_grad: the forward can be recorded before this phase's zero_grad)
            fake_images = self._generate(B, z_g, inject, 1)
            self._reduce_end(pending, d_params, self.discriminator_optimizer)
            self._zero()
        else:
            self._zero()
            fake_images = self._generate(B, z_g, inject, 1)
        # The reference lets autograd compute all discriminator weight gradients here and then discards them
        # (zero_grad of both optimisers precedes the next phase, :260-261,:379-380); they are unobservable, so the
        # discriminator is differentiated w.r.t. its input only.
        for p in d_params:
            p.requires_grad_(False)
        try:
            fake_pred, fake_pred_px = self.discriminator(fake_images, is_real=False, is_cut_mix=False)
        finally:
            for p in d_params:
                p.requires_grad_(True)
        top = self.top_k(fake_pred)
        if isinstance(top, tuple):
            fake_pred, indexes = top
            fake_pred_px = fake_pred_px[indexes]
        else:
            fake_pred = top
        l_g = self.generator_loss(fake_pred)
        l_g_px = self.generator_loss(fake_pred_px, weight=self._trap())
        (l_g + l_g_px).backward()
        self._optimize(g_params, self.generator_optimizer)
        out.update(loss_generator=l_g.detach(), loss_generator_pixel_wise=l_g_px.detach())

        # ---------------- lazy path length (:418-444) ----------------
        if lazy_pl:
            self._zero()
            n = max(1, int(hp["batch_size_shrink_path_length_regularization"] * B))
            grads = self._generate(n, z_pl, inject, 2, return_path_length_grads=True, path_length_noise=pl_noise)
            pl_loss, path_length = self.path_length_regularization(grads)
            (hp["w_generator_regularization"] * pl_loss).backward()
            self._optimize(g_params, self.generator_optimizer, extra=[self.path_length_regularization.mean_path_length])
            out.update(path_length=path_length.detach().mean(), loss_path_length_regularization=pl_loss.detach())

        # ---------------- EMA (:446) ----------------
        misc.exponential_moving_average(model_ema=self.generator_ema, model_train=self.generator)
        return out

    def _gan_training(self, resume_training: bool = False, top_k: nn.Module = nn.Identity()):
        """One epoch over `training_dataset` (model_wrapper.py:245-253); yields the per-iteration scalars."""
        self.resume_training, self.top_k = resume_training, top_k
        history = []
        for real_images in self.training_dataset:
            out = self.train_step(real_images.to(self.device, non_blocking=True))
            # with cuda_graphs the scalars alias the graph's static memory: keep values, not views of the next replay
            history.append({k: v.detach().clone() for k, v in out.items()})
        return history

    def train(self, epochs: int = 20, resume_training: bool = False, top_k: bool = False):
        """Epoch loop (model_wrapper.py:104-145) without the logging / validation / checkpoint shell."""
        self.epochs = epochs
        tk = nn.Identity()
        if top_k:
            n = len(self.training_dataset)
            tk = loss.TopK(starting_iteration=int(self.hyperparameters["top_k_start"] * epochs * n),
                           final_iteration=int(self.hyperparameters["top_k_finish"] * epochs * n))
            if resume_training:
                tk.starting_iteration, tk.final_iteration = 0, 1
        history = []
        for self.epoch in range(epochs):
            if hasattr(self.training_dataset, "set_epoch"):      # dataset.DeviceLoader: a new shared permutation per epoch
                self.training_dataset.set_epoch(self.epoch)
            self.generator.train()
            self.discriminator.train()
            history.extend(self._gan_training(resume_training=resume_training, top_k=tk))
        return history
