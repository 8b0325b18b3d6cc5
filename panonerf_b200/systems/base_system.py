"""Host-side mirror of systems/base_system.py without the Lightning dependency: model construction from the flat
`hparams` dict (same keys as configs/*.yaml after configs/config.py flattening), the optimiser (fused Adam on one
flat parameter buffer + MipLRDecay) and the data-parallel gradient all-reduce."""
import math

import torch
import torch.distributed as dist

from .. import field, ops
from ..datasets.base_datasets import Rays


def mip_lr_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1.0):
    """utils/lr_schedule.py:51-60."""
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(
            0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0.0), 1.0)
    return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


class FlatAdam:
    """torch.optim.Adam (defaults) semantics on ONE contiguous fp32 buffer: parameters and their gradients are
    re-homed as views of flat buffers so that the data-parallel all-reduce is a single NCCL call on 2.45 MB and the
    update is a single kernel (pnb_adam_step) instead of 24 (systems/base_system.py:81-87 + train.py:92 DDP)."""

    def __init__(self, params, lr_fn):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                p.grad = self.flat_g[off:off + k].view_as(p)
                p._pnb_direct_grad = True      # the fused backward accumulates straight into these views (field.py)
                off += k
        self.lr_fn = lr_fn
        self.step_count = 0
        field.invalidate_packs()

    def zero_grad(self):
        self._rebind_grads(fold=False)
        self.flat_g.zero_()

    def _rebind_grads(self, fold=True):
        """Every p.grad must stay a view of `flat_g` (the fused backward accumulates straight into it, field.py).
        `module.zero_grad()` / `set_to_none=True` or an external autograd call can replace it: stray gradients are
        folded into the flat buffer (so the update never runs on a silently empty buffer) and the views re-installed."""
        off = 0
        for p in self.params:
            k = p.numel()
            view = self.flat_g[off:off + k].view_as(p)
            g = p.grad
            if g is None or g.data_ptr() != view.data_ptr() or g.shape != view.shape:
                if fold and g is not None:
                    view.add_(g.to(view.dtype))
                p.grad = view
            off += k

    def all_reduce_grads(self):
        """Sum over ranks (mean taken inside the Adam kernel via grad_scale = 1/world)."""
        from ..parallel import all_reduce_flat_
        return all_reduce_flat_(self.flat_g)

    def step(self):
        self._rebind_grads()
        scale = self.all_reduce_grads()
        self.step_count += 1
        lr = self.lr_fn(self.step_count - 1)
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, lr, self.step_count, grad_scale=scale)
        field.invalidate_packs()
        return lr

    # ---- CUDA-graph variant: the step-dependent scalars live in device memory ---------------------------------
    def next_hyper(self, beta1=0.9, beta2=0.999):
        """Advance the step counter; returns (lr, 1 - beta1^t, sqrt(1 - beta2^t)) for the device-side update."""
        self.step_count += 1
        t = self.step_count
        return self.lr_fn(t - 1), 1.0 - beta1 ** t, math.sqrt(1.0 - beta2 ** t)

    def step_dev(self, hyper):
        if not torch.cuda.is_current_stream_capturing():
            self._rebind_grads()
        scale = self.all_reduce_grads()
        ops.adam_step_dev(self.flat_p, self.flat_g, self.m, self.v, hyper, grad_scale=scale)
        field.invalidate_packs()


class GraphedTrainStep:
    """One training step (zero_grad -> training_step -> backward -> all-reduce -> Adam) captured in a CUDA graph and
    replayed: ~240 kernel launches become one graph launch, which removes the host-side launch cost (it matters most
    with 8 ranks sharing the host cores).  Inputs are copied into static buffers, the learning rate and the Adam bias
    corrections reach the update kernel through a 3-float device tensor, so the schedule still advances.  Results are
    those of the eager step (same kernels, same order); `torch.rand` streams are graph-safe Philox streams.

    With more than one rank the NCCL all-reduce of the flat gradient buffer and the update kernel are captured too
    (NCCL supports stream capture), so a step is a single graph launch on every rank; `PNB_GRAPH_NCCL=0` keeps them
    outside the graph.

    Step-dependent Python branches of `training_step` (`train.surface_start_step`) are frozen into a captured graph,
    so the step key (`system.graph_key()`) is re-evaluated before every replay and the graph is re-captured when it
    flips; `system.global_step` advances once per step like Lightning's."""

    def __init__(self, system, opt, rays, gts, warmup=3):
        self.system, self.opt = system, opt
        dev = gts.device
        self.dev = dev
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        import os
        self.update_in_graph = (not multi) or os.environ.get("PNB_GRAPH_NCCL", "1") != "0"
        self.rays = type(rays)(*[x.clone() for x in rays])
        self.gts = gts.clone()
        self.hyper = torch.zeros(3, device=dev, dtype=torch.float32)
        self.launches_per_step = 0
        self.captures = 0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._set_hyper()
                self._body(capturing=False)
                system.global_step += 1
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._set_hyper()
        try:
            self._capture()
        except Exception:                      # noqa: BLE001 - e.g. a communicator that refuses stream capture
            if not (multi and self.update_in_graph):
                raise
            self.update_in_graph = False       # keep the collective and the update outside the graph
            torch.cuda.synchronize(dev)
            self._capture()
        self._replay()                         # the captured step itself has not run yet: run it once
        system.global_step += 1

    def _key(self):
        fn = getattr(self.system, "graph_key", None)
        return fn() if fn is not None else None

    def _capture(self):
        torch.cuda.synchronize(self.dev)
        self.key = self._key()
        l0 = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.launches_per_step = ops.launch_count() - l0
        self.captures += 1

    def _replay(self):
        self.graph.replay()
        if not self.update_in_graph:
            self.opt.step_dev(self.hyper)
        # the replayed update changed the parameters behind torch's back: cached bf16 weight packs are stale for any
        # eager forward that follows (e.g. a validation render); inside the graph the re-pack kernels are captured
        field.invalidate_packs()

    def _set_hyper(self):
        # a fresh pageable tensor per step: the driver stages it before returning, so the host may run many steps
        # ahead of the device without overwriting a value that is still to be copied
        self.hyper.copy_(torch.tensor(self.opt.next_hyper(), dtype=torch.float32))

    def _body(self, capturing=True):
        self.opt.zero_grad()
        loss = self.system.training_step((self.rays, self.gts))
        loss.backward()
        if self.update_in_graph or not capturing:
            self.opt.step_dev(self.hyper)
        return loss.detach()

    def __call__(self, rays=None, gts=None):
        """Run one step on (rays, gts) (or on the current contents of the static buffers); returns the loss tensor
        (static: read it before the next call)."""
        if rays is not None:
            for dst, src in zip(self.rays, rays):
                dst.copy_(src, non_blocking=True)
            self.gts.copy_(gts, non_blocking=True)
        if self._key() != self.key:            # a step-dependent branch flipped: the captured graph is out of date
            self._capture()
        self._set_hyper()
        self._replay()
        self.system.global_step += 1
        return self.loss


class AccumulatedTrainStep:
    """One optimiser step on a batch that is walked in `micro_batches` equal slices (gradient accumulation): the
    graph holds forward + backward of ONE slice with the loss pre-scaled by 1/micro_batches and is replayed per slice;
    the flat gradient buffer accumulates (field.py writes `+=`), then one all-reduce and one Adam update follow.
    Equal slices make the accumulated gradient the gradient of the whole-batch mean loss (systems/*_system.py divide
    by `mask.sum()` of the batch; `lossmult` is 1 for panoramas, datasets/pano_datasets.py:192).  This is how
    BASELINE.json's 65 536 rays/GPU step (config C4) runs in a bounded activation footprint: the saved planes of a
    65 536-ray step would be ~150 GB, those of an 8192-ray slice are ~19 GB."""

    def __init__(self, system, opt, rays, gts, micro_batches, warmup=2):
        n = gts.shape[0]
        if n % micro_batches:
            raise ValueError("the batch must divide into equal micro-batches")
        self.system, self.opt, self.k, self.mb = system, opt, micro_batches, n // micro_batches
        dev = gts.device
        self.dev = dev
        self.rays = type(rays)(*[x[:self.mb].clone() for x in rays])
        self.gts = gts[:self.mb].clone()
        self.hyper = torch.zeros(3, device=dev, dtype=torch.float32)
        self.loss_sum = torch.zeros((), device=dev, dtype=torch.float32)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._slice_body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        opt.zero_grad()
        # the bf16 weight packs must be rebuilt inside the graph: the parameters change between optimiser steps
        field.invalidate_packs()
        l0 = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._slice_body()
        self.launches_per_slice = ops.launch_count() - l0
        opt.zero_grad()

    def _slice_body(self):
        loss = self.system.training_step((self.rays, self.gts)) * (1.0 / self.k)
        loss.backward()
        self.loss_sum += loss.detach()

    def __call__(self, rays, gts):
        self.opt.zero_grad()
        self.loss_sum.zero_()
        for i in range(self.k):
            sl = slice(i * self.mb, (i + 1) * self.mb)
            for dst, src in zip(self.rays, rays):
                dst.copy_(src[sl], non_blocking=True)
            self.gts.copy_(gts[sl], non_blocking=True)
            self.graph.replay()
        self.hyper.copy_(torch.tensor(self.opt.next_hyper(), dtype=torch.float32))
        self.opt.step_dev(self.hyper)
        self.system.global_step += 1
        return self.loss_sum


class BaseSystem(torch.nn.Module):
    """systems/base_system.py:9-55 (model construction) + :81-87 (optimiser)."""

    def __init__(self, hparams):
        super().__init__()
        self.hparams = dict(hparams)
        hp = self.hparams
        self.train_randomized = hp["train.randomized"]
        self.val_randomized = hp["val.randomized"]
        self.white_bkgd = hp["train.white_bkgd"]
        self.val_chunk_size = hp["val.chunk_size"]
        self.batch_size = hp["train.batch_size"]
        self.global_step = 0
        if hp["nerf.mlp_name"] == "mipnerf":
            from ..models.mip_nerf import MipNeRF as NeRFModel
            num_density_channels = 1
        elif hp["nerf.mlp_name"] == "panonerf":
            from ..models.pano_mip_nerf import PanoMipNeRF as NeRFModel
            num_density_channels = 5
        else:
            raise ValueError(hp["nerf.mlp_name"])
        self.mip_nerf = NeRFModel(
            num_samples=hp["nerf.num_samples"], num_levels=hp["nerf.num_levels"],
            resample_padding=hp["nerf.resample_padding"], stop_resample_grad=hp["nerf.stop_resample_grad"],
            use_viewdirs=hp["nerf.use_viewdirs"], disparity=hp["nerf.disparity"], ray_shape=hp["nerf.ray_shape"],
            min_deg_point=hp["nerf.min_deg_point"], max_deg_point=hp["nerf.max_deg_point"],
            deg_view=hp["nerf.deg_view"], density_activation=hp["nerf.density_activation"],
            density_noise=hp["nerf.density_noise"], density_bias=hp["nerf.density_bias"],
            rgb_activation=hp["nerf.rgb_activation"], alb_activation=hp["nerf.alb_activation"],
            rgb_padding=hp["nerf.rgb_padding"], disable_integration=hp["nerf.disable_integration"],
            append_identity=hp["nerf.append_identity"], mlp_net_depth=hp["nerf.mlp.net_depth"],
            mlp_net_width=hp["nerf.mlp.net_width"], mlp_net_depth_condition=hp["nerf.mlp.net_depth_condition"],
            mlp_net_width_condition=hp["nerf.mlp.net_width_condition"], mlp_skip_index=hp["nerf.mlp.skip_index"],
            mlp_num_rgb_channels=hp["nerf.mlp.num_rgb_channels"], mlp_num_density_channels=num_density_channels,
            mlp_net_activation=hp["nerf.mlp.net_activation"], mlp_name=hp["nerf.mlp_name"],
            num_env_samples=hp["nerf.num_env_samples"], precision=hp.get("precision"))
        self.env_rays = None

    def render_rays_per_launch(self, n_rays):
        """Rays per forward of `render_image`.  Upstream walks a panorama in `val.chunk_size` = 512-ray chunks (1024
        sequential forwards, GPU mostly idle - SURVEY.md section 8f rank 2).  Rays are independent, so chunking never
        changes a pixel (tests/test_models_gpu.py checks bit-identity); here the whole image is ONE forward - a fixed,
        small number of launches over all H*W rays - unless its transient buffers (~0.8 KB per sample: Gaussians, raw
        outputs, the fp32 encoding gradient of the normals) would not fit in half of the free device memory, and
        `val.chunk_size` is a no-op knob."""
        per_ray = 768 * int(self.hparams["nerf.num_samples"]) + 8192
        free, _ = torch.cuda.mem_get_info()
        fit = max(1024, int(0.5 * free / per_ray) // 1024 * 1024)
        return min(int(n_rays), fit)

    def _render_into(self, rays, height, width, channels, forward, chunk_size=None):
        """Render driver: forwards over ray ranges (normally one), per-ray results scattered straight into the
        [sum(channels), H, W] image planes by one kernel per range (csrc/image.cu) - no list-of-chunks, no cat, no
        permute (replaces models/mip.py:530-547 rearrange_render_image + systems/*_system.py compose())."""
        n = height * width
        flat = type(rays)(*[ops._f32c(x.reshape(n, x.shape[-1])) for x in rays])
        per = int(chunk_size) if chunk_size else self.render_rays_per_launch(n)
        out = torch.empty(sum(channels), n, device=flat[0].device, dtype=torch.float32)
        with torch.no_grad():
            for r0 in range(0, n, per):
                part = type(rays)(*[x[r0:r0 + per] for x in flat])
                ops.pack_chw(forward(part), out, r0)
        images, c0 = [], 0
        for c in channels:
            images.append(out[c0:c0 + c].view(1, c, height, width))
            c0 += c
        return images

    def configure_optimizers(self):
        hp = self.hparams
        lr_fn = lambda s: mip_lr_decay(s, hp["optimizer.lr_init"], hp["optimizer.lr_final"], hp["optimizer.max_steps"],
                                       hp["optimizer.lr_delay_steps"], hp["optimizer.lr_delay_mult"])
        return FlatAdam(self.mip_nerf.mlp.parameters(), lr_fn)

    def graph_key(self):
        """Everything `training_step` branches on in Python (a captured CUDA graph is only valid while it is fixed)."""
        hp = self.hparams
        return (bool(self.global_step >= hp.get("train.surface_start_step", 0) and hp.get("train.surface", False)),
                bool(hp.get("loss.ort_loss", 0) > 0), bool(hp.get("loss.chrom_loss", 0) > 0), self.train_randomized)

    @staticmethod
    def _inv_mask_sum(mask):
        """1 / mask.sum() as a device scalar (systems/panonerf_system.py:44-50 divide by `mask.sum()`): a fixed-order
        device reduction, no host synchronisation, so the step stays CUDA-graph capturable."""
        return torch.reciprocal(ops.dsum(mask))

    # ---- losses shared by both systems ---------------------------------------------------------------------------
    @staticmethod
    def _gt_ldr(rgbs):
        return ops.hdr_to_ldr(ops._f32c(rgbs[..., :3]), quantize=True)     # systems/*_system.py:17 / :24

    @staticmethod
    def _masked_mse(pred_hdr, gt_ldr, mask, inv_mask_sum):
        return ops.tonemap_mse(pred_hdr, gt_ldr, mask, inv_mask_sum)


def default_hparams(mlp_name="panonerf", **over):
    """configs/{mipnerf,panonerf}.yaml flattened the way configs/config.py:14-32 does (values after literal_eval)."""
    hp = {
        "seed": 4, "train.batch_size": 512, "train.batch_type": "all_images", "train.randomized": True,
        "train.white_bkgd": False, "train.surface": mlp_name == "panonerf", "train.surface_start_step": 0,
        "val.randomized": False, "val.white_bkgd": False, "val.chunk_size": 512,
        "nerf.mlp_name": mlp_name, "nerf.num_env_samples": 10, "nerf.num_ray_samples": 10, "nerf.num_samples": 64,
        "nerf.num_levels": 2, "nerf.resample_padding": 0.01, "nerf.stop_resample_grad": True,
        "nerf.use_viewdirs": True, "nerf.disparity": False, "nerf.ray_shape": "cone", "nerf.min_deg_point": 0,
        "nerf.max_deg_point": 16, "nerf.deg_view": 4, "nerf.density_activation": "softplus",
        "nerf.density_noise": 0.0, "nerf.density_bias": -1.0, "nerf.rgb_activation": "softplus",
        "nerf.alb_activation": "sigmoid", "nerf.rgb_padding": 0, "nerf.disable_integration": False,
        "nerf.append_identity": "Ture", "nerf.mlp.num_density_channels": 5, "nerf.mlp.net_depth": 8,
        "nerf.mlp.net_width": 256, "nerf.mlp.net_depth_condition": 1, "nerf.mlp.net_width_condition": 128,
        "nerf.mlp.net_activation": "relu", "nerf.mlp.skip_index": 4, "nerf.mlp.num_rgb_channels": 3,
        "optimizer.lr_init": 2e-4, "optimizer.lr_final": 2e-5, "optimizer.lr_delay_steps": 120,
        "optimizer.lr_delay_mult": 0.01, "optimizer.max_steps": 44000,
        "loss.coarse_loss_mult": 0.1, "loss.surface_loss": 1, "loss.ort_loss": 0.1 if mlp_name == "panonerf" else 0,
        "loss.chrom_loss": 0.1, "range": (0, 10),
    }
    hp.update(over)
    return hp
