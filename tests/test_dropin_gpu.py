"""Drop-in behaviour of the module API in the reference's training context (SURVEY.md section 8b "precision context",
train.py:79-93): Lightning runs `training_step` under torch.autocast('cuda', fp16) + GradScaler with a stock
torch.optim.Adam on `mip_nerf.mlp.parameters()`, and wraps the system in DistributedDataParallel when there is more
than one GPU.  The reference tree is not on the GPU box, so the loss is the oracle's restatement of
systems/panonerf_system.py:15-75 (pinned to the verbatim training_step by tests/golden/make_golden.py and
tests/test_oracle_vs_reference.py) evaluated on OUR model's outputs with plain torch ops under autocast.

Also here: data-parallel gradient equivalence (SURVEY.md section 4 iv / train.py:92): the all-reduced mean of two
ranks' gradients equals the one-GPU gradient of the concatenated batch (needs 2 GPUs; `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden
from util import O, T, golden_rays, golden_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(g, precision):
    from panonerf_b200.models.pano_mip_nerf import PanoMipNeRF
    model = PanoMipNeRF(num_samples=int(g["n"]), rgb_activation="softplus", rgb_padding=0.0,
                        mlp_net_width=int(g["width"]), mlp_num_density_channels=5, num_env_samples=10,
                        precision=precision).to(DEV)
    model.mlp.load_state_dict(golden_state_dict(g))
    return model


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_step_under_autocast_gradscaler_and_stock_adam(precision):
    from panonerf_b200 import _lib
    if precision == "bf16" and not _lib.lib().pnb_tc_available():
        pytest.skip("not an sm_100 device")
    g = load_golden("panonerf_w256.npz")
    rays, env = golden_rays(g, DEV)
    env16 = type(env)(*[x.half() for x in env])                      # datasets/pano_datasets.py:263: env rays are fp16
    gt = T(g["gt"]).to(DEV)

    def step(amp):
        model = _model(g, precision)
        opt = torch.optim.Adam(model.mlp.parameters(), lr=2e-4)      # systems/base_system.py:81-87
        scaler = torch.amp.GradScaler("cuda", enabled=amp)
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            out = model(rays=rays, env_rays=env16, randomized=False, white_bkgd=False, enable_surf=True,
                        use_ort_loss=True)
            loss = O.panonerf_loss(out, rays, gt)                     # torch ops on CUDA tensors, under autocast
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        grads = {k: p.grad.detach().clone() for k, p in model.mlp.named_parameters()}
        before = {k: p.detach().clone() for k, p in model.mlp.named_parameters()}
        scaler.step(opt)
        scaler.update()
        moved = sum(float((p.detach() - before[k]).abs().sum()) for k, p in model.mlp.named_parameters())
        return float(loss), grads, moved, out

    loss_amp, g_amp, moved, out = step(True)
    loss_ref, g_ref, _, _ = step(False)
    assert out[1][0].dtype == torch.float32 and torch.isfinite(out[1][0]).all()
    tol = 1e-5 if precision == "fp32" else 1e-3
    assert abs(loss_amp - loss_ref) <= tol * abs(loss_ref), (loss_amp, loss_ref)
    if precision == "fp32":
        assert abs(loss_amp - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    assert moved > 0.0, "GradScaler skipped the step (non-finite gradients?)"
    for k in g_ref:
        assert torch.isfinite(g_amp[k]).all(), k
        a, b = g_amp[k].double().flatten(), g_ref[k].double().flatten()
        # the kernels run in their own precision whatever the autocast state: same gradient up to the fp32 rounding of
        # the 65536x loss scale and the float-atomic order of the head-bias sums
        assert float((a - b).norm()) <= 2e-3 * float(b.norm()) + 1e-12, (k, float((a - b).norm() / b.norm()))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_distributed_data_parallel_wrapper_world1():
    """DistributedDataParallel (what Lightning's strategy='ddp' builds, train.py:92) around the system: its autograd
    hooks see the gradients our hand-written backward returns, and the step equals the unwrapped one."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    g = load_golden("panonerf_w64.npz")
    rays, env = golden_rays(g, DEV)
    gt = T(g["gt"]).to(DEV)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV, 0))
    try:
        grads = []
        for wrap in (False, True):
            from panonerf_b200.models.pano_mip_nerf import PanoMipNeRF
            model = PanoMipNeRF(num_samples=int(g["n"]), rgb_activation="softplus", rgb_padding=0.0,
                                mlp_net_width=int(g["width"]), mlp_num_density_channels=5, num_env_samples=10,
                                precision="fp32").to(DEV)
            model.mlp.load_state_dict(golden_state_dict(g))
            net = DDP(model, device_ids=[0]) if wrap else model
            out = net(rays=rays, env_rays=env, randomized=False, white_bkgd=False, enable_surf=True, use_ort_loss=True)
            O.panonerf_loss(out, rays, gt).backward()
            grads.append({k: p.grad.detach().clone() for k, p in model.mlp.named_parameters()})
        for k in grads[0]:
            a, b = grads[1][k].double().flatten(), grads[0][k].double().flatten()
            assert float((a - b).norm()) <= 1e-5 * float(b.norm()) + 1e-12, k
    finally:
        dist.destroy_process_group()


def _dp_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from panonerf_b200.systems.base_system import default_hparams
    from panonerf_b200.systems.panonerf_system import PanoNeRFSystem
    g = load_golden("panonerf_w64.npz")
    rays, env = golden_rays(g, dev)
    gt = T(g["gt"]).to(dev)
    n = gt.shape[0] // world

    def system():
        hp = default_hparams("panonerf", precision="fp32")
        hp.update({"nerf.num_samples": int(g["n"]), "nerf.mlp.net_width": int(g["width"]), "train.randomized": False})
        s = PanoNeRFSystem(hp).to(dev)
        s.mip_nerf.mlp.load_state_dict(golden_state_dict(g))
        s.env_rays = env
        return s

    # this rank's shard: local mean loss -> flat gradient -> one NCCL all-reduce (sum), scale 1/world
    s = system()
    opt = s.configure_optimizers()
    opt.zero_grad()
    sl = slice(rank * n, (rank + 1) * n)
    s.training_step((type(rays)(*[x[sl].contiguous() for x in rays]), gt[sl].contiguous())).backward()
    scale = opt.all_reduce_grads()
    dp = (opt.flat_g * scale).cpu()
    if rank == 0:      # one-GPU gradient of the concatenated batch
        s1 = system()
        o1 = s1.configure_optimizers()
        o1.zero_grad()
        s1.training_step((type(rays)(*[x[:n * world].contiguous() for x in rays]), gt[:n * world].contiguous())).backward()
        ret["dp"], ret["single"], ret["scale"] = dp, o1.flat_g.cpu(), scale
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_concatenated_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    dp, single = ret["dp"].double(), ret["single"].double()
    assert ret["scale"] == 0.5
    assert float((dp - single).norm()) <= 1e-4 * float(single.norm()), float((dp - single).norm() / single.norm())
