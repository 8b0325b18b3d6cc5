"""CPU (no GPU needed): the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares;
the host-side mirror keeps the reference's interface; CPU tensors are refused (no fallback)."""
import ctypes
import inspect
import os
import subprocess

import pytest
import torch

from util import ROOT


def test_library_exports_every_declared_symbol():
    from panonerf_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 40
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/panonerf_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    assert _lib.lib().pnb_abi_version() == 1


def test_library_contains_blackwell_tensor_core_code():
    from panonerf_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing: tcgen05 / TMA path not compiled in"


def test_argument_validation_without_gpu():
    from panonerf_b200 import _lib
    lib = _lib.lib()
    rc = lib.pnb_resample(4, 300, None, None, 0.01, 1, None, 0, None, None, None)        # N > 256
    assert rc == 10001 and b"N <= 256" in lib.pnb_last_error()
    rc = lib.pnb_linear_tc(128, 256, 100, None, 256, None, 256, None, 256, 1, None, None, 0, None, 0, 0, None, None)
    assert rc == 10001                                                                  # K % 16 != 0


def test_cpu_tensors_are_refused():
    from panonerf_b200 import ops
    from panonerf_b200.models import mip
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.pos_enc(torch.zeros(2, 3), 4)
    with pytest.raises(RuntimeError):
        mip.volumetric_rendering(torch.zeros(2, 4, 3), torch.zeros(2, 4, 1), torch.zeros(2, 5), torch.zeros(2, 3), False)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from panonerf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LibraryMissing):
        _lib.lib()


def test_module_interface_mirrors_reference():
    from panonerf_b200.models.mip_nerf import MipNeRF
    from panonerf_b200.models.pano_mip_nerf import PanoMipNeRF
    from panonerf_b200.models import mip
    from oracle.panonerf_oracle import mlp_shapes
    m = MipNeRF(num_samples=64, rgb_activation="softplus", mlp_name="mipnerf", num_env_samples=10)   # extra kwargs swallowed
    p = PanoMipNeRF(num_samples=64, rgb_activation="softplus", mlp_num_density_channels=5)
    assert sum(x.numel() for x in m.mlp.parameters()) == 612740
    assert sum(x.numel() for x in p.mlp.parameters()) == 613768
    want = mlp_shapes(c_density=5)
    got = {k: tuple(v.shape) for k, v in p.mlp.state_dict().items()}
    assert list(got) == list(want) and got == dict(want)
    assert list(inspect.signature(m.forward).parameters) == ["rays", "randomized", "white_bkgd", "use_ort_loss"]
    assert list(inspect.signature(p.forward).parameters) == ["rays", "env_rays", "randomized", "white_bkgd",
                                                             "enable_surf", "use_ort_loss"]
    for fn, args in (("sample_along_rays", ["origins", "directions", "radii", "num_samples", "near", "far", "randomized",
                                            "disparity", "ray_shape"]),
                     ("resample_along_rays", ["origins", "directions", "radii", "t_samples", "weights", "randomized",
                                              "ray_shape", "stop_grad", "resample_padding"]),
                     ("cast_rays", ["t_samples", "origins", "directions", "radii", "ray_shape"]),
                     ("volumetric_rendering", ["rgb", "density", "t_samples", "dirs", "white_bkgd"]),
                     ("integrated_pos_enc", ["means_covs", "min_deg", "max_deg"]),
                     ("pos_enc", ["x", "min_deg", "max_deg"]),
                     # SURVEY 8f rank 4 variants (mip.py:197-237, 486-527)
                     ("sample_each_points_hemisp", ["point_origins", "directions", "num_samples", "near", "far", "radii",
                                                    "randomized"]),
                     ("volumetric_lighting_composing", ["rgb", "density", "t_samples", "dirs", "white_bkgd", "output_t"])):
        assert list(inspect.signature(getattr(mip, fn)).parameters)[:len(args)] == args, fn
    from panonerf_b200.utils import surface_rendering as sr, vector_rotation as vr
    for fn in ("microfeast_brdf", "blinn_phong_brdf"):                     # utils/surface_rendering.py:6, 64
        assert list(inspect.signature(getattr(sr, fn)).parameters) == ["albedo", "normal", "roughness", "l", "v"]
    assert list(inspect.signature(sr.surface_rendering).parameters) == ["env", "albedo", "normal", "roughness", "l", "v",
                                                                        "solid_angle", "output_sd"]
    assert list(inspect.signature(vr.RotToTarget().rot2t).parameters) == ["tvec"]   # utils/vector_rotation.py:57
    with pytest.raises(NotImplementedError):
        MipNeRF(rgb_activation="sigmoid")                                  # mip_nerf.py:155-158
    with pytest.raises(NotImplementedError):
        PanoMipNeRF(rgb_activation="softplus", mlp_net_activation="gelu")


def test_lr_schedule_matches_reference_formula():
    from panonerf_b200.systems.base_system import mip_lr_decay
    from oracle.panonerf_oracle import mip_lr
    for step in (0, 1, 60, 120, 121, 22000, 44000, 50000):
        assert abs(mip_lr_decay(step, 2e-4, 2e-5, 44000, 120, 0.01) - mip_lr(step)) < 1e-15


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` is the CPU arm: it must run without a GPU and print one JSON line."""
    import json
    out = subprocess.run(["python", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-rays", "16"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_rays_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["e2e"]["h2d_bytes_per_step"] == 0


def test_adam_hyper_parameters_for_the_graph_step():
    """FlatAdam.next_hyper feeds the device-side update of GraphedTrainStep: lr from the MipLRDecay schedule and the
    two Adam bias corrections exactly as pnb_adam_step forms them on the host (1 - b1^t, sqrt(1 - b2^t))."""
    import math
    from panonerf_b200.systems.base_system import FlatAdam, mip_lr_decay

    opt = FlatAdam.__new__(FlatAdam)                       # no parameters needed for the schedule bookkeeping
    opt.step_count = 0
    opt.lr_fn = lambda s: mip_lr_decay(s, 2e-4, 2e-5, 44000, 120, 0.01)
    for t in range(1, 6):
        lr, bc1, bc2s = opt.next_hyper()
        assert opt.step_count == t
        assert lr == mip_lr_decay(t - 1, 2e-4, 2e-5, 44000, 120, 0.01)
        assert abs(bc1 - (1 - 0.9 ** t)) < 1e-15 and abs(bc2s - math.sqrt(1 - 0.999 ** t)) < 1e-15


def test_bench_roofline_report_schema():
    """bench.py's roofline object from a (fake) per-kernel profile: every MLP family is TENSOR-bound (SURVEY.md
    section 8d) and measured in TFLOP/s against the sustained peak; the programs of mlp_fused_kernel<P> are one
    family; the headline is the family with the most device time; the MLP stage and the whole step are reported;
    the HBM view (designed bytes, GB/s) sits beside it and the fused-ideal I/O is the algorithmic byte count."""
    import bench

    class Ev:
        def __init__(self, ms):
            self.ms = ms

        def elapsed_time(self, other):
            return other.ms

    prof = {"wgrad_batch": dict(events=[(Ev(0), Ev(2.0))], bytes=10e9, flops=1e12),
            "mlp_fused": dict(events=[(Ev(0), Ev(1.0))], bytes=1e9, flops=1.2e12),
            "mlp_fused_bwd": dict(events=[(Ev(0), Ev(1.5))], bytes=1e9, flops=0.3e12)}
    pk = dict(hbm=6500.0, tf_burst=1685.0, tf_sust=1382.0, src="measured")
    r = bench.roofline_from_profile(prof, 1, pk, step_ms=5.0, step_flop=2.5e12, ideal_bytes_per_step=3e8)
    assert r["kernel"] == "mlp_fused_kernel" and r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["achieved"] - 600.0) < 1e-6 and abs(r["frac"] - 600.0 / 1382.0) < 1e-9 and r["peak"] == 1382.0
    for k in ("traffic", "launches_per_step", "avg_launch_ms", "kernel_ms_per_step", "other_kernels", "hbm"):
        assert k in r
    assert set(r["programs"]) == {"mlp_fused", "mlp_fused_bwd"}
    o = r["other_kernels"]["wgrad_batch_kernel"]
    assert o["bound"] == "tensor" and abs(o["achieved"] - 500.0) < 1e-6
    assert abs(o["hbm"]["gbps"] - 5000.0) < 1e-6 and abs(o["hbm"]["frac_of_copy_bandwidth"] - 5000.0 / 6500.0) < 1e-9
    assert abs(r["mlp_stage"]["kernel_ms_per_step"] - 4.5) < 1e-9 and abs(r["mlp_stage"]["achieved"] - 2.5e3 / 4.5) < 1e-6
    assert abs(r["whole_step"]["achieved"] - 500.0) < 1e-6 and abs(r["whole_step"]["frac"] - 500.0 / 1382.0) < 1e-9
    assert r["algorithmic_bytes_per_step"] == 3e8 and r["designed_bytes_per_step"] == 12e9


def test_bench_rank_conditional_work_has_no_collectives():
    """bench.py: work that only some ranks do must not contain a collective.  The C1 line (MipNeRF step with its own
    FlatAdam, whose step all-reduces when a process group exists) once ran on rank 0 alone and dead-locked every N > 1
    run against the other ranks' render barrier."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    tree = ast.parse(src)
    run = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    guards = [ast.unparse(node.test) for node in ast.walk(run)
              if isinstance(node, ast.If) and any(isinstance(c, ast.Call) and getattr(c.func, "id", "") == "measure_c1"
                                                  for b in node.body for c in ast.walk(b))]
    assert "world == 1" in guards, guards


def test_mip_lr_decay_scheduler_class_matches_the_reference_class():
    """utils.lr_schedule.MipLRDecay: same constructor, same rates as the upstream class on a stock Adam, step by step."""
    import warnings
    from oracle import ref_harness
    from panonerf_b200.utils.lr_schedule import MipLRDecay
    if not ref_harness.available():
        pytest.skip("reference tree not present (GPU box)")
    ns = ref_harness.load()
    ref_cls = ns.lr_schedule.MipLRDecay if hasattr(ns, "lr_schedule") else None
    if ref_cls is None:
        import importlib
        ref_cls = importlib.import_module("utils.lr_schedule").MipLRDecay
    rates = []
    for cls in (MipLRDecay, ref_cls):
        p = torch.nn.Parameter(torch.zeros(3))
        opt = torch.optim.Adam([p], lr=5e-4)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sch = cls(opt, 5e-4, 5e-6, 2000, 250, 0.01)
            seq = []
            for _ in range(400):
                seq.append(opt.param_groups[0]["lr"])
                opt.step()
                sch.step()
        rates.append(seq)
    for a, b in zip(*rates):
        assert abs(a - b) <= 1e-12 * max(abs(b), 1e-12), (a, b)
