"""Test infrastructure ONLY (never imported by the product path).

Imports the *unmodified* upstream reference from ``/root/reference`` so that the oracle restatement in
``oracle/panonerf_oracle.py`` can be pinned against it and golden vectors can be generated
(``tests/golden/make_golden.py``).  ``/root/reference`` exists only in the build container; on the GPU box
``available()`` is False and every caller must skip.

The reference needs three modules that are not installed here: ``pytorch_lightning``,
``OpenEXR``/``Imath`` and ``matplotlib``.  They are replaced by inert stand-ins that provide just the names the
reference touches at import time; none of them is on the arithmetic path.
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("PANONERF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "mip.py"))


class _AttrDict(dict):
    __getattr__ = dict.__getitem__


def _install_shims():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            global_step = 0

            def save_hyperparameters(self, hp):
                self._hp = _AttrDict(hp)

            @property
            def hparams(self):
                return self._hp

            def log(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.Trainer = object
        cb = types.ModuleType("pytorch_lightning.callbacks")
        cb.ModelCheckpoint = object
        cb.TQDMProgressBar = object
        pl.callbacks = cb
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.callbacks"] = cb
    if "OpenEXR" not in sys.modules:
        exr = types.ModuleType("OpenEXR")
        exr.InputFile = exr.OutputFile = exr.Header = object
        sys.modules["OpenEXR"] = exr
    if "Imath" not in sys.modules:
        im = types.ModuleType("Imath")
        im.PixelType = im.Channel = object
        sys.modules["Imath"] = im
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        cm = types.ModuleType("matplotlib.cm")
        mpl.pyplot, mpl.cm = plt, cm
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.cm": cm})
    try:
        import lpips  # noqa: F401
    except Exception:
        lp = types.ModuleType("lpips")
        lp.LPIPS = lambda *a, **k: None
        sys.modules["lpips"] = lp


_loaded = {}


def load():
    """Return a namespace with the reference's hot-path modules (imported, not copied)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if _loaded:
        return _loaded["ns"]
    _install_shims()
    # The reference uses top-level package names `models`, `utils`, `datasets`, `systems`, `configs`.
    # Guard against name clashes with anything already imported under those names.
    for clash in ("models", "utils", "datasets", "systems", "configs"):
        mod = sys.modules.get(clash)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF_ROOT):
            for k in [k for k in sys.modules if k == clash or k.startswith(clash + ".")]:
                del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        import models.mip as mip
        import models.mip_nerf as mip_nerf
        import models.pano_mip_nerf as pano_mip_nerf
        import utils.surface_rendering as surface_rendering
        import datasets.base_datasets as base_datasets
        import datasets.pano_datasets as pano_datasets
        import utils.lr_schedule as lr_schedule
        try:
            import systems.mipnerf_system as mipnerf_system
            import systems.panonerf_system as panonerf_system
        except Exception as e:  # systems are optional (need cv2 etc.)
            mipnerf_system = panonerf_system = None
            _loaded["systems_error"] = repr(e)
    finally:
        sys.path.remove(REF_ROOT)
    ns = types.SimpleNamespace(
        mip=mip, mip_nerf=mip_nerf, pano_mip_nerf=pano_mip_nerf, surface_rendering=surface_rendering,
        base_datasets=base_datasets, pano_datasets=pano_datasets, lr_schedule=lr_schedule,
        mipnerf_system=mipnerf_system, panonerf_system=panonerf_system, Rays=base_datasets.Rays)
    _loaded["ns"] = ns
    return ns


def make_pano_dataset(ns, h, w, c2w, near=0.0, far=10.0):
    """Build a reference PanoDataset without touching disk and run its own ray generator."""
    import numpy as np
    ds = ns.pano_datasets.PanoDataset.__new__(ns.pano_datasets.PanoDataset)
    ds.h, ds.w, ds.near, ds.far = h, w, near, far
    ds.reform_cam = False
    ds.camtoworlds = [np.asarray(c, dtype=np.float32) for c in c2w]
    ds._generate_rays()
    return ds
