"""Import the *unmodified* reference from /root/reference (TEST INFRASTRUCTURE, build container only).

The reference does not import under this image's library versions:
``networks/models.py:23`` needs torchmeta, ``trainer/losses.py:19`` needs
pytorch_metric_learning, ``networks/ResNet.py:23`` needs ``torchvision.models.utils`` and
``utils/__init__.py:19`` pulls imgaug.  None of those touch hot-path arithmetic, so four stub
modules registered in ``sys.modules`` are enough.  No reference source is copied.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root():
    """/root/reference in the build container; on the GPU box the byte-identical copy that
    ``oracle/make_ref.py`` materialised into ``oracle/_ref/`` (git-ignored, shipped with the snapshot)."""
    env = os.environ.get("B200NP_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/networks"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REFERENCE_ROOT = _default_root()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the stubs and put the reference on sys.path (idempotent)."""
    import torch.nn as nn

    if "torchmeta" not in sys.modules:
        class _Meta(nn.Module):
            pass
        mods = _stub("torchmeta.modules", MetaModule=_Meta, MetaSequential=nn.Sequential,
                     MetaConv2d=nn.Conv2d, MetaBatchNorm2d=nn.BatchNorm2d, MetaLinear=nn.Linear)
        _stub("torchmeta", modules=mods)
        def _unavailable(*a, **k):
            raise NotImplementedError("torchmeta is not installed in this image (MAML paths are outside the hot path)")
        gb = _stub("torchmeta.utils.gradient_based", gradient_update_parameters=_unavailable)
        _stub("torchmeta.utils", gradient_based=gb, gradient_update_parameters=_unavailable)
    if "pytorch_metric_learning" not in sys.modules:
        losses = _stub("pytorch_metric_learning.losses", NTXentLoss=object)
        _stub("pytorch_metric_learning", losses=losses)
    try:
        import torchvision.models.utils  # noqa: F401
    except Exception:
        import torchvision.models as tvm
        u = _stub("torchvision.models.utils", load_state_dict_from_url=lambda *a, **k: {})
        tvm.utils = u
    if "imgaug" not in sys.modules:
        aug = _stub("imgaug.augmenters")
        _stub("imgaug", augmenters=aug, seed=lambda *a, **k: None, ALL="ALL")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def make_config(method, task, tasks_per_batch, agg_mode, img_agg, dim_w=None, dim_r=None,
                dim_z=None, n_hidden_units_r=None, seed=2578, device="cpu", temperature=0.07):
    """Plain namespace carrying the attributes the model constructors read
    (configs/config.py:33-104) without the mkdir/log side effects of ``Config``."""
    img_size, input_dim, output_dim = {
        "shapenet_3d": ([64, 64, 4], 4, 4),
        "shapenet_1d": ([128, 128, 1], 3, 2),
        "distractor": ([128, 128, 1], 2, 2),
    }[task]
    return types.SimpleNamespace(
        method=method, task=task, tasks_per_batch=tasks_per_batch, agg_mode=agg_mode,
        img_agg=img_agg, dim_w=dim_w, dim_r=dim_r, dim_z=dim_z, n_hidden_units_r=n_hidden_units_r,
        seed=seed, device=device, temperature=temperature, img_size=img_size,
        input_dim=input_dim, output_dim=output_dim, loss_type="mse", beta=0, contrastive=False)


def reference_class(method):
    install()
    import importlib
    return getattr(importlib.import_module("networks." + method), method)


def reference_lossfunc():
    install()
    import importlib
    return importlib.import_module("trainer.losses").LossFunc
