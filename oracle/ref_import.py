"""Import the UNMODIFIED reference `models` package from /root/reference on a box that
lacks Lightning / kornia / piq / torch_optimizer / robust_loss_pytorch.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by oracle/make_golden.py (fixture
generation, in the build container) and by `-m "not gpu"` tests that cross-check the oracle
when the reference tree is present.  Nothing on the product path imports this file.

The reference imports third-party packages at module scope
(/root/reference/models/srmodel.py:9-19, models/srgan.py:4-13).  We register inert stand-ins
in ``sys.modules`` so that `import models` succeeds; `SRModel.__init__` then only touches
`nn.L1Loss` and `optim.Adam` (srmodel.py:134-137 with the default ctor args).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SRB200_REFERENCE", "/root/reference")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed third-party symbol called")


def _lenient_module(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)

    def __getattr__(attr):  # noqa: N807
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Dummy

    mod.__getattr__ = __getattr__  # type: ignore[attr-defined]
    return mod


class _LightningModule(nn.Module):
    """Just enough of lightning.pytorch.LightningModule for SRModel.__init__."""

    def save_hyperparameters(self, *a, **k):
        pass

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            import torch
            return torch.device("cpu")

    def log_dict(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "common.py"))


def import_reference_models():
    """Returns the reference's `models` package (module object), imported under the private
    name ``_ref_models`` so it can coexist with this repo's own drop-in `models` package."""
    if "_ref_models" in sys.modules:
        return sys.modules["_ref_models"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")

    stubs = {}
    pl = types.ModuleType("lightning.pytorch")
    pl.LightningModule = _LightningModule
    lightning = types.ModuleType("lightning")
    lightning.pytorch = pl
    loggers = types.ModuleType("lightning.pytorch.loggers")
    loggers.CometLogger = type("CometLogger", (), {})
    loggers.TensorBoardLogger = type("TensorBoardLogger", (), {})
    pl.loggers = loggers
    stubs["lightning"] = lightning
    stubs["lightning.pytorch"] = pl
    stubs["lightning.pytorch.loggers"] = loggers

    kornia = _lenient_module("kornia")
    kaug = _lenient_module("kornia.augmentation")
    kcol = _lenient_module("kornia.color")
    kfil = _lenient_module("kornia.filters")
    kornia.augmentation, kornia.color, kornia.filters = kaug, kcol, kfil
    stubs.update({"kornia": kornia, "kornia.augmentation": kaug, "kornia.color": kcol,
                  "kornia.filters": kfil})
    stubs["piq"] = _lenient_module("piq")
    stubs["torch_optimizer"] = _lenient_module("torch_optimizer")
    stubs["robust_loss_pytorch"] = _lenient_module("robust_loss_pytorch")
    # shadow the reference's own `losses` package (it imports kornia at module scope,
    # /root/reference/losses/edge_loss.py:5-6); only names are needed at import time.
    losses = _lenient_module("losses")
    losses_losses = _lenient_module("losses.losses")
    losses.losses = losses_losses
    stubs["losses"] = losses
    stubs["losses.losses"] = losses_losses

    saved = {k: sys.modules.get(k) for k in list(stubs) + ["models"]}
    saved_sub = {k: v for k, v in sys.modules.items() if k.startswith("models.")}
    for k in saved_sub:
        del sys.modules[k]
    sys.modules.pop("models", None)
    sys.modules.update(stubs)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ref = importlib.import_module("models")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # move the reference package out of the way of this repo's own `models`
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            sys.modules["_ref_" + k] = sys.modules.pop(k)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            elif k in sys.modules and k != "models":
                del sys.modules[k]
        sys.modules.update(saved_sub)
    return sys.modules["_ref_models"]
