"""SRModel — the plugin boundary (mirror of /root/reference/models/srmodel.py:67-621).

Same constructor signature, same step methods and the same extension recipe as the reference
(README.md:97-101: subclass SRModel, implement `forward`, register in models/__init__.py).
Works with Lightning when it is installed (`lightning.pytorch.LightningModule` base) and without
it (a minimal stand-in base), because the Lightning dependency is not available on every box
this runs on.  Only the torch built-in losses (l1/l2/mae/mse) and PSNR/SSIM are implemented
natively; the perceptual extras of the reference (piq / kornia / FLIP / adaptive) are looked up
lazily and raise if their package is missing (out of the hot-path scope).
"""
from __future__ import annotations

import itertools
import logging
import os
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Callable

import torch
import torch.nn as nn
import torch.optim as optim

try:  # pragma: no cover - depends on the environment
    import lightning.pytorch as pl
    _Base = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class _Base(nn.Module):
        """Stand-in for LightningModule: the attributes SRModel and its callers touch."""

        def __init__(self):
            super().__init__()
            self.hparams = {}
            self.current_epoch = 0
            self.global_step = 0
            self.trainer = None
            self.loggers = []
            self.logged = {}

        def save_hyperparameters(self, *args, **kwargs):
            import inspect
            frame = inspect.currentframe().f_back
            hp = {}
            while frame is not None and "self" in frame.f_locals and frame.f_locals["self"] is self:
                if frame.f_code.co_name == "__init__":
                    loc = dict(frame.f_locals)
                    extra = loc.pop("kwargs", {}) or {}
                    for k, v in itertools.chain(loc.items(), extra.items()):
                        if k not in ("self", "__class__") and not k.startswith("_"):
                            hp.setdefault(k, v)
                frame = frame.f_back
            self.hparams = hp

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        def log(self, name, value, **kwargs):
            self.logged[name] = value

        def log_dict(self, d, **kwargs):
            self.logged.update(d)


@dataclass
class _SubLoss:
    name: str
    loss: Callable
    weight: float = 1.


class _L1(nn.Module):
    """nn.L1Loss() (reference srmodel.py:37) — on CUDA fp32 NCHW inputs it is the fused
    libsrb200 loss+seed-gradient kernel, otherwise the torch built-in."""

    def forward(self, sr, hr):
        if sr.is_cuda and sr.dtype == torch.float32 and hr.dtype == torch.float32 and sr.shape == hr.shape:
            from srb200.functional import l1_loss
            return l1_loss(sr, hr)
        return nn.functional.l1_loss(sr, hr)


def _lazy(pkg: str, attr: str):
    def make(*a, **k):
        import importlib
        try:
            mod = importlib.import_module(pkg)
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"loss/metric '{attr}' needs the optional package '{pkg}': {e}") from e
        return getattr(mod, attr)(*a, **k)
    return make


class _LazyObject:
    """Callable built from `factory()` on first call."""

    def __init__(self, factory):
        self._factory, self._obj = factory, None

    def __call__(self, *a, **k):
        if self._obj is None:
            self._obj = self._factory()
        return self._obj(*a, **k)


def psnr(sr: torch.Tensor, hr: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """piq.psnr(data_range=1, reduction='mean') restatement (reference srmodel.py:52)."""
    mse = ((sr.float() - hr.float()) ** 2).flatten(1).mean(1)
    return (-10.0 * torch.log10(mse + eps)).mean()


def ssim(sr: torch.Tensor, hr: torch.Tensor, k1=0.01, k2=0.03, sigma=1.5, size=11) -> torch.Tensor:
    """piq.ssim defaults (gaussian 11x11, sigma 1.5, data_range 1, valid convolution, mean)."""
    sr, hr = sr.float(), hr.float()
    c = sr.shape[1]
    ax = torch.arange(size, dtype=torch.float32, device=sr.device) - (size - 1) / 2
    g = torch.exp(-(ax ** 2) / (2 * sigma ** 2))
    g = (g / g.sum())
    kern = (g[:, None] * g[None, :]).expand(c, 1, size, size).contiguous()
    f = max(1, round(min(sr.shape[-2:]) / 256))
    if f > 1:
        sr = nn.functional.avg_pool2d(sr, f)
        hr = nn.functional.avg_pool2d(hr, f)
    mu_x = nn.functional.conv2d(sr, kern, groups=c)
    mu_y = nn.functional.conv2d(hr, kern, groups=c)
    sxx = nn.functional.conv2d(sr * sr, kern, groups=c) - mu_x ** 2
    syy = nn.functional.conv2d(hr * hr, kern, groups=c) - mu_y ** 2
    sxy = nn.functional.conv2d(sr * hr, kern, groups=c) - mu_x * mu_y
    c1, c2 = k1 ** 2, k2 ** 2
    cs = (2 * sxy + c2) / (sxx + syy + c2)
    ss = (2 * mu_x * mu_y + c1) / (mu_x ** 2 + mu_y ** 2 + c1) * cs
    return ss.mean(dim=(1, 2, 3)).mean()


_supported_losses = {
    'adaptive': _lazy('robust_loss_pytorch', 'AdaptiveImageLossFunction'),
    'dists': _lazy('piq', 'DISTS'),
    'edge_loss': _lazy('losses', 'EdgeLoss'),
    'flip': _lazy('losses', 'FLIPLoss'),
    'haarpsi': _lazy('piq', 'HaarPSILoss'),
    'l1': _L1,
    'l2': nn.MSELoss,
    'lpips': _lazy('piq', 'LPIPS'),
    'mae': _L1,
    'mse': nn.MSELoss,
    'pencil_sketch': _lazy('losses', 'PencilSketchLoss'),
    'pieapp': _lazy('piq', 'PieAPP'),
}

_supported_metrics = {
    'BRISQUE': _lazy('piq', 'brisque'),
    'FLIP': _lazy('losses', 'FLIP'),
    'LPIPS': _lazy('piq', 'LPIPS'),
    'MS-SSIM': _lazy('piq', 'multi_scale_ssim'),
    'PSNR': psnr,
    'SSIM': ssim,
}

def _lazy_optimizer(name: str):
    class _Missing:
        __name__ = name

        def __new__(cls, *a, **k):
            import importlib
            try:
                mod = importlib.import_module('torch_optimizer')
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(f"optimizer '{name}' needs the optional package 'torch_optimizer': {e}") from e
            return getattr(mod, name)(*a, **k)
    return _Missing


# same names as the reference registry (srmodel.py:30-66); entries whose package (piq / kornia-based `losses` /
# torch_optimizer / robust_loss_pytorch) is not installed are registered lazily and raise only when USED
_supported_optimizers = {
    'ADAM': optim.Adam,
    'Ranger': _lazy_optimizer('Ranger'),
    'RangerVA': _lazy_optimizer('RangerVA'),
    'RangerQH': _lazy_optimizer('RangerQH'),
    'RMSprop': optim.RMSprop,
    'SGD': optim.SGD,
}


class SRModel(_Base, ABC):
    """Base module for super-resolution models (reference srmodel.py:67-143)."""

    def __init__(self,
                 batch_size: int = 16,
                 channels: int = 3,
                 default_root_dir: str = '.',
                 devices: None | list[int] | str | int = None,
                 eval_datasets: list[str] = ['DIV2K', 'Set5', 'Set14', 'B100', 'Urban100'],
                 log_loss_every_n_epochs: int = 5,
                 log_weights_every_n_epochs: int = 50,
                 losses: str = 'l1',
                 max_epochs: int = -1,
                 metrics: list[str] = ['PSNR', 'SSIM'],
                 metrics_for_pbar: list[str] = ['PSNR', 'SSIM'],
                 model_gpus: list[str] = [],
                 model_parallel: bool = False,
                 optimizer: str = 'ADAM',
                 optimizer_params: list[str] = [],
                 patch_size: int = 128,
                 precision: int = 32,
                 predict_datasets: list[str] = [],
                 save_results: int = -1,
                 save_results_from_epoch: str = 'last',
                 scale_factor: int = 4,
                 **kwargs: dict[str, Any]):
        super().__init__()
        self._logger = logging.getLogger(__name__)
        self.save_hyperparameters()
        self.example_input_array = torch.zeros(batch_size, channels, patch_size // scale_factor,
                                               patch_size // scale_factor)
        if model_parallel:
            raise ValueError('model_parallel is vestigial in the reference (srmodel.py:115-124) and is not '
                             'supported by the B200 path; use data-parallel devices instead')
        self._model_parallel = False
        self._model_gpus = None
        self._batch_size = batch_size
        self._channels = channels
        self._default_root_dir = default_root_dir
        self._eval_datasets = eval_datasets
        self._last_epoch = max_epochs
        self._log_loss_every_n_epochs = log_loss_every_n_epochs
        self._log_weights_every_n_epochs = log_weights_every_n_epochs
        self._losses = self._create_losses(losses, patch_size, precision)
        self._metrics = self._create_metrics(metrics)
        self._metrics_for_pbar = metrics_for_pbar
        self._optim, self._optim_params = self._parse_optimizer_config(optimizer, optimizer_params)
        self._predict_datasets = predict_datasets
        self._save_results = save_results
        self._save_results_from_epoch = save_results_from_epoch
        self._scale_factor = scale_factor
        self._training_step_outputs = []
        self._validation_step_outputs = []
        # B200 path: arithmetic type of the conv kernels ("bf16" tcgen05 path, "fp32" parity mode).
        # Not a hyper-parameter of the reference, so it never enters save_hyperparameters().
        cd = os.environ.get('SRB200_DTYPE', kwargs.get('compute_dtype', 'bf16'))
        self.compute_dtype = cd

    # ---- B200 path configuration ---------------------------------------------------------------
    @property
    def compute_dtype(self) -> str:
        return self._compute_dtype

    @compute_dtype.setter
    def compute_dtype(self, value: str):
        v = str(value).lower()
        if v not in ('bf16', 'fp32'):
            raise ValueError(f"compute_dtype must be 'bf16' or 'fp32', got {value}")
        self._compute_dtype = v

    @property
    def act_dtype(self) -> torch.dtype:
        return torch.bfloat16 if self._compute_dtype == 'bf16' else torch.float32

    # ---- optimisation -----------------------------------------------------------------------
    def configure_optimizers(self):
        """reference srmodel.py:145-154."""
        parameters_list = [self.parameters()]
        for loss in self._losses:
            if loss.name.find('adaptive') >= 0:
                parameters_list.append(loss.loss.parameters())
        trainable = filter(lambda p: p.requires_grad, itertools.chain(*parameters_list))
        return [self._optim(trainable, **self._optim_params)]

    @abstractmethod
    def forward(self, x):
        pass

    # ---- steps ------------------------------------------------------------------------------
    def training_step(self, batch, batch_idx):
        """reference srmodel.py:160-171."""
        img_sr = self.forward(batch['lr'])
        result = self._calculate_losses(img_sr=img_sr, img_hr=batch['hr'])
        self._training_step_outputs.append(result)
        return result

    def on_train_epoch_end(self):
        self._training_step_outputs.clear()

    def validation_step(self, batch, batch_idx, dataloader_idx=0):
        """reference srmodel.py:214-232 (image dumping, :234-343, is logger glue and left out)."""
        img_lr, img_hr = batch['lr'], batch['hr']
        img_sr = self.forward(img_lr)
        assert img_sr.size() == img_hr.size(), \
            f'Output size for image {self._eval_datasets[dataloader_idx]}/{batch.get("path")} should be ' \
            f'{img_hr.size()}, instead is {img_sr.size()}'
        img_hr = img_hr.clamp(0, 1)
        img_sr = img_sr.clamp(0, 1)
        result = self._calculate_metrics(img_sr=img_sr, img_hr=img_hr, dataloader_idx=dataloader_idx)
        self._validation_step_outputs.append(result)
        return result

    def on_validation_epoch_end(self):
        self._validation_step_outputs.clear()

    def predict_step(self, batch, batch_idx, dataloader_idx=0):
        """reference srmodel.py:375-380."""
        return self.forward(batch['lr']).clamp(0, 1)

    # ---- factories --------------------------------------------------------------------------
    def _create_losses(self, losses_str: str, patch_size: int, precision: int = 32) -> list[_SubLoss]:
        """'0.5*l1+0.5*l2' mini-DSL (reference srmodel.py:435-501)."""
        losses = []
        for loss in losses_str.split('+'):
            parts = loss.split('*')
            if len(parts) == 2:
                weight, loss_type = parts
                try:
                    weight = float(weight)
                except ValueError:
                    raise ValueError(f'{weight} is not a valid number to be used as weight for loss function '
                                     f'{loss_type.strip()}')
            else:
                weight, loss_type = 1., parts[0]
            loss_type = loss_type.strip().lower()
            if loss_type not in _supported_losses:
                raise AttributeError(f'Couldn\'t find loss {loss_type}. Supported losses: '
                                     f'{", ".join(_supported_losses)}')
            if loss_type == 'adaptive':
                fn = _supported_losses[loss_type](image_size=(patch_size, patch_size, 3),
                                                  float_dtype=torch.float32 if precision == 32 else torch.float16,
                                                  device=self.device)
            else:
                fn = _supported_losses[loss_type]()
            losses.append(_SubLoss(name=loss_type, loss=fn, weight=weight))
        return losses

    def _create_metrics(self, metrics: list[str]) -> list[tuple[str, Callable]]:
        used = []
        for metric in metrics:
            if metric not in _supported_metrics:
                raise AttributeError(f'Couldn\'t find metric {metric}. Supported metrics: '
                                     f'{", ".join(_supported_metrics)}')
            fn = _supported_metrics[metric]
            if metric in {'LPIPS', 'FLIP'}:      # metric objects (reference srmodel.py:507-509): built on first use, so a
                fn = _LazyObject(fn)             # missing optional package fails when the metric is computed, not at construction
            used.append((metric, fn))
        return used

    def _calculate_losses(self, img_sr: torch.Tensor, img_hr: torch.Tensor) -> dict[str, torch.Tensor]:
        """reference srmodel.py:519-565."""
        names, values = [], []
        for sub in self._losses:
            if sub.name in {'haarpsi', 'pieapp'}:
                loss = sub.loss(torch.clamp(img_sr, 0, 1), img_hr)
            elif sub.name == 'adaptive':
                loss = torch.mean(sub.loss.lossfun((img_sr - img_hr)).permute(0, 3, 2, 1))
            else:
                loss = sub.loss(img_sr, img_hr)
                if loss.dim() > 0:
                    loss = loss.mean()
            names.append(sub.name)
            values.append(sub.weight * loss if sub.weight != 1. else loss)
        out = {f'loss/{n}': v for n, v in zip(names, values)}
        if len(names) > 1:
            self.log_dict({n: v for n, v in zip(names, values)}, prog_bar=True, logger=False)
        out['loss'] = sum(values) if len(values) > 1 else values[0]
        return out

    def _calculate_metrics(self, img_sr, img_hr, dataloader_idx: int = 0):
        """reference srmodel.py:567-593."""
        out = {}
        for name, metric in self._metrics:
            value = metric(img_sr) if name in {'BRISQUE'} else metric(img_sr, img_hr)
            out[f'{self._eval_datasets[dataloader_idx]}/{name}'] = value
        pbar = {k: v for k, v in out.items() for m in self._metrics_for_pbar if m in k} or dict(out)
        self.log_dict(pbar, prog_bar=True, logger=False)
        return out

    def _parse_optimizer_config(self, optimizer: str, optimizer_params: list[str]):
        """reference srmodel.py:595-621.  NB: the reference shadows its `optimizer_params`
        argument (line 602) and therefore ALWAYS returns {} — every reference run uses the
        optimizer defaults (Adam lr=1e-3).  We parse the list as evidently intended; with the
        default `optimizer_params=[]` behaviour is identical."""
        if optimizer not in _supported_optimizers:
            raise ValueError(f'Optimizer not recognized: {optimizer}. Supported optimizers: '
                             f'{", ".join(_supported_optimizers)}')
        params = {}
        if optimizer_params:
            self._logger.warning('optimizer_params=%s are applied here; the reference shadows this argument '
                                 '(srmodel.py:602) and always trains with the optimizer defaults', optimizer_params)
        for item in optimizer_params:
            name, value = item.strip().split('=')
            name = name.strip()
            if name in ['eps', 'lr', 'lr_decay', 'weight_decay']:
                params[name] = float(value)
            elif name in ['betas']:
                params[name] = tuple(float(v) for v in value.split(','))
            else:
                params[name] = value
        return _supported_optimizers[optimizer], params
