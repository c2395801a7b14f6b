"""Lightning-free fit / validate / predict loops around the B200 path.

The reference drives its models with `LightningCLI` (/root/reference/main.py:87-93): Lightning's fit
loop calls `SRModel.training_step` (srmodel.py:160-171) + `configure_optimizers` (:145-154), the
validation loop calls `validation_step` (:214-232: clamp to [0,1], PSNR / SSIM per batch, mean over
the epoch :345-373), predict calls `predict_step` (:375-380).  Lightning is not installed in this
image, so the same three loops are provided here over plain iterables of `{'lr': ..., 'hr': ...}`
batches (what `srdata.py:120-133` yields).  Training goes through `srb200.trainer.TrainStep` (one CUDA
graph per step, flat fp32 parameters / Adam state, NCCL gradient all-reduce when a process group is
initialised); validation and prediction call the model's own `validation_step` / `predict_step`.

Checkpoints hold the model `state_dict` (reference layout: interchangeable with the reference's
`.ckpt['state_dict']`) plus the flat Adam moments and step count, so a resumed run continues
bit-identically (tests/test_runner_gpu.py).
"""
from __future__ import annotations

from typing import Iterable

import torch

from .trainer import TrainStep


class Runner:
    def __init__(self, model, lr_shape, scale: int | None = None, *, lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, use_graph: bool = True, process_group=None):
        """lr_shape: (N, C, h, w) of the training batches (fixed: the step is captured once)."""
        self.model = model
        self.scale = int(scale if scale is not None else model._scale_factor)
        self.lr_shape = tuple(lr_shape)
        # optimizer hyper-parameters the model was configured with (SRModel._parse_optimizer_config) win over the defaults
        op = dict(getattr(model, "_optim_params", None) or {})
        unknown = set(op) - {"lr", "betas", "eps", "weight_decay"}
        if unknown:
            raise NotImplementedError(f"optimizer_params {sorted(unknown)} are not supported by the fused Adam step")
        lr = float(op.get("lr", lr))
        betas = tuple(op.get("betas", betas))
        eps = float(op.get("eps", eps))
        weight_decay = float(op.get("weight_decay", weight_decay))
        self.step_runner = TrainStep(model, self.lr_shape, self.scale, lr=lr, betas=betas, eps=eps,
                                     weight_decay=weight_decay, use_graph=use_graph, process_group=process_group)
        self._captured = False
        self.global_step = 0

    @classmethod
    def from_config(cls, config, model: str, device="cuda", **overrides):
        """Builds model + Runner from a reference YAML (`configs/train_default_sr.yml`) or an equivalent dict.

        Honoured keys: `data.batch_size`, `data.patch_size` (the HR patch, srdata.py:57-80: LR = patch_size //
        scale_factor), `data.scale_factor`, and `model.init_args` (channels, losses, optimizer, metrics, ...;
        metrics this image cannot compute — piq's BRISQUE / LPIPS / MS-SSIM, FLIP — are dropped, PSNR / SSIM stay).
        `model` is the class name the reference passes as `--model` (main.py:87-93).  Trainer / logger / callback
        sections are Lightning glue and are ignored.  `overrides` are merged into the model's init args
        (e.g. n_resblocks=20) except `lr`, which goes to the optimizer step."""
        if isinstance(config, str):
            import yaml
            with open(config) as f:
                config = yaml.safe_load(f)
        import models
        data = dict(config.get("data") or {})
        init = dict((config.get("model") or {}).get("init_args") or {})
        lr = float(overrides.pop("lr", 1e-3))
        init.update(overrides)
        scale = int(init.pop("scale_factor", data.get("scale_factor", 4)))
        batch = int(init.pop("batch_size", data.get("batch_size", 16)))
        patch = int(init.pop("patch_size", data.get("patch_size", 128)))
        if "metrics" in init:
            init["metrics"] = [m for m in init["metrics"] if m in ("PSNR", "SSIM")] or ["PSNR"]
            init["metrics_for_pbar"] = [m for m in init.get("metrics_for_pbar", []) if m.split("/")[-1] in init["metrics"]]
        if not hasattr(models, model) or not isinstance(getattr(models, model), type):
            raise ValueError(f"unknown model {model!r}; available: {sorted(models.__all__)}")
        net = getattr(models, model)(scale_factor=scale, batch_size=batch, patch_size=patch, **init).to(device)
        channels = int(init.get("channels", 3))
        return cls(net, (batch, channels, patch // scale, patch // scale), scale, lr=lr)

    # ---- training -----------------------------------------------------------------------------
    def _ensure_captured(self, batch):
        if not self._captured:
            self.step_runner.load_batch(batch['lr'], batch['hr'])
            # warm-up steps would move the weights: capture on a copy of the state and restore it
            flat = self.step_runner.flat
            keep = [t.clone() for t in (flat.flat, flat.m, flat.v, flat.step_dev)]
            self.step_runner.capture()
            for dst, src in zip((flat.flat, flat.m, flat.v, flat.step_dev), keep):
                dst.copy_(src)
            self._captured = True

    def fit(self, batches: Iterable[dict], epochs: int = 1, on_step=None) -> list[float]:
        """Runs `epochs` passes over `batches`; returns the loss of every step.  `on_step(step, loss)`
        (optional) is called with the device loss tensor (reading it synchronises)."""
        losses = []
        pending = []
        def lookahead(it):
            """(batch, next batch or None)"""
            it = iter(it)
            try:
                cur = next(it)
            except StopIteration:
                return
            for nxt in it:
                yield cur, nxt
                cur = nxt
            yield cur, None
        for _ in range(epochs):
            for batch, nxt in lookahead(batches):
                if tuple(batch['lr'].shape) != self.lr_shape:
                    raise ValueError(f"batch shape {tuple(batch['lr'].shape)} differs from the captured {self.lr_shape}")
                self._ensure_captured(batch)
                pf = None
                if nxt is not None and not nxt['lr'].is_cuda and nxt['lr'].is_pinned() and nxt['hr'].is_pinned() \
                        and tuple(nxt['lr'].shape) == self.lr_shape:
                    pf = (nxt['lr'], nxt['hr'])                          # its H2D copy runs under this step
                loss = self.step_runner.step(batch['lr'], batch['hr'], prefetch=pf)
                self.global_step += 1
                pending.append(loss.clone())          # the graph overwrites its loss buffer every replay
                if on_step is not None:
                    on_step(self.global_step, pending[-1])
        torch.cuda.synchronize()
        losses.extend(float(t) for t in pending)
        return losses

    # ---- evaluation ---------------------------------------------------------------------------
    def _refresh_packed(self):
        """The captured step re-packs the bf16 / K-major weight copies BEFORE its forward, so after a step
        they are one Adam update behind the fp32 parameters: re-pack once before evaluating.  Packed copies that
        are not in the table (first used by an evaluation on another shape) are invalidated and re-pack lazily."""
        from . import ops
        ops.invalidate_packed()
        if self.step_runner.pack_table is not None:
            self.step_runner.pack_table.run()

    @torch.no_grad()
    def validate(self, batches: Iterable[dict]) -> dict:
        """Mean of `validation_step`'s metrics over the batches (srmodel.py:345-373)."""
        dev = self.step_runner.device
        self._refresh_packed()
        sums, count = {}, 0
        for i, batch in enumerate(batches):
            res = self.model.validation_step({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}, i)
            for k, v in res.items():
                sums[k] = sums.get(k, 0.0) + float(v)
            count += 1
        self.model.on_validation_epoch_end()
        return {k: v / max(count, 1) for k, v in sums.items()}

    @torch.no_grad()
    def predict(self, batches: Iterable[dict]) -> list[torch.Tensor]:
        dev = self.step_runner.device
        self._refresh_packed()
        return [self.model.predict_step({'lr': b['lr'].to(dev)}, i).cpu() for i, b in enumerate(batches)]

    # ---- checkpoint / resume ------------------------------------------------------------------
    def state(self) -> dict:
        flat = self.step_runner.flat
        return {"state_dict": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
                "adam_m": flat.m.cpu().clone(), "adam_v": flat.v.cpu().clone(),
                "adam_step": int(flat.step_dev.item()), "global_step": self.global_step}

    def save_checkpoint(self, path: str):
        torch.save(self.state(), path)

    def load_checkpoint(self, path_or_state):
        st = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, str) else path_or_state
        self.model.load_state_dict(st["state_dict"])            # parameters are views of the flat buffer: copied in place
        flat = self.step_runner.flat
        if "adam_m" in st:
            flat.m.copy_(st["adam_m"])
            flat.v.copy_(st["adam_v"])
            flat.step_dev.fill_(int(st["adam_step"]))
        self.global_step = int(st.get("global_step", 0))
        from . import ops
        ops.invalidate_packed()

    def close(self):
        self.step_runner.close()
