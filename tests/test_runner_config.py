"""Runner.from_config: the keys of the reference's YAML configs (configs/train_default_sr.yml) that concern the
hot path are honoured; Lightning glue is ignored.  Construction only (no kernel runs): CPU."""
import pytest
import torch

CFG = {
    "data": {"augment": True, "batch_size": 16, "datasets_dir": "/datasets", "eval_datasets": ["B100", "DIV2K"],
             "patch_size": 128, "scale_factor": 4, "train_datasets": ["DIV2K"]},
    "model": {"init_args": {"channels": 3, "log_loss_every_n_epochs": 50, "losses": "l1",
                            "metrics": ["BRISQUE", "FLIP", "LPIPS", "MS-SSIM", "PSNR", "SSIM"],
                            "metrics_for_pbar": ["DIV2K/PSNR", "DIV2K/SSIM"], "optimizer": "ADAM", "save_results": -1,
                            "save_results_from_epoch": "last"}},
    "trainer": {"max_epochs": 2000, "check_val_every_n_epoch": 200,
                "logger": [{"class_path": "lightning.pytorch.loggers.CometLogger"}]},
}


def test_from_config_builds_model_and_step_shapes(tmp_path):
    import yaml
    from srb200.runner import Runner
    path = tmp_path / "train.yml"
    path.write_text(yaml.safe_dump(CFG))
    r = Runner.from_config(str(path), "RCAN", device="cpu", n_resblocks=2, n_resgroups=2, lr=1e-4)
    m = r.model
    assert type(m).__name__ == "RCAN" and len(m.body) == 3 and len(m.body[0].body) == 3
    assert r.lr_shape == (16, 3, 32, 32) and r.scale == 4                 # LR patch = patch_size // scale_factor
    assert r.step_runner.hp["lr"] == 1e-4
    assert [n for n, _ in m._metrics] == ["PSNR", "SSIM"]                 # piq-only metrics are dropped
    assert m._metrics_for_pbar == ["DIV2K/PSNR", "DIV2K/SSIM"]
    assert tuple(r.step_runner.hr.shape) == (16, 3, 128, 128)
    # parameters were re-homed into the flat buffer without changing the state_dict layout
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) <= r.step_runner.flat.numel
    assert list(m.state_dict())[0] == "sub_mean.weight"


def test_from_config_dict_and_errors():
    from srb200.runner import Runner
    r = Runner.from_config({"data": {"batch_size": 4, "patch_size": 96, "scale_factor": 2}}, "EDSR", device="cpu",
                           n_resblocks=1)
    assert r.lr_shape == (4, 3, 48, 48) and r.scale == 2
    with pytest.raises(ValueError, match="unknown model"):
        Runner.from_config(CFG, "SRGAN", device="cpu")
    with pytest.raises(ValueError, match="batch shape"):
        r.fit([{"lr": torch.zeros(2, 3, 48, 48), "hr": torch.zeros(2, 3, 96, 96)}])
