"""Drop-in boundary (CPU): state_dict layout, constructor/error behaviour of the SRModel plugin
API, and the C-ABI library exporting every symbol include/srb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from golden_util import Golden, golden_names
from oracle.ref_import import import_reference_models, reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_layout_matches_golden(name):
    """Keys, order and shapes recorded from the reference classes when the fixtures were made."""
    import models
    g = Golden(name)
    m = getattr(models, g.cls)(**g.kwargs)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g.shapes.keys())
    for k, v in sd.items():
        want = torch.int64 if k.endswith("num_batches_tracked") else torch.float32       # (BatchNorm's step counter, SRResNet)
        assert tuple(v.shape) == g.shapes[k] and v.dtype == want, k
    trainable = [k for k, p in m.named_parameters() if p.requires_grad]
    assert trainable == g.grad_names
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g.state_dict().items()})


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this box")
@pytest.mark.parametrize("cls,kw", [
    ("EDSR", dict(n_feats=256, n_resblocks=32, res_scale=0.1, scale_factor=4)),
    ("EDSR", dict(scale_factor=8)),
    ("RCAN", dict(n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16, scale_factor=4)),
    ("RDN", dict(rdn_config="B", scale_factor=3)),
    ("SRCNN", dict(scale_factor=3)),
    ("WDSR", dict(type="B", n_feats=128, n_resblocks=16, scale_factor=4)),
    ("WDSR", dict(type="A", n_feats=32, n_resblocks=3, scale_factor=2)),
    ("SRResNet", dict(n_resblocks=16, n_feats=64, scale_factor=4)),
    ("SRResNet", dict(n_resblocks=2, n_feats=64, scale_factor=3)),
])
def test_state_dict_layout_matches_reference(cls, kw):
    import models
    ref = import_reference_models()
    a, b = getattr(models, cls)(**kw), getattr(ref, cls)(**kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    assert all(sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype for k in sa)
    assert [k for k, p in a.named_parameters() if p.requires_grad] == [k for k, p in b.named_parameters() if p.requires_grad]
    a.load_state_dict(sb)          # checkpoints interchange in both directions
    b.load_state_dict(a.state_dict())
    # frozen MeanShift values
    for k in sa:
        if k.startswith(("sub_mean", "add_mean")):
            assert torch.equal(sa[k], sb[k]), k
    assert a.example_input_array.shape == b.example_input_array.shape


def test_registry_and_signatures():
    import inspect
    import models
    assert set(models.__all__) == {"EDSR", "RCAN", "RDN", "SRCNN", "SRModel", "SRResNet", "WDSR"}
    for cls in ("EDSR", "RCAN", "RDN", "SRCNN", "SRResNet", "WDSR"):
        assert issubclass(getattr(models, cls), models.SRModel)
    assert list(inspect.signature(models.EDSR.__init__).parameters)[1:4] == ["n_feats", "n_resblocks", "res_scale"]
    assert list(inspect.signature(models.RCAN.__init__).parameters)[1:6] == ["n_feats", "n_resblocks", "n_resgroups", "reduction", "res_scale"]
    assert list(inspect.signature(models.RDN.__init__).parameters)[1:4] == ["rdn_config", "G0", "kernel_size"]
    assert list(inspect.signature(models.WDSR.__init__).parameters)[1:5] == ["type", "n_feats", "n_resblocks", "res_scale"]      # wdsr.py:59
    m = models.RCAN()
    assert len(m.body) == 11 and len(m.body[0].body) == 17        # reference defaults (rcan.py:82)
    assert "compute_dtype" not in m.hparams and m.hparams["scale_factor"] == 4


def test_error_behaviour():
    import models
    with pytest.raises(ValueError, match="scale must be 2 or 3 or 4"):
        models.RDN(scale_factor=8)                                  # rdn.py:97
    with pytest.raises(AssertionError):
        models.EDSR(scale_factor=5)                                 # common.py:125
    with pytest.raises(AttributeError, match="Couldn't find loss"):
        models.EDSR(losses="nope")                                  # srmodel.py:483
    with pytest.raises(ValueError, match="not a valid number"):
        models.EDSR(losses="x*l1")                                  # srmodel.py:444-448
    with pytest.raises(AttributeError, match="Couldn't find metric"):
        models.EDSR(metrics=["nope"])                               # srmodel.py:514
    with pytest.raises(ValueError, match="Optimizer not recognized"):
        models.EDSR(optimizer="LION")                               # srmodel.py:599
    with pytest.raises(KeyError):
        models.RDN(rdn_config="C")                                  # rdn.py:51-54
    with pytest.raises(ValueError):
        models.EDSR().compute_dtype = "fp16"
    m = models.EDSR(losses="0.5*l1+0.5*l2", optimizer="SGD", optimizer_params=["lr=0.1"])
    assert [s.weight for s in m._losses] == [0.5, 0.5]
    opt = m.configure_optimizers()[0]
    assert isinstance(opt, torch.optim.SGD) and opt.param_groups[0]["lr"] == 0.1
    # MeanShift parameters are frozen and never reach the optimizer (common.py:70-71, srmodel.py:151-152)
    ids = {id(p) for g in opt.param_groups for p in g["params"]}
    assert id(m.sub_mean.bias) not in ids and id(m.head[0].weight) in ids


def test_product_path_has_no_cpu_fallback(monkeypatch):
    """The models compute only through libsrb200: CPU tensors are refused, and a missing library is an error
    at load time (never a silent PyTorch/oracle fallback)."""
    import models
    from srb200 import lib
    for cls, kw in (("EDSR", dict(n_resblocks=1)), ("RCAN", dict(n_resblocks=1, n_resgroups=1)), ("RDN", {}), ("SRCNN", dict(scale_factor=2)),
                    ("WDSR", dict(n_feats=16, n_resblocks=1))):
        m = getattr(models, cls)(**kw)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(torch.rand(1, 3, 8, 8))
    src = open(os.path.join(ROOT, "sr-pytorch-lightning_b200", "srb200", "ops.py")).read() + \
        open(os.path.join(ROOT, "sr-pytorch-lightning_b200", "srb200", "functional.py")).read()
    assert "oracle" not in src                                  # the product never imports the test oracle
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libsrb200.so")
    monkeypatch.setattr(lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.load()


def test_cabi_exports_every_declared_symbol():
    """The shared library loads and exports exactly the entry points the header declares."""
    from srb200 import lib
    header = open(os.path.join(ROOT, "include", "srb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(srb_[a-z0-9_]+)\s*\(", header))
    declared -= {"srb_ctx"}
    assert {"srb_conv", "srb_conv_wgrad", "srb_conv_wgrad_batched", "srb_ca_fwd", "srb_ca_bwd", "srb_pack_table",
            "srb_l1_loss", "srb_adam_step", "srb_create"} <= declared
    handle = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, missing
    assert declared == set(lib.PROTOTYPES), (declared ^ set(lib.PROTOTYPES))
    loaded = lib.load()
    assert loaded.srb_abi_version() == 1
    assert ctypes.sizeof(lib.ConvDesc) == 22 * 4 and ctypes.sizeof(lib.WgradDesc) == 15 * 4
    assert ctypes.sizeof(lib.PackItem) == 40


def test_psnr_ssim_metrics_cpu():
    from models.srmodel import psnr, ssim
    a = torch.rand(2, 3, 32, 32)
    assert psnr(a, a).item() == pytest.approx(80.0, abs=1e-4)
    assert ssim(a, a).item() == pytest.approx(1.0, abs=1e-5)
    b = (a + 0.05 * torch.randn_like(a)).clamp(0, 1)
    assert 15 < psnr(a, b).item() < 40 and 0.3 < ssim(a, b).item() < 1.0


def _run(code, cwd=None):
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=cwd, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout


def test_integration_option_a_launcher_shadows_the_models_package(tmp_path):
    """scripts/run_reference_main.py run from a checkout makes `import models` inside main.py resolve to
    the B200 classes while sibling modules still come from the checkout (a stand-in checkout: Lightning is not installed here)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (tmp_path / "models").mkdir()
    (tmp_path / "models" / "__init__.py").write_text("raise ImportError('the checkout\\'s own models package was imported')\n")
    (tmp_path / "srdata.py").write_text("WHO = 'checkout'\n")
    (tmp_path / "main.py").write_text(
        "import sys\nimport models\nfrom srdata import WHO\n"
        "assert 'sr-pytorch-lightning_b200' in models.__file__, models.__file__\n"
        "cls = getattr(models, sys.argv[2])\nassert issubclass(cls, models.SRModel)\n"
        "m = cls(n_resblocks=1, n_resgroups=1)\nprint('OK', WHO, cls.__name__, len(m.state_dict()))\n")
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "run_reference_main.py"), "fit", "RCAN"], capture_output=True,
                       text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0 and "OK checkout RCAN" in r.stdout, r.stdout + r.stderr


def test_integration_option_b_registry_snippet_runs(tmp_path):
    """the registry edit, executed verbatim inside a stand-in `models/__init__.py`."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "sr-pytorch-lightning_b200")
    (tmp_path / "models").mkdir()
    (tmp_path / "models" / "__init__.py").write_text(
        "import importlib.util, sys\n"
        f"_B200 = {pkg!r}\n"
        "sys.path.insert(0, _B200)\n"
        "_spec = importlib.util.spec_from_file_location('models_b200', _B200 + '/models/__init__.py', submodule_search_locations=[_B200 + '/models'])\n"
        "models_b200 = importlib.util.module_from_spec(_spec)\n"
        "sys.modules['models_b200'] = models_b200\n"
        "_spec.loader.exec_module(models_b200)\n"
        "from models_b200 import EDSR, RCAN, RDN, SRCNN, SRResNet, WDSR, SRModel\n")
    out = _run("import models; m = models.EDSR(n_resblocks=1); assert issubclass(models.EDSR, models.SRModel); "
               "print('OK', type(m).__module__, len(m.state_dict()))", cwd=str(tmp_path))
    assert "OK models_b200.edsr" in out
