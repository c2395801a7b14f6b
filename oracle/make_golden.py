"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes from
/root/reference (fp64, CPU) on deterministic synthetic weights/inputs (oracle/synth.py).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  Each fixture stores the case config, the seeds, the reference
output `sr` (fp32), the L1 loss (fp64), the input gradient, and for every trainable parameter
its gradient's L2 norm and its projection on a fixed synthetic probe vector (full gradients
for all parameters of RCAN would be 62 MB; a few small tensors are stored in full).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.ref_import import import_reference_models  # noqa: E402
from oracle.synth import synth_image_batch, synth_state_dict, synth_tensor  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name -> (class, ctor kwargs, input NCHW shape, weight gain)
CASES = {
    "edsr_base_x4":   ("EDSR", dict(n_feats=64, n_resblocks=16, res_scale=1.0, scale_factor=4), (2, 3, 24, 24), 0.577),
    "edsr_small_x2":  ("EDSR", dict(n_feats=64, n_resblocks=2, res_scale=1.0, scale_factor=2), (2, 3, 16, 24), 1.0),
    "edsr_small_x3":  ("EDSR", dict(n_feats=64, n_resblocks=2, res_scale=0.5, scale_factor=3), (1, 3, 16, 16), 1.0),
    "edsr_wide_x4":   ("EDSR", dict(n_feats=256, n_resblocks=4, res_scale=0.1, scale_factor=4), (1, 3, 16, 24), 0.577),
    "rcan_small_x4":  ("RCAN", dict(n_feats=64, n_resblocks=3, n_resgroups=2, reduction=16, scale_factor=4), (2, 3, 16, 16), 1.0),
    "rcan_full_x4":   ("RCAN", dict(n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16, scale_factor=4), (1, 3, 16, 16), 0.577),
    "rdn_b_x4":       ("RDN", dict(rdn_config="B", scale_factor=4), (1, 3, 16, 16), 0.577),
    "rdn_a_x2":       ("RDN", dict(rdn_config="A", scale_factor=2), (1, 3, 16, 16), 0.577),
    "srcnn_x2":       ("SRCNN", dict(scale_factor=2), (2, 3, 16, 16), 0.577),
    "wdsr_b_x4":      ("WDSR", dict(type="B", n_feats=64, n_resblocks=3, res_scale=1, scale_factor=4), (2, 3, 16, 24), 0.577),
    # input shape chosen so that no PReLU pre-activation of the fp64 run lies within 1e-6 of the kink (a (2,3,16,24) input has six
    # such values; fp32 summation-order noise then flips their slope and moves whole gradient tensors by 1e-3, run to run)
    "srresnet_x4":    ("SRResNet", dict(n_resblocks=3, n_feats=64, scale_factor=4), (1, 3, 16, 16), 0.577),
    "wdsr_a_x2":      ("WDSR", dict(type="A", n_feats=32, n_resblocks=2, res_scale=1, scale_factor=2), (1, 3, 16, 16), 0.577),
}

FULL_GRAD_MAX_ELEMS = 40_000


def frozen_entries(sd):
    return {k: v.detach().numpy() for k, v in sd.items() if k.startswith(("sub_mean", "add_mean"))}


def run_case(name, ref):
    cls_name, kwargs, xshape, gain = CASES[name]
    model = getattr(ref, cls_name)(**kwargs)
    sd0 = model.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd0.items()}
    sd = synth_state_dict(shapes, seed=0, gain=gain, frozen=frozen_entries(sd0))
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.double()
    scale = kwargs["scale_factor"]
    n, c, h, w = xshape
    x = torch.from_numpy(synth_image_batch(n, c, h, w, key=name + "/lr", seed=0)).double().requires_grad_(True)
    hr = torch.from_numpy(synth_image_batch(n, c, h * scale, w * scale, key=name + "/hr", seed=1)).double()
    sr = model.forward(x)
    loss = torch.nn.L1Loss()(sr, hr)          # srmodel.py:37,549
    loss.backward()
    out = {
        "config": np.frombuffer(json.dumps(dict(cls=cls_name, kwargs=kwargs, xshape=xshape, gain=gain,
                                                keys=list(shapes), shapes=[list(s) for s in shapes.values()])).encode(),
                                dtype=np.uint8),
        "sr": sr.detach().numpy().astype(np.float32),
        "loss": np.array(loss.item(), dtype=np.float64),
        "grad/input": x.grad.numpy().astype(np.float32),
    }
    gnames, gnorm, gproj = [], [], []
    for k, p in model.named_parameters():
        if not p.requires_grad:
            continue
        g = p.grad.detach().numpy().astype(np.float64)
        probe = synth_tensor(g.shape, "probe/" + k, seed=7).astype(np.float64)
        gnames.append(k)
        gnorm.append(np.sqrt((g * g).sum()))
        gproj.append((g * probe).sum())
        if g.size <= FULL_GRAD_MAX_ELEMS and (".3.conv_du" in k or k.startswith(("head", "tail.1", "SFENet1", "UPNet.4", "_net"))
                                              or k.endswith("bias") and g.size <= 64 and k.count(".") <= 3):
            out["grad/" + k] = g.astype(np.float32)
    out["grad_names"] = np.frombuffer(json.dumps(gnames).encode(), dtype=np.uint8)
    out["grad_norm"] = np.array(gnorm)
    out["grad_proj"] = np.array(gproj)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: loss={loss.item():.6f} sr[{sr.min().item():.3f},{sr.max().item():.3f}] "
          f"|g_in|={x.grad.norm().item():.3e} -> {os.path.getsize(path)/1024:.0f} KiB")


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = import_reference_models()
    names = sys.argv[1:] or list(CASES)
    for name in names:
        run_case(name, ref)


if __name__ == "__main__":
    main()
