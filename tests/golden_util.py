"""Helpers shared by the tests: load golden fixtures, rebuild their synthetic weights/inputs."""
import json
import os

import numpy as np

from oracle.synth import synth_image_batch, synth_state_dict, synth_tensor

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RGB_MEAN = (0.4488, 0.4371, 0.4040)


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def frozen_meanshift(shapes):
    out = {}
    for k in shapes:
        if k.startswith(("sub_mean", "add_mean")):
            sign = -1.0 if k.startswith("sub_mean") else 1.0
            if k.endswith("weight"):
                out[k] = np.eye(3, dtype=np.float32).reshape(3, 3, 1, 1)
            else:
                out[k] = (sign * np.array(RGB_MEAN, dtype=np.float64)).astype(np.float32)
    return out


class Golden:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        cfg = json.loads(bytes(z["config"]).decode())
        self.cls = cfg["cls"]
        self.kwargs = cfg["kwargs"]
        self.xshape = tuple(cfg["xshape"])
        self.gain = cfg["gain"]
        self.shapes = {k: tuple(s) for k, s in zip(cfg["keys"], cfg["shapes"])}
        self.scale = self.kwargs["scale_factor"]
        self.sr = z["sr"]
        self.loss = float(z["loss"])
        self.grad_names = json.loads(bytes(z["grad_names"]).decode())
        self.grad_norm = z["grad_norm"]
        self.grad_proj = z["grad_proj"]

    def state_dict(self):
        return synth_state_dict(self.shapes, seed=0, gain=self.gain, frozen=frozen_meanshift(self.shapes))

    def inputs(self):
        n, c, h, w = self.xshape
        x = synth_image_batch(n, c, h, w, key=self.name + "/lr", seed=0)
        hr = synth_image_batch(n, c, h * self.scale, w * self.scale, key=self.name + "/hr", seed=1)
        return x, hr

    def oracle_cfg(self):
        k = dict(self.kwargs)
        cfg = {"scale": k.pop("scale_factor")}
        if self.cls == "EDSR":
            cfg.update(n_resblocks=k["n_resblocks"], res_scale=k["res_scale"])
        elif self.cls == "RCAN":
            cfg.update(n_resblocks=k["n_resblocks"], n_resgroups=k["n_resgroups"])
        elif self.cls == "RDN":
            cfg.update(rdn_config=k["rdn_config"])
        elif self.cls == "SRResNet":
            cfg.update(n_resblocks=k["n_resblocks"], n_feats=k["n_feats"])
        elif self.cls == "WDSR":
            cfg.update(type=k["type"], n_feats=k["n_feats"], n_resblocks=k["n_resblocks"], res_scale=k["res_scale"])
        return cfg

    def full_grads(self):
        return {k[5:]: self.z[k] for k in self.z.files if k.startswith("grad/")}

    def probe(self, name, shape):
        return synth_tensor(shape, "probe/" + name, seed=7).astype(np.float64)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
