"""GPU data path (csrc/data.cu + srb200/data.py) — the reference's per-sample training data path
(srdata.py:57-92,136-169: crop, rotate, flips, to_tensor).

CPU: the numpy restatement (oracle/data_oracle.py) is pinned bit for bit against the calls the reference itself makes
(torchvision.transforms.functional on PIL images, driven by Python's `random` exactly as srdata.py does).
GPU: `srb_patch_batch` against that restatement, bit for bit, for every angle / flip combination, non-square images, and crop
boxes that leave the image (the reference's (w, h) mix-up makes that reachable)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sr-pytorch-lightning_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import data_oracle  # noqa: E402


def _images(seed=0, sizes=((60, 72), (96, 64), (48, 48), (50, 130))):
    """[(lr uint8 [h,w,3], hr uint8 [4h,4w,3])] random images (scale 4)."""
    rs = np.random.RandomState(seed)
    return [(rs.randint(0, 256, (h, w, 3), dtype=np.uint8), rs.randint(0, 256, (4 * h, 4 * w, 3), dtype=np.uint8)) for h, w in sizes]


def _reference_item(lr_pil, hr_pil, patch_size, scale):
    """srdata.py:57-92,136-169 + :125-128 verbatim in behaviour, with the module-level `random` the reference uses."""
    import torchvision.transforms.functional as TF
    lr_patch_size = patch_size // scale
    lr_h, lr_w = lr_pil.size                                        # sic (srdata.py:152-153)
    lr_x = random.randrange(0, lr_h - lr_patch_size + 1)
    lr_y = random.randrange(0, lr_w - lr_patch_size + 1)
    lr = TF.crop(lr_pil, lr_x, lr_y, lr_patch_size, lr_patch_size)
    hr = TF.crop(hr_pil, scale * lr_x, scale * lr_y, patch_size, patch_size)
    angle = random.choice((0, 90, 180, 270))
    if angle != 0:
        hr, lr = TF.rotate(hr, angle=angle), TF.rotate(lr, angle=angle)
    if random.choice((True, False)):
        hr, lr = TF.hflip(hr), TF.hflip(lr)
    if random.choice((True, False)):
        hr, lr = TF.vflip(hr), TF.vflip(lr)
    return TF.to_tensor(lr).numpy(), TF.to_tensor(hr).numpy()


def test_oracle_matches_torchvision_on_pil_images():
    Image = pytest.importorskip("PIL.Image")
    pytest.importorskip("torchvision")
    scale, lr_patch = 4, 24
    for k, (lr, hr) in enumerate(_images()):
        lr_pil, hr_pil = Image.fromarray(lr), Image.fromarray(hr)
        for seed in range(12):
            random.seed(1000 * k + seed)
            want_lr, want_hr = _reference_item(lr_pil, hr_pil, lr_patch * scale, scale)
            rng = random.Random(1000 * k + seed)
            choice = data_oracle.draw(rng, lr_pil.size, lr_patch)
            got_lr, got_hr = data_oracle.get_item(lr, hr, choice, lr_patch, scale)
            assert np.array_equal(got_lr, want_lr) and np.array_equal(got_hr, want_hr), (k, seed, choice)


def test_oracle_every_augmentation_against_torchvision():
    Image = pytest.importorskip("PIL.Image")
    TF = pytest.importorskip("torchvision.transforms.functional")
    lr, _ = _images(3, sizes=((40, 40),))[0]
    pil = Image.fromarray(lr)
    for angle in (0, 90, 180, 270):
        for hflip in (False, True):
            for vflip in (False, True):
                for top, left in ((0, 0), (7, 3), (30, 25), (-5, 33)):      # the last two boxes leave the image
                    want = TF.crop(pil, top, left, 16, 16)
                    if angle:
                        want = TF.rotate(want, angle=angle)
                    if hflip:
                        want = TF.hflip(want)
                    if vflip:
                        want = TF.vflip(want)
                    got = data_oracle.to_tensor(data_oracle.augment(data_oracle.crop(lr, top, left, 16), angle, hflip, vflip))
                    assert np.array_equal(got, TF.to_tensor(want).numpy()), (angle, hflip, vflip, top, left)


def test_sampler_draws_what_the_reference_draws():
    """PatchSampler.draw consumes Python's random stream exactly as srdata.py:165-166,77-92 does."""
    rs = random.Random(5)
    want = [data_oracle.draw(rs, (72, 60), 24) for _ in range(20)]

    class Fake:      # draw() only needs the LR image's shape
        shape = (60, 72, 3)
    from srb200.data import PatchSampler
    s = PatchSampler.__new__(PatchSampler)
    s.rng, s.lr_patch, s.augment, s.images = random.Random(5), 24, True, [(Fake(), None)]
    assert [s.draw(0) for _ in range(20)] == want


@pytest.mark.gpu
def test_patch_batch_kernel_bit_exact():
    from srb200.data import PatchSampler
    scale, lr_patch, n = 4, 24, 64
    imgs = _images(7)
    s = PatchSampler(scale, lr_patch, device="cuda:0", seed=11)
    for lr, hr in imgs:
        s.add(hr, lr)
    lr_out = torch.empty(n, 3, lr_patch, lr_patch, device="cuda:0")
    hr_out = torch.empty(n, 3, lr_patch * scale, lr_patch * scale, device="cuda:0")
    made = s.fill(lr_out, hr_out)
    assert {c[2] for _, c in made} == {0, 90, 180, 270} and {c[3] for _, c in made} == {True, False}
    lr_h, hr_h = lr_out.cpu().numpy(), hr_out.cpu().numpy()
    for k, (i, choice) in enumerate(made):
        want_lr, want_hr = data_oracle.get_item(imgs[i][0], imgs[i][1], choice, lr_patch, scale)
        assert np.array_equal(lr_h[k], want_lr) and np.array_equal(hr_h[k], want_hr), (k, i, choice)
    # boxes that leave the image, every augmentation, explicitly
    choices = [(top, left, a, h, v) for a in (0, 90, 180, 270) for h in (False, True) for v in (False, True)
               for top, left in ((-3, 2), (40, 60), (0, 0), (26, 110))]
    idx = [3] * len(choices)
    lr2 = torch.empty(len(choices), 3, lr_patch, lr_patch, device="cuda:0")
    hr2 = torch.empty(len(choices), 3, lr_patch * scale, lr_patch * scale, device="cuda:0")
    s.fill(lr2, hr2, indices=idx, choices=choices)
    for k, ch in enumerate(choices):
        want_lr, want_hr = data_oracle.get_item(imgs[3][0], imgs[3][1], ch, lr_patch, scale)
        assert np.array_equal(lr2[k].cpu().numpy(), want_lr) and np.array_equal(hr2[k].cpu().numpy(), want_hr), ch


@pytest.mark.gpu
def test_sampler_makes_the_lr_image_as_the_reference_and_feeds_a_training_step():
    """HR-only images: the LR image is PIL's antialiased bicubic of the whole image (srdata.py:226-229); the batch goes
    straight into TrainStep's static buffers (no host batch)."""
    pytest.importorskip("PIL.Image")
    from sklearn.datasets import load_sample_image
    import torchvision.transforms.functional as TF
    from torchvision.transforms import InterpolationMode
    from PIL import Image
    import models
    from srb200.data import PatchSampler
    from srb200.trainer import TrainStep
    s = PatchSampler(4, 48, device="cuda:0", seed=3)
    for name in ("china.jpg", "flower.jpg"):
        hr = load_sample_image(name)[:424, :640].copy()
        i = s.add(hr)
        pil = Image.fromarray(hr)
        want = np.asarray(TF.resize(pil, [106, 160], interpolation=InterpolationMode.BICUBIC))
        assert np.array_equal(s.images[i][0].cpu().numpy(), want)
    torch.manual_seed(0)
    m = models.EDSR(n_feats=64, n_resblocks=2, scale_factor=4)
    m.compute_dtype = "bf16"
    ts = TrainStep(m.cuda(), (8, 3, 48, 48), 4, lr=1e-4)
    ts.prepare()
    losses = []
    for _ in range(6):
        s.fill(ts.x, ts.hr)
        losses.append(float(ts.run().item()))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    ts.close()
