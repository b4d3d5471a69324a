"""GPU: b2c_resized_crop_flip bit-exact against the goldens from the real Pillow / torchvision transform
of datasets.py:24-34, against the oracle on seeded cases, and the loader mirror end to end."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import resample_oracle as R

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment_ref.npz"))


@pytest.fixture(scope="module")
def D():
    from iclr_17_compression_b200 import datasets
    return datasets


def _photo(meta):
    h, w, seed, crc = (int(v) for v in meta)
    img = R.synth_photo(h, w, seed)
    assert zlib.crc32(img.tobytes()) == crc
    return img


def test_golden_batch_from_the_real_transform(D):
    """all golden cases as ONE ragged batch: same draws, same bytes"""
    imgs, params = [], []
    for n in range(int(G["n"])):
        img = _photo(G[f"img{n}"])
        torch.manual_seed(int(G[f"seed{n}"]))
        p = D.sample_params(img.shape[0], img.shape[1])
        assert [int(v) for v in p] == [int(v) for v in G[f"params{n}"]]
        imgs.append(img)
        params.append(p)
    y = D.resized_crop_flip(imgs, params, 256)
    assert y.shape == (len(imgs), 3, 256, 256) and y.dtype == torch.float32
    y8 = y.mul(255).round().to(torch.uint8).cpu()
    assert torch.equal(y8.float().div(255), y.cpu())              # ToTensor's exact uint8 / 255 values
    for n in range(len(imgs)):
        assert zlib.crc32(y8[n].numpy().tobytes()) == int(G[f"outcrc{n}"]), n
        if f"out{n}" in G:
            assert np.array_equal(y8[n].numpy(), G[f"out{n}"])


def test_golden_plain_resizes(D):
    for n in range(int(G["nresize"])):
        img = _photo(G[f"rs_in{n}"])
        oh, ow, crc = (int(v) for v in G[f"rs_outcrc{n}"])
        if oh != ow or oh % 32:
            continue
        y = D.resized_crop_flip([img], [(0, 0, img.shape[0], img.shape[1], False, False)], oh)
        y8 = y[0].mul(255).round().to(torch.uint8).permute(1, 2, 0).contiguous().cpu().numpy()
        assert zlib.crc32(y8.tobytes()) == crc, n


@pytest.mark.parametrize("size", [32, 64, 256, 512])
def test_against_oracle_ragged(D, size):
    g = np.random.default_rng(size)
    imgs, params = [], []
    for n in range(9):
        h, w = int(g.integers(1, 700)), int(g.integers(1, 900))
        img = R.synth_photo(h, w, 100 * size + n)
        ch, cw = int(g.integers(1, h + 1)), int(g.integers(1, w + 1))
        params.append((int(g.integers(0, h - ch + 1)), int(g.integers(0, w - cw + 1)), ch, cw, bool(g.integers(2)), bool(g.integers(2))))
        imgs.append(img)
    imgs.append(R.synth_photo(size, size, 7))                      # identity: both passes skipped by Pillow
    params.append((0, 0, size, size, True, True))
    y = D.resized_crop_flip(imgs, params, size).cpu()
    buf, desc, mh, mw = D.pack_images(imgs, params, crop_only=False)          # whole images resident, crop on the device
    y2 = D.resized_crop_flip_packed(buf.cuda(), desc.cuda(), size, mh, mw).cpu()
    assert torch.equal(y, y2)
    for n, (img, p) in enumerate(zip(imgs, params)):
        assert torch.equal(y[n], R.resized_crop_flip(img, *p, size=size)), (n, img.shape, p)


def test_large_downscale_and_full_size_property(D):
    """a 3000 x 4000 source (scale 15.6, 33 taps): against the oracle, and flips commute with the transform"""
    img = R.synth_photo(3000, 4000, 99)
    p = (5, 3, 2990, 3991)
    y = D.resized_crop_flip([img] * 4, [p + (False, False), p + (True, False), p + (False, True), p + (True, True)], 256)
    assert torch.equal(y[0].cpu(), R.resized_crop_flip(img, *p, False, False))
    assert torch.equal(y[1], y[0].flip(2)) and torch.equal(y[2], y[0].flip(1)) and torch.equal(y[3], y[0].flip(1, 2))


def test_bad_descriptors_fail_loudly(D):
    from iclr_17_compression_b200._lib import B2CError
    img = R.synth_photo(40, 40, 1)
    buf, desc, mh, mw = D.pack_images([img], [(0, 0, 40, 40, False, False)], crop_only=False)
    desc = desc.clone()
    desc[0, 5] = 41                                               # crop taller than the image
    with pytest.raises(B2CError):
        D.resized_crop_flip_packed(buf.cuda(), desc.cuda(), 32, 41, 40)
    with pytest.raises(B2CError):                                 # out_size not a multiple of 32
        D.resized_crop_flip_packed(buf.cuda(), desc.cuda(), 48, 40, 40)
    with pytest.raises(B2CError):                                 # shared memory: crop rows of 60000 pixels
        D.resized_crop_flip_packed(buf.cuda(), desc.cuda(), 32, 40, 60000)


def test_loader_matches_reference_compose(D, tmp_path):
    """get_train_loader over a directory of PNGs == the reference's Datasets + DataLoader with the same seeds"""
    Image = pytest.importorskip("PIL.Image")
    transforms = pytest.importorskip("torchvision.transforms")
    for n, (h, w) in enumerate([(300, 200), (256, 256), (120, 500), (640, 480), (90, 91)]):
        Image.fromarray(R.synth_photo(h, w, 40 + n), "RGB").save(tmp_path / f"im{n}.png")
    _, loader = D.get_train_loader(str(tmp_path), 256, 2)
    torch.manual_seed(77)
    got = [b.cpu() for b in loader]

    class Ref(torch.utils.data.Dataset):                          # datasets.py:14-37 restated with the real transforms
        paths = sorted(str(p) for p in tmp_path.glob("*.*"))
        tf = transforms.Compose([transforms.RandomResizedCrop(256), transforms.RandomHorizontalFlip(),
                                 transforms.RandomVerticalFlip(), transforms.ToTensor()])

        def __getitem__(self, i):
            return self.tf(Image.open(self.paths[i]).convert("RGB"))

        def __len__(self):
            return len(self.paths)

    torch.manual_seed(77)
    ref = list(torch.utils.data.DataLoader(Ref(), batch_size=2, shuffle=True))
    assert len(got) == len(ref) == 3
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
