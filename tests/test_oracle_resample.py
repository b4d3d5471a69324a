"""CPU: the input-transform oracle (oracle/resample_oracle.py) against golden vectors made with the
real Pillow / torchvision the reference calls (tests/golden/make_augment_golden.py), against the live packages when
importable, and the host-side draws / packing of iclr_17_compression_b200.datasets."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import resample_oracle as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment_ref.npz"))


def _photo(meta):
    h, w, seed, crc = (int(v) for v in meta)
    img = R.synth_photo(h, w, seed)
    assert zlib.crc32(img.tobytes()) == crc, "synthetic source image drifted: regenerate the goldens"
    return img


@pytest.mark.parametrize("n", range(int(G["nresize"])))
def test_resize_matches_pillow_golden(n):
    img = _photo(G[f"rs_in{n}"])
    oh, ow, crc = (int(v) for v in G[f"rs_outcrc{n}"])
    out = R.resize_bilinear_u8(img, oh, ow)
    assert zlib.crc32(np.ascontiguousarray(out).tobytes()) == crc
    if f"rs_out{n}" in G:
        assert np.array_equal(out, G[f"rs_out{n}"])


@pytest.mark.parametrize("n", range(int(G["n"])))
def test_transform_matches_torchvision_golden(n):
    img = _photo(G[f"img{n}"])
    torch.manual_seed(int(G[f"seed{n}"]))
    p = R.sample_params(img.shape[0], img.shape[1])
    assert [int(v) for v in p] == [int(v) for v in G[f"params{n}"]]
    y = R.resized_crop_flip(img, *p)
    y8 = y.mul(255).round().to(torch.uint8).numpy()
    assert torch.equal(torch.from_numpy(y8).float().div(255), y)
    assert zlib.crc32(y8.tobytes()) == int(G[f"outcrc{n}"])
    if f"out{n}" in G:
        assert np.array_equal(y8, G[f"out{n}"])


def test_oracle_against_live_packages():
    Image = pytest.importorskip("PIL.Image")
    transforms = pytest.importorskip("torchvision.transforms")
    tf = transforms.Compose([transforms.RandomResizedCrop(64), transforms.RandomHorizontalFlip(),
                             transforms.RandomVerticalFlip(), transforms.ToTensor()])
    for seed, (h, w) in enumerate([(50, 70), (200, 130), (64, 64), (333, 517)]):
        img = R.synth_photo(h, w, 900 + seed)
        torch.manual_seed(seed)
        ref = tf(Image.fromarray(img, "RGB"))
        torch.manual_seed(seed)
        p = R.sample_params(h, w)
        assert torch.equal(R.resized_crop_flip(img, *p, size=64), ref)


def test_host_draws_match_oracle_and_consume_the_same_stream():
    from iclr_17_compression_b200 import datasets as D
    for seed, (h, w) in enumerate([(300, 451), (97, 1031), (2100, 260), (16, 16), (5, 400)]):
        torch.manual_seed(seed)
        a = [D.sample_params(h, w) for _ in range(5)] + [torch.rand(1).item()]
        torch.manual_seed(seed)
        b = [R.sample_params(h, w) for _ in range(5)] + [torch.rand(1).item()]
        assert a == b


def test_pack_images_layout():
    from iclr_17_compression_b200 import datasets as D
    imgs = [R.synth_photo(7, 5, 1), R.synth_photo(33, 21, 2)]
    buf, desc, mh, mw = D.pack_images(imgs, [(1, 0, 5, 4, True, False), (0, 3, 33, 10, False, True)], crop_only=False)
    assert desc.tolist() == [[0, 7, 5, 1, 0, 5, 4, 1], [112, 33, 21, 0, 3, 33, 10, 2]]
    assert (mh, mw) == (33, 10) and buf.numel() == 112 + (33 * 21 * 3 + 15) // 16 * 16
    assert np.array_equal(buf[112:112 + 33 * 21 * 3].numpy().reshape(33, 21, 3), imgs[1])
    buf, desc, mh, mw = D.pack_images(imgs, [(1, 0, 5, 4, True, False), (0, 3, 33, 10, False, True)])     # crops only
    assert desc.tolist() == [[0, 5, 4, 0, 0, 5, 4, 1], [64, 33, 10, 0, 0, 33, 10, 2]] and buf.numel() == 64 + 992
    assert np.array_equal(buf[64:64 + 990].numpy().reshape(33, 10, 3), imgs[1][:, 3:13])
    with pytest.raises(ValueError):
        D.pack_images(imgs, [(1, 0, 7, 4, False, False), (0, 0, 1, 1, False, False)])
    with pytest.raises(ValueError):
        D.pack_images([np.zeros((4, 4, 3), np.float32)], [(0, 0, 1, 1, False, False)])


def test_datasets_mirror_on_cpu(tmp_path):
    """host side of the loader mirror without a GPU: items are (decoded uint8 image, draws in torchvision's RNG order);
    the device call refuses CPU tensors (no CPU path)"""
    Image = pytest.importorskip("PIL.Image")
    from iclr_17_compression_b200 import datasets as D
    from iclr_17_compression_b200._lib import B2CError
    for n, (h, w) in enumerate([(40, 60), (33, 33)]):
        Image.fromarray(R.synth_photo(h, w, n), "RGB").save(tmp_path / f"im{n}.png")
    ds = D.Datasets(str(tmp_path), 32)
    assert len(ds) == 2 and len(D.TestKodakDataset(str(tmp_path))) == 2
    # datasets.py:63-80: TestKodakDataset items are ToTensor() outputs - fp32 [3, H, W] in [0, 1]
    from torchvision import transforms
    kd = D.TestKodakDataset(str(tmp_path))
    ref_item = transforms.ToTensor()(Image.open(tmp_path / "im0.png").convert("RGB"))
    assert kd[0].dtype == torch.float32 and tuple(kd[0].shape) == (3, 40, 60) and torch.equal(kd[0], ref_item)
    raw = D.TestKodakDataset(str(tmp_path), raw=True)[0]
    assert raw.dtype == torch.uint8 and torch.equal(raw.float().div(255), ref_item)
    batch = next(iter(torch.utils.data.DataLoader(D.TestKodakDataset(str(tmp_path)), batch_size=1)))
    assert tuple(batch.shape) == (1, 3, 40, 60) and batch.dtype == torch.float32
    torch.manual_seed(5)
    img, p = ds[0]
    torch.manual_seed(5)
    assert p == R.sample_params(40, 60) and img.dtype == torch.uint8 and tuple(img.shape) == (40, 60, 3)
    assert np.array_equal(img.numpy(), R.synth_photo(40, 60, 0))
    with pytest.raises(Exception):
        D.Datasets(str(tmp_path / "missing"))
    buf, desc, mh, mw = D.pack_images([img], [p])
    with pytest.raises(B2CError):
        D.resized_crop_flip_packed(buf, desc, 32, mh, mw)
