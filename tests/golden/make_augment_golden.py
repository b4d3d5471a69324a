"""Golden vectors for the training input transform from the REAL packages the reference calls
(datasets.py:24-34): torchvision's Compose([RandomResizedCrop(256), RandomHorizontalFlip(), RandomVerticalFlip(),
ToTensor()]) on PIL images, run in this container (Pillow / torchvision versions stored in the file).

    python tests/golden/make_augment_golden.py        # writes tests/golden/augment_ref.npz

Stored per case: the geometry and seed of the deterministic source image (oracle.resample_oracle.synth_photo) with
its crc32, the torch seed, the crop box and flips the real transform drew (recorded by re-running the draws with the
same seed), the crc32 of the transform's output as uint8 (ToTensor output * 255 is exact), and for the small cases
the output itself, so the fixture stays small."""
import os
import sys
import zlib

import numpy as np
import PIL
import torch
import torchvision
from PIL import Image
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import resample_oracle as R  # noqa: E402  (only for the deterministic synthetic photos)

CASES = [(64, 96, 11), (300, 451, 12), (256, 256, 13), (700, 513, 14), (1100, 1700, 15), (97, 1031, 16), (2100, 260, 17)]


def main():
    out = {"pillow": np.array(PIL.__version__), "torchvision": np.array(torchvision.__version__), "n": np.array(len(CASES))}
    tf = transforms.Compose([transforms.RandomResizedCrop(256), transforms.RandomHorizontalFlip(),
                             transforms.RandomVerticalFlip(), transforms.ToTensor()])
    for n, (h, w, seed) in enumerate(CASES):
        img = R.synth_photo(h, w, seed)
        torch.manual_seed(1000 + seed)
        y = tf(Image.fromarray(img, "RGB"))
        torch.manual_seed(1000 + seed)
        i, j, ch, cw = transforms.RandomResizedCrop.get_params(Image.fromarray(img, "RGB"), (0.08, 1.0), (3 / 4, 4 / 3))
        hf, vf = bool(torch.rand(1) < 0.5), bool(torch.rand(1) < 0.5)
        y8 = (y * 255).round().to(torch.uint8)
        assert torch.equal(y8.float().div(255), y)
        out[f"img{n}"] = np.array([h, w, seed, zlib.crc32(img.tobytes())])
        out[f"seed{n}"] = np.array(1000 + seed)
        out[f"params{n}"] = np.array([i, j, ch, cw, int(hf), int(vf)])
        out[f"outcrc{n}"] = np.array(zlib.crc32(y8.numpy().tobytes()))
        if n < 3:
            out[f"out{n}"] = y8.numpy()
    # plain resizes (no crop) at awkward ratios, incl. upscaling and a skipped pass
    sizes = [(64, 96, 256, 256), (300, 451, 256, 256), (1100, 1700, 256, 256), (256, 999, 256, 256), (513, 256, 256, 256),
             (37, 41, 64, 48), (4000, 300, 256, 256)]
    out["nresize"] = np.array(len(sizes))
    for n, (h, w, oh, ow) in enumerate(sizes):
        img = R.synth_photo(h, w, 50 + n)
        res = np.asarray(Image.fromarray(img, "RGB").resize((ow, oh), Image.BILINEAR))
        out[f"rs_in{n}"] = np.array([h, w, 50 + n, zlib.crc32(img.tobytes())])
        out[f"rs_outcrc{n}"] = np.array([oh, ow, zlib.crc32(np.ascontiguousarray(res).tobytes())])
        if n in (0, 5):
            out[f"rs_out{n}"] = res
    path = os.path.join(ROOT, "tests", "golden", "augment_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
