"""Mirror of the reference's model_twoSteps.py: ImageCompressor_twoSteps is ImageCompressor with the transforms
stored as ``Encoder1`` / ``Decoder1`` (model_twoSteps.py:38-81; state_dict keys ``Encoder1.*``, ``Decoder1.*``), and a
``load_model`` that parses the iteration number out of the checkpoint name (model_twoSteps.py:22-35).  Same kernels,
same forward."""
import torch

from .model import *  # noqa: F401,F403  (star-import surface of model_twoSteps.py:1-15)
from .model import ImageCompressor, save_model  # noqa: F401


def load_model(model, f):
    """model_twoSteps.py:22-35: partial load filtered by key; returns the step parsed from 'iter_<n>.pth...'"""
    with open(f, 'rb') as fh:
        pretrained_dict = torch.load(fh, map_location="cpu")
        model_dict = model.state_dict()
        pretrained_dict = {k: v for k, v in pretrained_dict.items() if k in model_dict}
        model_dict.update(pretrained_dict)
        model.load_state_dict(model_dict)
    f = str(f)
    if f.find('iter_') != -1 and f.find('.pth') != -1:
        st = f.find('iter_') + 5
        ed = f.find('.pth', st)
        return int(f[st:ed])
    return 0


class ImageCompressor_twoSteps(ImageCompressor):
    _ENC, _DEC = "Encoder1", "Decoder1"
