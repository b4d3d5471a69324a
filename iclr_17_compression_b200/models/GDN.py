"""Host-side mirror of the reference's models/GDN.py: same class name, constructor, parameter names /
shapes / init (GDN.py:29-62) so state_dicts interchange; forward runs libb2c's fused kernel."""
import torch
from torch import nn

from .. import functional as B2F


class GDN(nn.Module):
    """Generalized divisive normalization: y_i = x_i / sqrt(beta_i + sum_j gamma[i,j] x_j^2)
    (inverse=True multiplies instead).  beta/gamma are stored reparametrised exactly as in the
    reference (GDN.py:46-62) and bounded inside the kernel (GDN.py:73-79)."""

    def __init__(self, ch, inverse=False, beta_min=1e-6, gamma_init=0.1, reparam_offset=2 ** -18):
        super().__init__()
        if beta_min != 1e-6 or reparam_offset != 2 ** -18:
            raise ValueError("libb2c's GDN kernel is compiled for the reference defaults beta_min=1e-6, "
                             "reparam_offset=2**-18")
        self.inverse = inverse
        self.beta_min = beta_min
        self.gamma_init = gamma_init
        self.reparam_offset = reparam_offset
        self.build(ch)

    def build(self, ch):
        self.pedestal = self.reparam_offset ** 2
        self.beta_bound = (self.beta_min + self.reparam_offset ** 2) ** 0.5
        self.gamma_bound = self.reparam_offset
        self.beta = nn.Parameter(torch.sqrt(torch.ones(ch) + self.pedestal))
        self.gamma = nn.Parameter(torch.sqrt(self.gamma_init * torch.eye(ch) + self.pedestal))

    def forward(self, inputs):
        unfold = inputs.dim() == 5                       # GDN.py:66-69
        if unfold:
            bs, ch, d, w, h = inputs.size()
            inputs = inputs.view(bs, ch, d * w, h)
        outputs = B2F.gdn(inputs, self.beta, self.gamma, self.inverse)
        if unfold:
            outputs = outputs.view(bs, ch, d, w, h)
        return outputs
