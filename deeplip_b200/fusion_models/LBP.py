"""Drop-in for models/fusion_models/LBP.py: LowFER (:8-54).

The reference computes an MFB bilinear term from U, V and then discards it; the value it returns
is cat([e1, sigmoid(e2), e1*sigmoid(e2)], 1) (:46-50).  U, V, bn0, bn1 are kept as (dead)
parameters so checkpoints load; only the live path is computed (dl_lowfer)."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops


class LowFER(nn.Module):
    def __init__(self, d1, d2, o):
        super().__init__()
        k = 30
        self.U = nn.Parameter(torch.tensor(np.random.uniform(-1, 1, (d1, k * o)), dtype=torch.float))
        self.V = nn.Parameter(torch.tensor(np.random.uniform(-1, 1, (d2, k * o)), dtype=torch.float))
        self.bn0 = nn.BatchNorm1d(d1)
        self.bn1 = nn.BatchNorm1d(d1)
        self.k, self.o = k, o

    def forward(self, e1, e2):
        return ops.lowfer(e1, e2)
