"""`rdm.util` helpers the sampling path touches (`rdm/util.py:17`, used at ddpm.py:723)."""
import numpy as np
import torch


def isimage(x):
    return isinstance(x, (torch.Tensor, np.ndarray)) and x.ndim == 4 and x.shape[1] in (1, 3)


def ischannellastimage(x):
    return isinstance(x, (torch.Tensor, np.ndarray)) and x.ndim == 4 and x.shape[-1] in (1, 3)
