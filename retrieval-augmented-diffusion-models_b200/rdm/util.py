"""`rdm.util` helpers the sampling path touches (`rdm/util.py:17`, used at ddpm.py:723)."""
import numpy as np
import torch


def isimage(x):
    return isinstance(x, (torch.Tensor, np.ndarray)) and x.ndim == 4 and x.shape[1] in (1, 3)


def ischannellastimage(x):
    return isinstance(x, (torch.Tensor, np.ndarray)) and x.ndim == 4 and x.shape[-1] in (1, 3)


class SampleLogs(dict):
    """The dictionary a sampling entry point returns.  Iterating it (the reference's scripts save one image file per key and element:
    `scripts/rdm_sample.py:253-261,301-309`) yields exactly the reference's keys; the extra device-side results of this implementation
    (`nns`: neighbour indices, `latents`: un-decoded samples, `sampled_indices`: RARM token ids) live in `.extras` and are still reachable
    by subscription (`logs["nns"]`)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.extras = {}

    def __missing__(self, key):
        return self.extras[key]
