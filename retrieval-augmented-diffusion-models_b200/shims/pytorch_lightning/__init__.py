"""`pytorch_lightning.seed_everything` stand-in (scripts/rdm_sample.py:235-236 seeds python, numpy and torch)."""
import os
import random

import numpy as np
import torch


def seed_everything(seed=None, workers=False):
    seed = int(seed if seed is not None else 0)
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed
