"""Puts the stand-ins of `shims/` on sys.path for every third-party package that cannot be imported here."""
import importlib.util
import os
import sys

SHIMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shims")


def install_shims():
    missing = [m for m in ("ldm", "omegaconf", "pytorch_lightning", "clip", "taming") if importlib.util.find_spec(m) is None]
    if missing and SHIMS not in sys.path:
        sys.path.append(SHIMS)          # appended: a real installation always wins
    return missing
