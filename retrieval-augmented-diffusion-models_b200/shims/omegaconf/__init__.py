"""Tiny `omegaconf` stand-in: YAML -> mutable attribute-access mappings (what scripts/rdm_sample.py:155-160 needs)."""
import yaml

from .listconfig import ListConfig


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __delattr__(self, k):
        del self[k]

    def get(self, k, d=None):
        return self[k] if k in self else d

    def pop(self, k, *d):
        return dict.pop(self, k, *d)


def _wrap(v):
    if isinstance(v, (DictConfig, ListConfig)):
        return v
    if isinstance(v, dict):
        return DictConfig({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return ListConfig([_wrap(x) for x in v])
    return v


class OmegaConf:
    @staticmethod
    def load(path):
        with open(str(path)) as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def create(obj=None):
        return _wrap(obj if obj is not None else {})

    @staticmethod
    def to_container(cfg, resolve=True):
        if isinstance(cfg, dict):
            return {k: OmegaConf.to_container(v) for k, v in cfg.items()}
        if isinstance(cfg, list):
            return [OmegaConf.to_container(v) for v in cfg]
        return cfg

    @staticmethod
    def merge(*cfgs):
        out = DictConfig()

        def rec(dst, src):
            for k, v in src.items():
                if isinstance(v, dict) and isinstance(dst.get(k), dict):
                    rec(dst[k], v)
                else:
                    dst[k] = _wrap(v)
        for c in cfgs:
            rec(out, c)
        return out
