"""CPU: host-side U-Net logic (parameter inventory) against the oracle's state dict."""
from oracle import unet as ounet


def _check(cfg):
    from rdm_b200.unet import unet_param_shapes
    shapes = unet_param_shapes(**cfg)
    sd = ounet.UNetModel(**cfg).state_dict()
    assert list(shapes) == list(sd), "names/order differ from the reference key layout"
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k


def test_param_inventory_matches_oracle_tiny():
    _check(ounet.TINY_UNET)


def test_param_inventory_matches_oracle_imagenet():
    from rdm_b200.unet import unet_param_shapes
    _check(ounet.IMAGENET_UNET)
    n = 0
    for shp in unet_param_shapes(**ounet.IMAGENET_UNET).values():
        m = 1
        for s in shp:
            m *= s
        n += m
    assert n == 400_920_579          # scripts/demo_rdm.ipynb:128
