"""The reference's config.py keys that the forward path reads (config.py:83-94,132; CONST.IMG_*), with
the reference defaults.  Any attribute-style mapping works as `cfg` (easydict, argparse.Namespace of
namespaces, this AttrDict); the modules only ever do cfg.NETWORK.<KEY> / cfg.TEST.VOXEL_THRESH."""


class AttrDict(dict):
    def __getattr__(self, key):
        try:
            val = self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc
        if isinstance(val, dict) and not isinstance(val, AttrDict):
            val = AttrDict(val)
            self[key] = val
        return val

    def __setattr__(self, key, val):
        self[key] = val


def make_cfg(**network):
    net = AttrDict(
        LEAKY_VALUE=.2,
        TCONV_USE_BIAS=False,
        USE_REFINER=True,
        USE_MERGER=True,
        USE_SWIN_T_MULTI_STAGE=True,
        SWIN_T_STAGES=[0, 1, 2, 3],
        USE_CROSS_VIEW_ATTENTION=True,
        CROSS_ATT_REDUCTION_RATIO=4,
        ATT_SPATIAL_DOWNSAMPLE_RATIO=2,
        CROSS_ATT_NUM_HEADS=4,
    )
    net.update(network)
    return AttrDict(NETWORK=net, TEST=AttrDict(VOXEL_THRESH=[.2, .3, .4, .5]),
                    CONST=AttrDict(IMG_W=224, IMG_H=224, DEVICE='0'))


cfg = make_cfg()
