# API mirror: the class / field / function names and argument lists in this file follow ocr4all_pixel_classifier
# (https://github.com/ocr-d-modul-2-segmentierung/page-segmentation, (c) its authors, licensed Apache-2.0 OR
# GPL-3.0-or-later) so that it drops in for the reference; the arithmetic underneath is this repository's own
# (pcs_* calls into libpcseg_b200.so).
"""Mirror of ocr4all_pixel_classifier/lib/architecture.py:5-68 for the
architectures in scope (fcn_skip, fcn, unet); the ImageNet-pretrained encoders
are out of scope."""
import enum


class Architecture(enum.Enum):
    FCN_SKIP = 'fcn_skip'
    FCN = 'fcn'
    RES_NET = 'image_res_net'
    RES_UNET = 'res_unet'
    MOBILE_NET = 'mobile_net'
    UNET = 'unet'
    EFFNETB0 = 'effb0'
    EFFNETB1 = 'effb1'
    EFFNETB2 = 'effb2'
    EFFNETB3 = 'effb3'
    EFFNETB4 = 'effb4'
    EFFNETB5 = 'effb5'
    EFFNETB6 = 'effb6'
    EFFNETB7 = 'effb7'

    def supported(self) -> bool:
        return self in (Architecture.FCN_SKIP, Architecture.FCN, Architecture.UNET)

    def preprocess(self):
        """architecture.py:45-64 -> (preprocess_fn, rgb)."""
        if not self.supported():
            raise NotImplementedError(f"architecture {self.value} is outside the B200 hot-path scope")
        return default_preprocess, False


def default_preprocess(x):
    """architecture.py:67-68."""
    return x / 255.0
