"""emlight_b200 -- B200 (sm_100a) implementation of EMLight's illumination-estimation hot path.

Public surface mirrors the reference's:

    from emlight_b200 import DenseNet            # RegressionNetwork/DenseNet.py: DenseNet
    from emlight_b200 import SamplesLoss         # RegressionNetwork/geomloss: SamplesLoss  (gmloss: GMSamplesLoss)
    from emlight_b200 import sphere_points, convert_to_panorama     # RegressionNetwork/util.py
    from emlight_b200 import SphereConv2D, SPADE, SPADEResnetBlock, SPADEGenerator   # GenProjector/models/networks/*
    from emlight_b200 import MultiscaleDiscriminator, GANLoss, VGGLoss, Pix2PixModel # discriminator.py, loss.py, pix2pix_model.py
    from emlight_b200.needlets import SNvertex, NeedletTransform                     # Needlets/sphere_needlets.py, mat_gen2.py (needs scipy)

Module-name shims for unchanged reference scripts live in ``emlight_b200/dropin`` (put it on sys.path).
All arithmetic runs in hand-written CUDA reached through the C ABI of include/emlight_b200.h.
"""
from .panorama import convert_to_panorama, genprojector_guide, render_from_params, sphere_points  # noqa: F401
from .samples_loss import GMSamplesLoss, SamplesLoss  # noqa: F401
from .densenet import DenseNet  # noqa: F401
from .genprojector import (SPADE, ConvEncoder, GANLoss, MultiscaleDiscriminator, NLayerDiscriminator, Pix2PixModel,  # noqa: F401
                           SPADEGenerator, SPADEResnetBlock, SphereConv2D, VGG19, VGGLoss)

__all__ = ["DenseNet", "SamplesLoss", "GMSamplesLoss", "sphere_points", "convert_to_panorama", "render_from_params", "genprojector_guide",
           "SphereConv2D", "SPADE", "SPADEResnetBlock", "ConvEncoder", "SPADEGenerator", "MultiscaleDiscriminator", "NLayerDiscriminator",
           "GANLoss", "VGG19", "VGGLoss", "Pix2PixModel"]
