"""`models.networks.loss.{GANLoss, VGGLoss}` (GenProjector/models/networks/loss.py:16-114)."""
from emlight_b200.genprojector import GANLoss, VGGLoss  # noqa: F401
