"""`from geomloss import SamplesLoss` (RegressionNetwork/train.py:14)."""
from emlight_b200.samples_loss import SamplesLoss  # noqa: F401

__all__ = ["SamplesLoss"]
