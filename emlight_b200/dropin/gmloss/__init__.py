"""`from gmloss import SamplesLoss` (RegressionNetwork/test.py:14): forward(x, y, geometry)."""
from emlight_b200.samples_loss import GMSamplesLoss as SamplesLoss  # noqa: F401

__all__ = ["SamplesLoss"]
