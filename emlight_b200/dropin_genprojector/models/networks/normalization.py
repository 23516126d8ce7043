"""`models.networks.normalization.SPADE` (GenProjector/models/networks/normalization.py:68-115)."""
from emlight_b200.genprojector import SPADE  # noqa: F401
