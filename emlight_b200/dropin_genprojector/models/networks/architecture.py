"""`models.networks.architecture.{SPADEResnetBlock, VGG19}` (GenProjector/models/networks/architecture.py:22-120)."""
from emlight_b200.genprojector import SPADEResnetBlock, VGG19  # noqa: F401
