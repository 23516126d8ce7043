"""`models.networks.spherenet.SphereConv2D` (GenProjector/models/networks/spherenet/sphere_cnn.py:87-124)."""
from emlight_b200.genprojector import SphereConv2D  # noqa: F401
