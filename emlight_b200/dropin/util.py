"""`import util; util.sphere_points(n); util.convert_to_panorama(dirs, sizes, colors)` (RegressionNetwork/train.py:67,111,122)."""
from emlight_b200.panorama import convert_to_panorama, sphere_points  # noqa: F401
