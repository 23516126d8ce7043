"""`import util` for the GenProjector scripts (GenProjector/data.py:12, test.py, trainers): the names they use from
GenProjector/util.py that touch the hot path or its wire formats -- `TonemapHDR`, `load_exr`, `write_exr`, `sphere_points`,
`convert_to_panorama`, `tonemapping`, `PanoramaHandler` -- without the OpenEXR / Imath / vtk imports of the reference file."""
from emlight_b200.handlers import PanoramaHandler, cartesian_to_polar, polar_to_cartesian, tonemapping  # noqa: F401
from emlight_b200.panorama import convert_to_panorama, sphere_points  # noqa: F401
from emlight_b200.tonemap import TonemapHDR  # noqa: F401
from emlight_b200.wire import load_exr, write_exr  # noqa: F401
