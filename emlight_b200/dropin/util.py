"""`import util; util.sphere_points(n); util.convert_to_panorama(dirs, sizes, colors)` (RegressionNetwork/train.py:67,111,122),
`util.TonemapHDR(gamma, percentile, max_mapping)` (RegressionNetwork/data.py:62-73 / util.py:36-66; CUDA tensors in, CUDA tensors out),
`util.load_exr(path)` / `util.write_exr(path, data)` (util.py:301-306, GenProjector/util.py:248-277; host-side file IO without the
OpenEXR bindings)."""
from emlight_b200.panorama import convert_to_panorama, sphere_points  # noqa: F401
from emlight_b200.tonemap import TonemapHDR  # noqa: F401
from emlight_b200.wire import load_exr, write_exr  # noqa: F401
from emlight_b200.handlers import (PanoramaHandler, cartesian_to_polar, polar_to_cartesian, print_model_parm_nums,  # noqa: F401
                                   tonemapping)
