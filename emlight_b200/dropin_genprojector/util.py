"""`import util` for the GenProjector scripts (GenProjector/data.py:12, test.py, trainers): the names they use from
GenProjector/util.py that touch the hot path or its wire formats -- `TonemapHDR`, `load_exr`, `write_exr`, `sphere_points`,
`convert_to_panorama`, `tonemapping(im, sv_path)`, `save_test_images`, `save_current_images`, `print_current_errors`,
`save_network` / `load_network`, `PanoramaHandler` -- without the OpenEXR / Imath / vtk imports of the reference file."""
from emlight_b200.handlers import (PanoramaHandler, cartesian_to_polar, convert_visuals_to_numpy, load_network,  # noqa: F401
                                   polar_to_cartesian, print_current_errors, save_current_images, save_network, save_test_images)
from emlight_b200.handlers import copyconf, mkdir, mkdirs, natural_sort, str2bool  # noqa: F401
from emlight_b200.handlers import tonemapping_to_file as tonemapping  # noqa: F401  (GenProjector's variant writes the image file)
from emlight_b200.panorama import convert_to_panorama, sphere_points  # noqa: F401
from emlight_b200.tonemap import TonemapHDR  # noqa: F401
from emlight_b200.wire import load_exr, write_exr  # noqa: F401


def find_class_in_module(target_cls_name, module):
    from models.networks import find_class_in_module as _find                               # util.py:158-170
    return _find(target_cls_name, module)
