"""`models.pix2pix_model.Pix2PixModel` (GenProjector/models/pix2pix_model.py:12-186) -> emlight_b200.genprojector.Pix2PixModel."""
import models.networks as networks  # noqa: F401  (the reference module exposes it too)
from emlight_b200.genprojector import Pix2PixModel as _Pix2PixModel


class Pix2PixModel(_Pix2PixModel):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        networks.modify_commandline_options(parser, is_train)
        return parser
