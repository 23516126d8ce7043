"""`models.networks.generator.SPADEGenerator` (GenProjector/models/networks/generator.py:16-88)."""
from emlight_b200.genprojector import SPADEGenerator as _SPADEGenerator
from models.networks.base_network import BaseNetwork


class SPADEGenerator(_SPADEGenerator, BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.set_defaults(norm_G="spectralspadesyncbatch3x3")                       # generator.py:20
        parser.add_argument("--num_upsampling_layers", choices=("normal", "more", "most"), default="normal",
                            help="only 'normal' (the EMLight default) is implemented on sm_100a")
        return parser
