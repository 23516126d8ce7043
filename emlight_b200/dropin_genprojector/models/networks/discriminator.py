"""`models.networks.discriminator.{MultiscaleDiscriminator, NLayerDiscriminator}` (GenProjector/models/networks/discriminator.py:16-125)."""
from emlight_b200.genprojector import MultiscaleDiscriminator as _Multi, NLayerDiscriminator as _NLayer
from models.networks.base_network import BaseNetwork


class NLayerDiscriminator(_NLayer, BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument("--n_layers_D", type=int, default=4, help="# layers in each discriminator")      # discriminator.py:72
        return parser


class MultiscaleDiscriminator(_Multi, BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument("--netD_subarch", type=str, default="n_layer", help="architecture of each discriminator")
        parser.add_argument("--num_D", type=int, default=2, help="number of discriminators to be used in multiscale")
        NLayerDiscriminator.modify_commandline_options(parser, is_train)              # the only sub-architecture (discriminator.py:26-28)
        return parser
