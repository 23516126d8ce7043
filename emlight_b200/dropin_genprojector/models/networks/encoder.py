"""`models.networks.encoder.ConvEncoder`: the reference looks up 'conv' + 'encoder' here (networks/__init__.py:34,58); the class itself
lives in generator.py:90-126 of the reference."""
from emlight_b200.genprojector import ConvEncoder as _ConvEncoder
from models.networks.base_network import BaseNetwork


class ConvEncoder(_ConvEncoder, BaseNetwork):
    pass
