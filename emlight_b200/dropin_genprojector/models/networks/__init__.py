"""`models.networks.define_G / define_D / define_E` and the by-name class lookup (GenProjector/models/networks/__init__.py:15-61)."""
import importlib

import torch

from models.networks.base_network import BaseNetwork
from models.networks.loss import *           # noqa: F401,F403
from models.networks.discriminator import *  # noqa: F401,F403
from models.networks.generator import *      # noqa: F401,F403
from models.networks.encoder import *        # noqa: F401,F403


def find_class_in_module(target_cls_name, module):
    """GenProjector/util.py:158-170: case-insensitive lookup, underscores ignored."""
    want = target_cls_name.replace("_", "").lower()
    lib = importlib.import_module(module)
    for name, obj in vars(lib).items():
        if name.lower() == want:
            return obj
    raise ImportError("In %s, there should be a class whose name matches %s in lowercase without underscore(_)" % (module, want))


def find_network_using_name(target_network_name, filename):
    network = find_class_in_module(target_network_name + filename, "models.networks." + filename)
    assert issubclass(network, BaseNetwork), "Class %s should be a subclass of BaseNetwork" % network
    return network


def modify_commandline_options(parser, is_train):
    opt, _ = parser.parse_known_args()
    parser = find_network_using_name(opt.netG, "generator").modify_commandline_options(parser, is_train)
    if is_train:
        parser = find_network_using_name(opt.netD, "discriminator").modify_commandline_options(parser, is_train)
    parser = find_network_using_name("conv", "encoder").modify_commandline_options(parser, is_train)
    return parser


def create_network(cls, opt):
    net = cls(opt)
    net.print_network()
    if len(opt.gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.cuda()
    net.init_weights(opt.init_type, opt.init_variance)
    return net


def define_G(opt):
    return create_network(find_network_using_name(opt.netG, "generator"), opt)


def define_D(opt):
    return create_network(find_network_using_name(opt.netD, "discriminator"), opt)


def define_E(opt):
    return create_network(find_network_using_name("conv", "encoder"), opt)
