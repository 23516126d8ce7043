"""`BaseNetwork` (GenProjector/models/networks/base_network.py:10-59): the base class `find_network_using_name` insists on."""
import torch.nn as nn
from torch.nn import init

_SCHEMES = {
    "normal": lambda w, gain: init.normal_(w, 0.0, gain),
    "xavier": lambda w, gain: init.xavier_normal_(w, gain=gain),
    "xavier_uniform": lambda w, gain: init.xavier_uniform_(w, gain=1.0),
    "kaiming": lambda w, gain: init.kaiming_normal_(w, a=0, mode="fan_in"),
    "orthogonal": lambda w, gain: init.orthogonal_(w, gain=gain),
}


class BaseNetwork(nn.Module):
    def __init__(self):
        super().__init__()

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def print_network(self):
        n = sum(p.numel() for p in self.parameters())
        print("Network [%s] was created. Total number of parameters: %.1f million. To see the architecture, do print(network)."
              % (type(self).__name__, n / 1e6))

    def init_weights(self, init_type="normal", gain=0.02):
        """Same effect as base_network.py:28-59: BatchNorm affine ~ N(1, gain) / 0, conv and linear weights by the named scheme with zero
        bias.  A spectral-norm wrapped layer only gets its bias zeroed: the reference initialises the DERIVED `weight` attribute of such
        a module, which leaves `weight_orig` (what is trained and saved) untouched."""
        if init_type not in _SCHEMES and init_type != "none":
            raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
        for m in self.modules():
            cname = type(m).__name__
            if "BatchNorm2d" in cname:
                if getattr(m, "weight", None) is not None:
                    init.normal_(m.weight.data, 1.0, gain)
                if getattr(m, "bias", None) is not None:
                    init.constant_(m.bias.data, 0.0)
            elif "Conv" in cname or "Linear" in cname:
                spectral = hasattr(m, "weight_orig")
                if not spectral and getattr(m, "weight", None) is None:
                    continue
                if not spectral:
                    if init_type == "none":
                        m.reset_parameters()
                    else:
                        _SCHEMES[init_type](m.weight.data, gain)
                if getattr(m, "bias", None) is not None:
                    init.constant_(m.bias.data, 0.0)
