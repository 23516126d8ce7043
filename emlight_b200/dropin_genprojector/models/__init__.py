"""`models.create_model(opt)` (GenProjector/models/__init__.py:10-48): model class looked up by name in `models.<name>_model`."""
import importlib

import torch


def find_model_using_name(model_name):
    lib = importlib.import_module("models." + model_name + "_model")
    want = (model_name.replace("_", "") + "model").lower()
    for name, cls in vars(lib).items():
        if name.lower() == want and isinstance(cls, type) and issubclass(cls, torch.nn.Module):
            return cls
    raise ImportError("models/%s_model.py must define a torch.nn.Module subclass named like %r (case-insensitive)" % (model_name, want))


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    instance = find_model_using_name(opt.model)(opt)
    print("model [%s] was created" % type(instance).__name__)
    return instance
