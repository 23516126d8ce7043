"""`models.pix2pix_model.Pix2PixModel` (GenProjector/models/pix2pix_model.py:12-186) -> emlight_b200.genprojector.Pix2PixModel, built the
way the reference builds it: networks through `networks.define_G / define_D` (print_network, init_weights(opt.init_type,
opt.init_variance): :80-88), checkpoints loaded when testing or continuing (:84-87), modules left in training mode unless the script
calls `.eval()` (the reference never calls `.train()`), `save(epoch)` (:72-74).  With `opt.isTrain` the loss dictionaries carry the
autograd node of emlight_b200/gp_train.py, which is what lets `model_trainer.Trainer` run unchanged."""
import torch

import models.networks as networks  # noqa: F401  (the reference module exposes it too)
import util
from emlight_b200.genprojector import Pix2PixModel as _Pix2PixModel


class Pix2PixModel(_Pix2PixModel):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        networks.modify_commandline_options(parser, is_train)
        return parser

    def __init__(self, opt):
        super().__init__(opt)                      # builds the networks through initialize_networks() below
        self.FloatTensor = torch.cuda.FloatTensor if self.use_gpu() else torch.FloatTensor
        self.train(True)
        self.autograd = bool(opt.isTrain)

    def initialize_networks(self, opt):
        training = bool(opt.isTrain)
        nets = {"G": networks.define_G(opt), "D": networks.define_D(opt) if training else None}
        if not training or getattr(opt, "continue_train", False):                 # test time, or resuming: read the checkpoints
            for label in ("G", "D"):
                if nets[label] is not None:
                    nets[label] = util.load_network(nets[label], label, opt.which_epoch, opt)
        return nets["G"], nets["D"]

    def save(self, epoch):
        for label, net in (("G", self.netG), ("D", self.netD)):
            util.save_network(net, label, epoch, self.opt)
