"""CPU: the GenProjector module-name shims (emlight_b200/dropin_genprojector) satisfy the reference's by-name factory contract
(GenProjector/models/networks/__init__.py:15-61, models/__init__.py:10-48).  Run in a subprocess: the shim package is
called `models`, like the reference's."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

SCRIPT = textwrap.dedent('''
    import argparse, sys
    sys.path.insert(0, %r); sys.path.insert(0, %r)
    import torch
    import models, models.networks as N
    from models.networks.base_network import BaseNetwork
    from oracle import genprojector_oracle as GO

    parser = argparse.ArgumentParser()
    parser.add_argument("--netG", default="spade"); parser.add_argument("--netD", default="multiscale"); parser.add_argument("--norm_G", default="x")
    parser = N.modify_commandline_options(parser, True)
    o = parser.parse_args([])
    assert o.norm_G == "spectralspadesyncbatch3x3" and o.num_upsampling_layers == "normal" and o.num_D == 2 and o.n_layers_D == 4

    opt = argparse.Namespace(ngf=8, ndf=8, norm_G=o.norm_G, norm_E="spectralinstance", norm_D="spectralinstance", semantic_nc=3, label_nc=3,
                             output_nc=3, num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0, netG="spade", netD="multiscale",
                             netD_subarch="n_layer", num_D=2, n_layers_D=4, gpu_ids=[], init_type="xavier", init_variance=0.02,
                             no_ganFeat_loss=False, contain_dontcare_label=False, no_instance=True)
    cls = N.find_network_using_name(opt.netG, "generator")
    assert cls.__name__ == "SPADEGenerator" and issubclass(cls, BaseNetwork)
    assert issubclass(N.find_network_using_name(opt.netD, "discriminator"), BaseNetwork)
    assert issubclass(N.find_network_using_name("conv", "encoder"), BaseNetwork)
    torch.manual_seed(0)
    G = cls(opt)
    before = {k: v.clone() for k, v in G.state_dict().items()}
    G2 = N.define_G(opt)                                          # print_network + init_weights('xavier', 0.02)
    sd = G2.state_dict()
    ref = GO.init_generator_state_dict(0, 8)
    assert set(sd) == set(ref) and all(sd[k].shape == ref[k].shape for k in ref)        # the 253-entry state_dict contract
    w = sd["up_3.norm_0.mlp_gamma.weight"]
    assert abs(float(w.std()) - 0.02 * (2.0 / (w.shape[0] * 9 + w.shape[1] * 9)) ** 0.5) < 2e-4   # xavier_normal_(gain=0.02)
    assert float(sd["up_3.norm_0.mlp_gamma.bias"].abs().max()) == 0.0 and float(sd["up_3.conv_0.bias"].abs().max()) == 0.0
    assert float(sd["netE.fc.weight"].std()) < 1e-3                                           # Linear layers are initialised too
    D = N.define_D(opt)
    assert type(D).__name__ == "MultiscaleDiscriminator" and len(D.state_dict()) > 0
    assert models.find_model_using_name("pix2pix").__name__ == "Pix2PixModel"
    assert callable(models.get_option_setter("pix2pix"))
    try:
        N.find_network_using_name("nosuch", "generator")
        raise SystemExit("lookup of a missing class must fail")
    except ImportError:
        pass
    print("ok")
''') % (ROOT, os.path.join(ROOT, "emlight_b200", "dropin_genprojector"))


def test_genprojector_factories_resolve_to_the_sm100a_classes():
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-1500:] + r.stderr[-3000:]
