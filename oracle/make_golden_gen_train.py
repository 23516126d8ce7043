"""Golden vectors for the TRAIN-mode forward of the GenProjector generator (test infrastructure): imports the reference's
generator.py / architecture.py / normalization.py / sphere_cnn.py unchanged from /root/reference (IO modules stubbed),
runs SPADEGenerator(...).train() on CPU -- batch-statistic (Sync)BatchNorm falls back to F.batch_norm on one device, spectral_norm does
its power iteration -- and writes tests/golden/generator_train.npz.  Run from the repo root: python oracle/make_golden_gen_train.py"""
import argparse
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    for name in ("OpenEXR", "Imath", "imageio", "imageio.plugins", "imageio.plugins.freeimage", "vtk", "vtk.util", "vtk.util.numpy_support"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["imageio.plugins.freeimage"].download = lambda: None
    sys.modules["imageio"].plugins = sys.modules["imageio.plugins"]
    sys.modules["imageio.plugins"].freeimage = sys.modules["imageio.plugins.freeimage"]
    sys.modules["vtk"].util = sys.modules["vtk.util"]
    sys.modules["vtk.util"].numpy_support = sys.modules["vtk.util.numpy_support"]
    sys.path.insert(0, "/root/reference/GenProjector")
    from models.networks.generator import SPADEGenerator
    from oracle import genprojector_oracle as GO
    ngf = 8
    opt = argparse.Namespace(ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                             num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)
    G = SPADEGenerator(opt).train()
    sd = GO.init_generator_state_dict(seed=2, ngf=ngf)
    G.load_state_dict(sd)
    gen = torch.Generator().manual_seed(21)
    guide = torch.rand(2, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(2, 3, 128, 128, generator=gen)
    with torch.no_grad():
        out1 = G(guide, crop)
        out2 = G(guide * 0.5, crop)                    # second step: uses the buffers updated by the first
    new = G.state_dict()
    keep = ["head_0.conv_0.weight_u", "head_0.conv_0.weight_v", "up_3.conv_s.weight_u", "netE.layer1.0.weight_v", "netE.layer5.0.weight_u",
            "head_0.norm_0.param_free_norm.running_mean", "head_0.norm_0.param_free_norm.running_var",
            "up_3.norm_1.param_free_norm.running_mean", "up_3.norm_1.param_free_norm.running_var",
            "up_1.norm_s.param_free_norm.running_var", "up_3.norm_1.param_free_norm.num_batches_tracked"]
    np.savez_compressed(os.path.join(OUT, "generator_train.npz"), ngf=np.int64(ngf), sd_seed=np.int64(2), in_seed=np.int64(21),
                        out1=out1.numpy()[:, :, ::2, ::2], out2=out2.numpy()[:, :, ::2, ::2],
                        **{"buf_" + k: new[k].numpy() for k in keep})
    print("wrote generator_train.npz", out1.mean().item(), out2.mean().item())


if __name__ == "__main__":
    main()
