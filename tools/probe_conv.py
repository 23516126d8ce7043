"""First-contact diagnostics for the tcgen05 conv kernel: prints structured error maps instead of a bare assert.
Usage (GPU box): python tools/probe_conv.py > gpurun_out/probe.log 2>&1"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from emlight_b200 import build, _lib
    build.build()
    lib = _lib.load()
    print("device_ok", lib.eml_device_ok(), torch.cuda.get_device_name(0))
    cuda = torch.device("cuda:0")
    from test_conv_gpu import conv_ref, run_conv
    torch.manual_seed(0)

    def case(tag, mode, B, H, W, c_in, pitch, c_out, relu, precision, onehot=False):
        x = torch.randn(B, H, W, pitch)
        scale = torch.ones(c_in); shift = torch.zeros(c_in)
        taps = 3 if mode == "3x3" else 1
        if onehot:
            w = torch.zeros(c_out, c_in, taps, taps)
            for n in range(c_out):
                w[n, (7 * n + 3) % c_in, taps // 2, taps // 2] = 1.0
        else:
            w = torch.randn(c_out, c_in, taps, taps) / np.sqrt(c_in * taps * taps)
        try:
            out, _ = run_conv(lib, cuda, x, c_in, scale, shift, w, mode, relu, precision, c_out, 0, False)
        except Exception as e:      # noqa: BLE001
            print(tag, "EXCEPTION", repr(e)); return False
        ref = conv_ref(x, c_in, scale, shift, w, mode, relu)
        got = out[..., :c_out].double()
        err = (got - ref).abs()
        rel = err.max().item() / ref.abs().max().item()
        print("%-40s rel_err %.3e  nan %d" % (tag, rel, int(torch.isnan(got).sum())))
        if rel > 5e-2 or torch.isnan(got).any():
            e2 = err.reshape(-1, c_out)
            rows = e2.max(1).values
            print("   bad rows (of %d):" % rows.numel(), (rows > 1e-2).nonzero().flatten()[:32].tolist())
            print("   bad cols:", (e2.max(0).values > 1e-2).nonzero().flatten()[:32].tolist())
            print("   got[0,:8]", got.reshape(-1, c_out)[0, :8].tolist())
            print("   ref[0,:8]", ref.reshape(-1, c_out)[0, :8].tolist())
            if onehot:
                # which input channel did each output column actually pick up (row 0..3)?
                xr = x.reshape(-1, pitch)[:, :c_in].double()
                for r in (0, 1, 9):
                    picks = []
                    for n in range(min(c_out, 8)):
                        d = (xr[r] - got.reshape(-1, c_out)[r, n]).abs()
                        picks.append(int(d.argmin()) if d.min() < 1e-2 else -1)
                    print("   row %d picks channels" % r, picks, "expected", [(7 * n + 3) % c_in for n in range(min(c_out, 8))])
            return False
        return True

    ok = True
    ok &= case("fp32   1x1 c64->16 one tile", "1x1", 1, 8, 16, 64, 64, 16, False, "fp32")
    ok &= case("bf16   1x1 c64->16 onehot", "1x1", 1, 8, 16, 64, 64, 16, False, "bf16", onehot=True)
    ok &= case("bf16   1x1 c64->16 one tile", "1x1", 1, 8, 16, 64, 64, 16, False, "bf16")
    ok &= case("bf16x3 1x1 c64->16 one tile", "1x1", 1, 8, 16, 64, 64, 16, False, "bf16x3")
    ok &= case("bf16x3 1x1 c24->48", "1x1", 1, 8, 16, 24, 216, 48, True, "bf16x3")
    ok &= case("bf16x3 1x1 c128->48 (2 chunks)", "1x1", 1, 8, 16, 128, 128, 48, True, "bf16x3")
    ok &= case("bf16x3 1x1 c330->48 (6 chunks)", "1x1", 2, 8, 16, 330, 344, 48, True, "bf16x3")
    ok &= case("bf16x3 3x3 c48->12", "3x3", 1, 8, 16, 48, 48, 12, False, "bf16x3")
    ok &= case("bf16x3 pool c216->108", "pool", 1, 16, 16, 216, 216, 108, True, "bf16x3")
    ok &= case("bf16x3 pool c342->171", "pool", 1, 16, 16, 342, 344, 171, True, "bf16x3")
    print("PROBE", "OK" if ok else "FAILED")


if __name__ == "__main__":
    main()
