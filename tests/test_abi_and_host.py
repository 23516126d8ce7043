"""CPU: the C-ABI library loads and exports every declared symbol; host-side drop-in contract."""
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        txt = open(os.path.join(inc, f)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(eml_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol(lib):
    from emlight_b200 import _lib
    decl = declared_symbols()
    assert decl, "no declarations found"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (eml_[a-z0-9_]+)", out))
    assert decl <= exported, "declared but not exported: %s" % sorted(decl - exported)
    assert decl == set(_lib.SIGNATURES), "ctypes table out of sync with the header"
    assert lib.eml_version() == _lib.ABI_VERSION
    assert lib.eml_error_string(-2) and b"shape" in lib.eml_error_string(-2)


def test_library_is_sm100a_tcgen05(lib):
    """The shipped cubin targets sm_100a and the conv kernel really is tcgen05 + bulk-TMA (SASS mnemonics)."""
    from emlight_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass.stdout
    assert "UTCHMMA" in sass.stdout          # tcgen05.mma kind::f16
    assert "LDTM" in sass.stdout             # tcgen05.ld
    assert "UBLKCP" in sass.stdout           # cp.async.bulk


def test_argument_errors_need_no_gpu(lib):
    from emlight_b200._lib import ConvParams
    assert lib.eml_sg_render_fwd(None, 0, None, 0, None, None, None, 1, 8, None) == -1
    assert lib.eml_sinkhorn_fwdbwd(None, None, None, None, None, 1, 96, 0.025, 0.5, 0.0, None, 0, None) == -1
    assert lib.eml_conv_forward(None, None) == -1
    p = ConvParams()
    assert lib.eml_conv_forward(p, None) == -1
    assert lib.eml_conv_wpack_bytes(12, 48, 9) == 9 * 2 * 16 * 128 + 2 * 9 * 6 * 16 * 16 + 9 * 6 * 32 * 16    # generic chunks + planar hi|lo + [hi;lo] concatenated image
    assert lib.eml_conv_wpack_bytes(48, 150, 1) == 3 * 2 * 48 * 128


def test_new_entry_points_validate_arguments_without_a_gpu(lib):
    """Host-side logic of the round-1 additions: shape predicates and NULL / range checks return before any launch."""
    from emlight_b200._lib import DenseLayerParams
    G = 12
    # the one-kernel dense layer: blocks 1-2 (W 256 / 128, C_in % 4 == 0), block 3 as image pairs (W 64, C_in % 2 == 0), <= 5 weight chunks
    assert [lib.eml_dense_layer_supported(192, 256, 24 + 12 * l, G, 1) for l in range(16)] == [1] * 16
    assert [lib.eml_dense_layer_supported(96, 128, 108 + 12 * l, G, 1) for l in range(16)] == [1] * 16
    assert [lib.eml_dense_layer_supported(48, 64, 150 + 12 * l, G, 1) for l in range(16)] == [1] * 15 + [0]
    assert lib.eml_dense_layer_supported(192, 256, 26, G, 1) == 0 and lib.eml_dense_layer_supported(48, 64, 151, G, 1) == 0
    assert lib.eml_dense_layer_supported(192, 256, 24, G, 2) == 0            # fp32 SIMT mode: two-kernel path
    assert lib.eml_dense_layer_forward(None, None) == -1
    assert lib.eml_dense_layer_forward(DenseLayerParams(), None) == -1
    assert lib.eml_needlet_basis(None, 1, None, None, 1, None, 1, 1, None, 2, None) == -1
    assert lib.eml_split_bf16(None, 1, 1, 1, None, None, 64, None) == -1
    assert lib.eml_needlet_sparsify(None, 1, 1, 1, None, 1, 0.1, None) == -1
    assert lib.eml_gemm_bf16_splitk(None, None, 1, 64, None, 1, None, None, 4, 0, 1, 1, None) == -1
    assert lib.eml_channel_stats(None, 4, 1, 4, None, None) == -1
    # round-2 entry points: NULL pointers are refused before anything is launched
    assert lib.eml_gemm_bf16_slices(None, None, 1, 64, None, 0, 2, 16, None, None, 36, 0, 1, 1, None) == -1
    assert lib.eml_gemm_pack_slices(None, None, 2, 16, 64, 4096, None) == -1
    assert lib.eml_spectral_norm(None, 4, 4, None, None, 1, 1e-12, None, None, None) == -1
    assert lib.eml_bias_act(None, 4, None, 1, None, 4, 1, 4, None) == -1
    assert lib.eml_wgrad_3x3(None, 16, 12, None, 48, 48, None, None, None, 1, 4, 64, 1, None) == -1
    assert lib.eml_dense_layer_compose(None, None, None, None, 48, 24, 12, None, None, None) == -1
    assert lib.eml_extract_params(None, None, None, 1, 128, 256, 64, None, None, None, None, None, None) == -1
    assert lib.eml_tonemap_hdr(None, None, None, 1, 1, 2.4, 50.0, 0.5, 1, 1, 0, None) == -1


def test_state_dict_contract_matches_reference_names():
    import emlight_b200 as E
    from oracle.densenet_oracle import init_state_dict
    sd = init_state_dict(0, 96)
    net = E.DenseNet()
    mine = net.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    assert all(tuple(mine[k].shape) == tuple(sd[k].shape) for k in sd)
    assert sum(p.numel() for p in net.parameters()) == 9336711
    net.load_state_dict(sd)
    assert [n for n, _ in net.features.named_children()][:6] == ["conv0", "norm0", "relu0", "denseblock1", "transition1", "last_norm1"]
    assert [n for n, _ in net.features.denseblock1.denselayer1.named_children()] == ["norm1", "relu1", "conv1", "norm2", "conv2"]


def test_no_cpu_fallback():
    import emlight_b200 as E
    with pytest.raises(RuntimeError, match="CUDA"):
        E.DenseNet()(torch.zeros(1, 3, 192, 256))
    with pytest.raises(RuntimeError, match="CUDA"):
        E.convert_to_panorama(torch.zeros(1, 24), torch.ones(1, 8), torch.zeros(1, 24))
    with pytest.raises(RuntimeError, match="CUDA"):
        E.SamplesLoss("sinkhorn", p=2, blur=.025)(torch.zeros(1, 96, 1), torch.zeros(1, 96, 1))
    with pytest.raises(ValueError):
        E.SamplesLoss("sinkhorn", p=2, blur=.025)(torch.zeros(1, 96, 1))
    with pytest.raises(ValueError):
        E.SamplesLoss("gaussian")


def test_dropin_module_names_resolve():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "import DenseNet, util; from geomloss import SamplesLoss; from gmloss import SamplesLoss as G;"
            "import emlight_b200 as E; assert DenseNet.DenseNet is E.DenseNet and SamplesLoss is E.SamplesLoss;"
            "assert util.convert_to_panorama is E.convert_to_panorama and util.sphere_points(96).shape == (96, 3);"
            "assert G is E.GMSamplesLoss") % (ROOT, os.path.join(ROOT, "emlight_b200", "dropin"))
    subprocess.check_call([sys.executable, "-c", code])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "emlight_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(d, f)).read(), f


def test_support_queries_answer_without_a_gpu():
    """The host-side shape rules that decide the execution plan (no device work): every block-1 / block-2 layer of the 192x256 network
    takes the one-kernel path, C_in = 330 (block 3, layer 16) does not; transition 1 reads channel planes, transitions 2 / 3 (C_out =
    160 / 171: not on the TMA pipeline) do not; fp32 never does."""
    from emlight_b200 import _lib
    lib = _lib.load()
    P = _lib.PRECISIONS
    assert all(lib.eml_dense_layer_supported(192, 256, 24 + 12 * l, 12, P["bf16x3"]) == 1 for l in range(16))
    assert all(lib.eml_dense_layer_supported(96, 128, 108 + 12 * l, 12, P["bf16"]) == 1 for l in range(16))
    assert lib.eml_dense_layer_supported(48, 64, 330, 12, P["bf16x3"]) == 0
    assert lib.eml_dense_layer_supported(192, 256, 24, 12, P["fp32"]) == 0
    assert lib.eml_dense_layer_supported(192, 200, 24, 12, P["bf16x3"]) == 0           # W must be 64, 128 or 256
    assert lib.eml_transition_planes_supported(192, 256, 216, 108, P["bf16x3"]) == 1
    assert lib.eml_transition_planes_supported(96, 128, 320, 160, P["bf16x3"]) == 0
    assert lib.eml_transition_planes_supported(48, 64, 342, 171, P["bf16x3"]) == 0
    assert lib.eml_transition_planes_supported(192, 256, 216, 108, P["fp32"]) == 0
    assert lib.eml_dense_layer_wpack_bytes(24) > 0 and lib.eml_conv_wpack_bytes(256, 1152, 1) == 18 * 2 * 256 * 128
