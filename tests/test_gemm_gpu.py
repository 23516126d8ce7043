"""GPU: the TMA-fed tcgen05 GEMM (`csrc/gemm_tma.cu`) on its own -- the convolution half of SphereConv2D (sphere_cnn.py:123): one
slice, several slices in one launch (bit-identical to per-slice launches), split-K, ragged M / N, both precision tiers."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _operands(cuda, M, K, N, seed):
    from emlight_b200 import _lib
    lib, P, st = _lib.load(), _lib.ptr, _lib.stream_ptr()
    gen = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=gen).to(cuda)
    W = (torch.randn(N, K, generator=gen) / K ** 0.5).to(cuda)
    bias = torch.randn(N, generator=gen).to(cuda)
    Kp = (K + 63) // 64 * 64
    hi = torch.zeros(M, Kp, dtype=torch.bfloat16, device=cuda)
    lo = torch.zeros(M, Kp, dtype=torch.bfloat16, device=cuda)
    _lib.check(lib.eml_split_bf16(P(A), M, K, K, P(hi), P(lo), Kp, st), "eml_split_bf16")
    Wp = torch.zeros(N, Kp, device=cuda)
    Wp[:, :K] = W
    return A, W, bias, hi, lo, Wp, Kp


def _pack(lib, Wp, n0, n, Kp, buf=None):
    from emlight_b200 import _lib
    w_s = Wp[n0:n0 + n].contiguous()
    if buf is None:
        buf = torch.empty(lib.eml_conv_wpack_bytes(n, Kp, 1), dtype=torch.uint8, device=Wp.device)
    _lib.check(lib.eml_conv_pack_weights(_lib.ptr(w_s), _lib.ptr(buf), n, Kp, 1, _lib.stream_ptr()), "pack")
    torch.cuda.synchronize()
    return buf


@pytest.mark.parametrize("precision,tol", [("bf16x3", 5e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("M,K,N", [(300, 576, 256), (128, 64, 16), (1000, 1152, 100), (37, 9216, 3)])
def test_gemm_single_slice(cuda, lib, M, K, N, precision, tol):
    from emlight_b200 import _lib
    P, st = _lib.ptr, _lib.stream_ptr()
    A, W, bias, hi, lo, Wp, Kp = _operands(cuda, M, K, N, M + N)
    pack = _pack(lib, Wp, 0, N, Kp)
    pitch = (N + 3) // 4 * 4 + 8
    out = torch.full((M, pitch), 7.0, device=cuda)
    _lib.check(lib.eml_gemm_bf16(P(hi), P(lo) if precision == "bf16x3" else None, M, Kp, P(pack), N, P(bias), P(out), pitch, 4,
                                 _lib.PRECISIONS[precision], st), "eml_gemm_bf16")
    want = A.double() @ W.double().t() + bias.double()
    assert float((out[:, 4:4 + N].double() - want).abs().max()) <= tol * float(want.abs().max())
    assert bool((out[:, :4] == 7.0).all()) and bool((out[:, 4 + N:] == 7.0).all())            # nothing outside the slice is touched


@pytest.mark.parametrize("M,K,nsl,N", [(2048, 1152, 4, 256), (100, 640, 2, 256), (513, 128, 3, 64)])
def test_gemm_slices_in_one_launch_equal_per_slice_launches(cuda, lib, M, K, nsl, N):
    """eml_gemm_bf16_slices == nsl calls of eml_gemm_bf16, bit for bit (same K order per output), and both match fp64."""
    from emlight_b200 import _lib
    P, st = _lib.ptr, _lib.stream_ptr()
    A, W, bias, hi, lo, Wp, Kp = _operands(cuda, M, K, nsl * N, M + K)
    sb = lib.eml_conv_wpack_bytes(N, Kp, 1)
    pack_all = torch.empty(sb * nsl, dtype=torch.uint8, device=cuda)
    for s in range(nsl):
        _pack(lib, Wp, s * N, N, Kp, pack_all[s * sb:(s + 1) * sb])
    pitch = nsl * N + 4
    one = torch.zeros(M, pitch, device=cuda)
    ref = torch.zeros(M, pitch, device=cuda)
    _lib.check(lib.eml_gemm_bf16_slices(P(hi), P(lo), M, Kp, P(pack_all), sb, nsl, N, P(bias), P(one), pitch, 4, _lib.PRECISIONS["bf16x3"], 1, st),
               "eml_gemm_bf16_slices")
    for s in range(nsl):
        _lib.check(lib.eml_gemm_bf16(P(hi), P(lo), M, Kp, P(pack_all[s * sb:(s + 1) * sb]), N, P(bias[s * N:(s + 1) * N]), P(ref), pitch, 4 + s * N,
                                     _lib.PRECISIONS["bf16x3"], st), "eml_gemm_bf16")
    assert torch.equal(one, ref)
    want = A.double() @ W.double().t() + bias.double()
    assert float((one[:, 4:].double() - want).abs().max()) <= 2e-5 * float(want.abs().max())
    # argument checks: slices must be whole 16-column groups and fit the row pitch
    assert lib.eml_gemm_bf16_slices(P(hi), P(lo), M, Kp, P(pack_all), sb, nsl, N, P(bias), P(one), pitch - 8, 4, _lib.PRECISIONS["bf16x3"], 1, st) < 0
    assert lib.eml_gemm_bf16_slices(P(hi), P(lo), M, Kp, P(pack_all), sb, nsl, N - 4, P(bias), P(one), pitch, 4, _lib.PRECISIONS["bf16x3"], 1, st) < 0
    # one pack launch for all slices == per-slice packs, byte for byte; split-K over the slices matches to summation order
    again = torch.zeros_like(pack_all)
    _lib.check(lib.eml_gemm_pack_slices(P(Wp), P(again), nsl, N, Kp, sb, st), "eml_gemm_pack_slices")
    assert torch.equal(again, pack_all)
    sk = torch.zeros(M, pitch, device=cuda)
    _lib.check(lib.eml_gemm_bf16_slices(P(hi), P(lo), M, Kp, P(pack_all), sb, nsl, N, P(bias), P(sk), pitch, 4, _lib.PRECISIONS["bf16x3"],
                                        min(3, Kp // 64), st), "eml_gemm_bf16_slices(split-K)")
    assert float((sk[:, 4:].double() - want).abs().max()) <= 2e-5 * float(want.abs().max())


def test_gemm_split_k_matches_single_pass(cuda, lib):
    from emlight_b200 import _lib
    P, st = _lib.ptr, _lib.stream_ptr()
    M, K, N = 48, 8192, 200
    A, W, bias, hi, lo, Wp, Kp = _operands(cuda, M, K, N, 5)
    pack = _pack(lib, Wp, 0, N, Kp)
    out = torch.zeros(M, N, device=cuda)
    _lib.check(lib.eml_gemm_bf16_splitk(P(hi), P(lo), M, Kp, P(pack), N, P(bias), P(out), N, 0, _lib.PRECISIONS["bf16x3"], 37, st), "splitk")
    want = A.double() @ W.double().t() + bias.double()
    assert float((out.double() - want).abs().max()) <= 2e-5 * float(want.abs().max())


@pytest.mark.parametrize("M,K,N", [(1, 8192, 256), (7, 1023, 10), (16, 8192, 256), (17, 512, 100), (200, 1024, 37)])
def test_linear_fp32(cuda, lib, M, K, N):
    """eml_linear_fp32 (nn.Linear: fc layers and heads, DenseNet.py:139-150, generator.py:124): the few-rows kernel (M <= 16) and the
    tiled one, ragged K / N."""
    from emlight_b200 import _lib
    P, st = _lib.ptr, _lib.stream_ptr()
    gen = torch.Generator().manual_seed(M * 31 + N)
    a = torch.randn(M, K, generator=gen).to(cuda)
    w = (torch.randn(N, K, generator=gen) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=gen).to(cuda)
    out = torch.full((M, N), float("nan"), device=cuda)
    _lib.check(lib.eml_linear_fp32(P(a), P(w), P(b), P(out), M, N, K, st), "eml_linear_fp32")
    want = a.double() @ w.double().t() + b.double()
    assert float((out.double() - want).abs().max()) <= 1e-5 * float(want.abs().max())
    _lib.check(lib.eml_linear_fp32(P(a), P(w), None, P(out), M, N, K, st), "eml_linear_fp32")
    assert float((out.double() - (want - b.double())).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.parametrize("N,K", [(256, 1152), (100, 70), (16, 64), (37, 9216)])
def test_matrix_pack_equals_the_elementwise_pack(cuda, lib, N, K, monkeypatch):
    """eml_conv_pack_weights with taps = 1 runs the 16-byte-chunk kernel; EML_PACK_V1=1 keeps the element-per-thread one: same bytes."""
    from emlight_b200 import _lib
    gen = torch.Generator().manual_seed(N + K)
    W = torch.randn(N, K, generator=gen).to(cuda)
    nbytes = lib.eml_conv_wpack_bytes(N, K, 1)
    a = torch.full((nbytes,), 0xAB, dtype=torch.uint8, device=cuda)
    b = torch.full((nbytes,), 0xCD, dtype=torch.uint8, device=cuda)
    _lib.check(lib.eml_conv_pack_weights(_lib.ptr(W), _lib.ptr(a), N, K, 1, _lib.stream_ptr()), "pack")
    monkeypatch.setenv("EML_PACK_V1", "1")
    _lib.check(lib.eml_conv_pack_weights(_lib.ptr(W), _lib.ptr(b), N, K, 1, _lib.stream_ptr()), "pack v1")
    assert torch.equal(a, b)
