"""Spherical needlets on sm_100a -- the hot path of the reference's ``Needlets/`` directory behind its own function names.

Reference (paths under Needlets/):
  ``fun_b``                                   sphere_needlets.py:10-29    C-infinity window b(x) (host, scipy.integrate.quad like the reference)
  ``spneedlet_pair(jmax)``                    sphere_needlets.py:107-127  antipodal cubature pairs
  ``SNvertex(theta, phi, jmax)``              sphere_needlets.py:196-238  -> (SN_matrix1, SN_matrix2, SN_matrix), float64
  ``getSolidAngleMap(width)``                 utils.py:35-50
  projection / sparsification / reconstruction   gt_gen_j3.py:39-43, mat_gen2.py:36-51, 55 -> ``NeedletTransform``

What runs where: the window function and the HEALPix cubature points (a few hundred numbers) are host-side numpy; the
(grid points x 1021) basis matrix -- hours of Python loops in the reference, which is why it ships the call commented out and
loads ``SN_Matrix3.npy`` instead (mat_gen2.py:27-30) -- is one CUDA kernel (``eml_needlet_basis``: addition theorem + Legendre
recurrence in float64); projection and reconstruction are tcgen05 GEMMs (``eml_gemm_bf16``, bf16x3 = fp32-grade) on operands packed
once per transform.  healpy is not needed: the RING-scheme pixel centres are evaluated in closed form (Gorski et al. 2005).
No CPU fallback: tensors live on the GPU and a missing library raises.
"""
import math

import numpy as np
import torch
from scipy.integrate import quad

from . import _lib

PANO_H, PANO_W = 128, 256


# --------------------------------------------------------------------------------------------- host-side tables (tiny)
def _f2(u):
    g = lambda x: math.exp(-1.0 / (1.0 - x * x))                                      # noqa: E731
    return quad(g, -1, u + 1e-10)[0] / quad(g, -1, 1)[0]


def _f3(x, B):
    if x <= 1.0 / B:
        return 1.0
    if x <= 1:
        return _f2(1 - 2 * B / (B - 1) * (x - 1 / B))
    return 0.0


def fun_b(x, B=2.0):
    """Window function of sphere_needlets.py:28-29."""
    return math.sqrt(_f3(x / B, B) - _f3(x, B))


def level_nside(j, B=2.0):
    return 2 ** math.ceil(math.log(math.floor(B ** (j + 1)) / 2, 2))                  # sphere_needlets.py:48


def healpix_centres(nside):
    """(npix, 3) float64 unit vectors of the HEALPix RING-scheme pixel centres (what healpy.pix2vec returns), per pixel in closed
    form: north cap p < 2n(n-1), equatorial belt, south cap by mirror symmetry."""
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    p = np.arange(npix, dtype=np.int64)
    z = np.empty(npix)
    phi = np.empty(npix)
    north = p < ncap
    south = p >= npix - ncap
    belt = ~(north | south)
    # polar caps: ring i (1-based) holds 4i pixels and starts at 2i(i-1)
    for mask, q, sign in ((north, p, 1.0), (south, npix - 1 - p, -1.0)):
        qq = q[mask]
        i = np.floor((1 + np.sqrt(1 + 2.0 * qq)) / 2).astype(np.int64)
        i = np.where(2 * i * (i - 1) > qq, i - 1, i)
        i = np.where(2 * (i + 1) * i <= qq, i + 1, i)
        jj = qq - 2 * i * (i - 1)                                                     # 0-based position in the ring
        z[mask] = sign * (1.0 - i * i / (3.0 * nside * nside))
        ph = (jj + 0.5) * np.pi / (2.0 * i)
        phi[mask] = ph if sign > 0 else 2 * np.pi - ph                                # mirrored rings run backwards from the end
    ip = p[belt] - ncap
    ring = ip // (4 * nside) + nside
    jj = ip % (4 * nside)
    shift = np.where(((ring + nside) & 1) == 1, 0.0, 0.5)                             # fodd = 1 (odd) / 0.5 (even): phi = (j + 1 - fodd) ...
    z[belt] = (2.0 * nside - ring) * 2.0 / (3.0 * nside)
    phi[belt] = (jj + shift) * np.pi / (2.0 * nside)
    s = np.sqrt(np.clip(1.0 - z * z, 0.0, None))
    return np.stack((s * np.cos(phi), s * np.sin(phi), z), 1)


def cubature_points(jmax, B=2.0):
    """All levels concatenated: (K,3) centres and (K,) level index (sphere_needlets.py:109-116)."""
    pts = [healpix_centres(level_nside(j, B)) for j in range(jmax + 1)]
    lev = np.concatenate([np.full(len(p), j, dtype=np.int32) for j, p in enumerate(pts)])
    return np.concatenate(pts, 0), lev


def level_coefficients(jmax, B=2.0):
    """c[j][l] = sqrt(lambda_j) b(l/B^j) (2l+1)/(4 pi) inside the level's band [ceil(B^(j-1)), min(floor(B^(j+1)), lmax)], else 0."""
    lmax = int(math.floor(B ** (jmax + 1)))
    c = np.zeros((jmax + 1, lmax + 1))
    for j in range(jmax + 1):
        lamb = 4 * math.pi / (12 * level_nside(j, B) ** 2)
        l_st, l_en = int(math.ceil(B ** (j - 1))), int(min(math.floor(B ** (j + 1)), lmax))
        for l in range(l_st, l_en + 1):
            c[j, l] = math.sqrt(lamb) * fun_b(l / 2.0 ** j, 2.0) * (2 * l + 1) / (4 * math.pi)
    return c


def spneedlet_pair(jmax, B=2.0):
    """sphere_needlets.py:107-127: index of the antipodal cubature point of every point, and the representatives."""
    pix, _ = cubature_points(jmax, B)
    corr = pix @ pix.T
    pair, use = [], []
    for i in range(pix.shape[0]):
        p = int(np.where(corr[i] + 1 < 1e-10)[0][0])
        pair.append(p)
        if p > i:
            use.append(i)
    return pair, use


def getSolidAngleMap(width):
    """utils.py:35-50 (numpy, host): (width/2, width) solid angle of every equirect pixel."""
    height = int(width / 2)
    y = np.arange(0, height)
    theta = (1.0 - ((y + 0.5) / height)) * np.pi
    sa = (np.pi * 2) / width * (np.cos(theta - (np.pi / height / 2.0)) - np.cos(theta + (np.pi / height / 2.0)))
    return np.repeat(sa[:, np.newaxis], width, axis=1)


# --------------------------------------------------------------------------------------------- device: basis matrix
def needlet_matrix(theta, phi, jmax, B=2.0, device=None):
    """SN_matrix (npoints, 1 + sum_j Npix_j) float64 on the GPU for grid points (theta[k], phi[k])."""
    lib = _lib.load()
    device = torch.device(device if device is not None else "cuda")
    if device.type != "cuda":
        raise RuntimeError("emlight_b200 runs on CUDA tensors only (sm_100a kernels; no CPU fallback)")
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    phi = np.asarray(phi, dtype=np.float64).reshape(-1)
    if theta.shape != phi.shape:
        raise ValueError("theta and phi must have the same length")
    xyz = np.stack((np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)), 1)
    centres, lev = cubature_points(jmax, B)
    coef = level_coefficients(jmax, B)
    with torch.cuda.device(device):
        d_xyz = torch.from_numpy(np.ascontiguousarray(xyz)).to(device)
        d_c = torch.from_numpy(np.ascontiguousarray(centres)).to(device)
        d_lev = torch.from_numpy(lev).to(device)
        d_coef = torch.from_numpy(np.ascontiguousarray(coef)).to(device)
        K = centres.shape[0]
        out = torch.empty(len(theta), K + 1, dtype=torch.float64, device=device)
        _lib.check(lib.eml_needlet_basis(_lib.ptr(d_xyz), len(theta), _lib.ptr(d_c), _lib.ptr(d_lev), K, _lib.ptr(d_coef),
                                         coef.shape[0], coef.shape[1] - 1, _lib.ptr(out), K + 1, _lib.stream_ptr()), "eml_needlet_basis")
    return out


def SNvertex(theta, phi, jmax, B=2.0, device=None):
    """sphere_needlets.py:196-238: returns (SN_matrix1, SN_matrix2, SN_matrix) like the reference -- [Y_00 | psi at the pair
    representatives], [Y_00 | psi at their antipodes], [Y_00 | all psi]."""
    SN = needlet_matrix(theta, phi, jmax, B, device)
    pair, use = spneedlet_pair(jmax, B)
    dev = SN.device
    use_i = torch.tensor([0] + [u + 1 for u in use], device=dev)
    pair_i = torch.tensor([0] + [pair[u] + 1 for u in use], device=dev)
    return SN.index_select(1, use_i), SN.index_select(1, pair_i), SN


def pano_grid(h=PANO_H, w=PANO_W):
    """theta / phi of the reference's evaluation grid (mat_gen2.py:22-25: endpoint-inclusive linspace, row-major)."""
    X, Y = np.meshgrid(np.linspace(0, 2, w) * np.pi, np.linspace(0, 1, h) * np.pi)
    return Y.reshape(-1), X.reshape(-1)


# --------------------------------------------------------------------------------------------- device: projection / reconstruction
class NeedletTransform:
    """Needlet analysis / synthesis of equirect panoramas (gt_gen_j3.py:39-43, mat_gen2.py:36-55) for a whole batch at once.

    ``project(pano)``:      pano (B, h*w, 3) [the reference's ``im.reshape((-1, 3))``] or (B, 3, h, w) -> coef (B, nCoeffs, 3),
                            coef[b,i,ch] = sum_p pano[b,p,ch] * SN[p,i] * omega[p]
    ``sparsify(coef)``:     mat_gen2.py:43-51, the j = jmax and j = jmax-1 blocks keep only |c| > 0.1 max|c| (per image)
    ``reconstruct(coef)``:  rec (B, h*w, 3) = SN @ coef
    """

    def __init__(self, jmax=3, h=PANO_H, w=PANO_W, B=2.0, device=None, precision="bf16x3"):
        if precision not in ("bf16x3", "bf16"):
            raise ValueError("precision must be 'bf16x3' (fp32-grade) or 'bf16'")
        lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        self.h, self.w, self.P, self.jmax = h, w, h * w, jmax
        self.precision = _lib.PRECISIONS[precision]
        self._split = precision == "bf16x3"
        theta, phi = pano_grid(h, w)
        self.SN = needlet_matrix(theta, phi, jmax, B, self.device)                    # (P, n) float64
        self.n = self.SN.shape[1]
        npix = [12 * level_nside(j, B) ** 2 for j in range(jmax + 1)]
        off = np.concatenate(([1], 1 + np.cumsum(npix)))
        self.level_slices = [(int(off[j]), int(off[j + 1])) for j in range(jmax + 1)]
        self.omega = torch.from_numpy(getSolidAngleMap(w).reshape(-1)).to(self.device)        # float64 (P,)
        st = _lib.stream_ptr()
        with torch.cuda.device(self.device):
            # projection weights: rows = coefficients, K = pixels, in slices of <= 256 rows
            self._Kp_pix = (self.P + 63) // 64 * 64
            wproj = (self.SN * self.omega[:, None]).t().contiguous().float()           # (n, P)
            self._proj_packs = []
            for n0 in range(0, self.n, 256):
                rows = min(256, self.n - n0)
                buf = torch.empty(lib.eml_conv_wpack_bytes(rows, self.P, 1), dtype=torch.uint8, device=self.device)
                _lib.check(lib.eml_conv_pack_weights(_lib.ptr(wproj[n0:n0 + rows]), _lib.ptr(buf), rows, self.P, 1, st),
                           "eml_conv_pack_weights(needlet projection)")
                self._proj_packs.append((n0, rows, buf))
            # reconstruction operand: SN rows (pixels) x K = coefficients, bf16 hi / lo
            self._Kp_n = (self.n + 63) // 64 * 64
            sn32 = self.SN.float().contiguous()
            self._sn_hi = torch.empty(self.P, self._Kp_n, dtype=torch.bfloat16, device=self.device)
            self._sn_lo = torch.empty_like(self._sn_hi) if self._split else None
            _lib.check(lib.eml_split_bf16(_lib.ptr(sn32), self.P, self.n, self.n, _lib.ptr(self._sn_hi), _lib.ptr(self._sn_lo),
                                          self._Kp_n, st), "eml_split_bf16(SN)")
            torch.cuda.current_stream().synchronize()                                 # sn32 / wproj may be freed after this

    def _planes(self, pano):
        _lib.require_cuda(pano)
        if pano.dim() == 4 and pano.shape[1] == 3 and pano.shape[2] * pano.shape[3] == self.P:
            return pano.reshape(pano.shape[0] * 3, self.P).float().contiguous()       # NCHW: already (B*3, P)
        if pano.dim() == 3 and pano.shape[1] == self.P and pano.shape[2] == 3:
            return pano.float().permute(0, 2, 1).reshape(-1, self.P).contiguous()     # the reference's (P, 3) per image
        raise ValueError("expected pano (B, %d, 3) or (B, 3, %d, %d), got %s" % (self.P, self.h, self.w, tuple(pano.shape)))

    @torch.no_grad()
    def project(self, pano):
        lib = _lib.load()
        st = _lib.stream_ptr()
        planes = self._planes(pano)                                                   # (M = B*3, P)
        M = planes.shape[0]
        a_hi = torch.empty(M, self._Kp_pix, dtype=torch.bfloat16, device=self.device)
        a_lo = torch.empty_like(a_hi) if self._split else None
        _lib.check(lib.eml_split_bf16(_lib.ptr(planes), M, self.P, self.P, _lib.ptr(a_hi), _lib.ptr(a_lo), self._Kp_pix, st),
                   "eml_split_bf16(pano)")
        pitch = (self.n + 3) // 4 * 4
        # short and deep (M = 3B rows, K = 32768 pixels): split K so that every SM gets a piece (partial sums by float atomics)
        out = torch.zeros(M, pitch, dtype=torch.float32, device=self.device)
        mtiles = (M + 127) // 128
        ksplit = max(1, min(self._Kp_pix // 64, 148 // mtiles))
        for n0, rows, buf in self._proj_packs:
            _lib.check(lib.eml_gemm_bf16_splitk(_lib.ptr(a_hi), _lib.ptr(a_lo), M, self._Kp_pix, _lib.ptr(buf), rows, None, _lib.ptr(out),
                                                pitch, n0, self.precision, ksplit, st), "eml_gemm_bf16_splitk(needlet projection)")
        return out[:, :self.n].reshape(-1, 3, self.n).permute(0, 2, 1).contiguous()    # (B, n, 3)

    @torch.no_grad()
    def sparsify(self, coef, frac=0.1, levels=None):
        lib = _lib.load()
        _lib.require_cuda(coef)
        if coef.dim() != 3 or coef.shape[1] != self.n:
            raise ValueError("expected coef (B, %d, ch)" % self.n)
        levels = levels if levels is not None else [j for j in (self.jmax, self.jmax - 1) if j >= 0]
        out = coef.float().contiguous().clone()
        ranges = torch.tensor([v for j in levels for v in self.level_slices[j]], dtype=torch.int32, device=self.device)
        _lib.check(lib.eml_needlet_sparsify(_lib.ptr(out), out.shape[0], self.n, out.shape[2], _lib.ptr(ranges), len(levels),
                                            float(frac), _lib.stream_ptr()), "eml_needlet_sparsify")
        return out

    @torch.no_grad()
    def reconstruct(self, coef):
        lib = _lib.load()
        st = _lib.stream_ptr()
        _lib.require_cuda(coef)
        if coef.dim() != 3 or coef.shape[1] != self.n or coef.shape[2] != 3:
            raise ValueError("expected coef (B, %d, 3)" % self.n)
        Bn = coef.shape[0]
        wrec = coef.float().permute(0, 2, 1).reshape(Bn * 3, self.n).contiguous()      # rows = (image, channel), K = coefficients
        N = Bn * 3
        pitch = (N + 3) // 4 * 4
        out = torch.empty(self.P, pitch, dtype=torch.float32, device=self.device)
        for n0 in range(0, N, 256):
            rows = min(256, N - n0)
            buf = torch.empty(lib.eml_conv_wpack_bytes(rows, self.n, 1), dtype=torch.uint8, device=self.device)
            _lib.check(lib.eml_conv_pack_weights(_lib.ptr(wrec[n0:n0 + rows]), _lib.ptr(buf), rows, self.n, 1, st),
                       "eml_conv_pack_weights(needlet coefficients)")
            _lib.check(lib.eml_gemm_bf16(_lib.ptr(self._sn_hi), _lib.ptr(self._sn_lo), self.P, self._Kp_n, _lib.ptr(buf), rows, None,
                                         _lib.ptr(out), pitch, n0, self.precision, st), "eml_gemm_bf16(needlet reconstruction)")
        return out[:, :N].reshape(self.P, Bn, 3).permute(1, 0, 2)                      # (B, P, 3) view, pixel-major like rec.reshape((h,w,3))
