"""HDR tonemapping on the GPU -- drop-in for ``TonemapHDR`` of the reference (RegressionNetwork/util.py:36-66; the step right in front
of the DenseNet: ``data.py:62-73`` tonemaps every crop and scales the intensity / ambient targets by the returned ``alpha``).
SURVEY 8f rank 2.  Same constructor and call signature.  A numpy ``(H, W, C)`` image -- what the reference's callers pass
(``train.py:124,135``: ``tone(env)[0].transpose(...).astype('float32')``) -- is uploaded, tone-mapped by the kernel and comes back as
``(numpy float32 array, float alpha)`` exactly like the reference; a CUDA tensor ``(H, W, C)`` or batch ``(B, H, W, C)`` stays on the
device and returns ``(float32 tensor, alpha)`` -- ``alpha`` a Python float for one image, a ``(B,)`` tensor for a batch.  The percentile is an exact per-image radix selection (``eml_tonemap_hdr``), interpolated like ``np.percentile``."""
import numpy as np
import torch

from . import _lib


class TonemapHDR:
    def __init__(self, gamma=2.4, percentile=50, max_mapping=0.5):
        self.gamma = gamma
        self.percentile = percentile
        self.max_mapping = max_mapping

    @_lib.on_tensor_device
    @torch.no_grad()
    def __call__(self, img, clip=True, alpha=None, gamma=True):
        lib = _lib.load()
        if isinstance(img, np.ndarray):                 # the reference's calling convention: numpy in -> numpy out (no CPU arithmetic:
            if not torch.cuda.is_available():           # the array is uploaded and the kernel does the work)
                raise RuntimeError("emlight_b200.TonemapHDR needs a CUDA device (sm_100a kernel; no CPU fallback)")
            y, a = self(torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).cuda(), clip=clip, alpha=alpha, gamma=gamma)
            return y.cpu().numpy(), (a if not torch.is_tensor(a) else a.cpu().numpy())
        _lib.require_cuda(img)
        single = img.dim() == 3
        if img.dim() not in (3, 4):
            raise ValueError("expected an image (H, W, C) or a batch (B, H, W, C), got %s" % (tuple(img.shape),))
        x = (img[None] if single else img).float().contiguous()
        B = x.shape[0]
        per = x[0].numel()
        out = torch.empty_like(x)
        given = alpha is not None
        if given:
            a = torch.as_tensor(alpha, dtype=torch.float32, device=x.device).reshape(-1)
            a = a.expand(B).contiguous() if a.numel() == 1 else a.contiguous()
            if a.numel() != B:
                raise ValueError("alpha must be a scalar or have one value per image")
        else:
            a = torch.empty(B, dtype=torch.float32, device=x.device)
        _lib.check(lib.eml_tonemap_hdr(_lib.ptr(x), _lib.ptr(out), _lib.ptr(a), B, per, float(self.gamma), float(self.percentile),
                                       float(self.max_mapping), int(bool(gamma)), int(bool(clip)), int(given), _lib.stream_ptr()),
                   "eml_tonemap_hdr")
        if single:
            return out[0], (alpha if given else float(a[0]))
        return out, (alpha if given else a)
