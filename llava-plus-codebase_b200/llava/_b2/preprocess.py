"""GPU image preprocessing for the LLaVA path — the device counterpart of the reference's `process_images`
(/root/reference/llava/mm_utils.py:16-44): `expand2square` (when `image_aspect_ratio == 'pad'`) followed by
`CLIPImageProcessor.preprocess` as the reference's pinned transformers 4.31 runs it (PIL bicubic resize of the shortest
edge, centre crop, rescale 1/255, normalise). Input: PIL images or uint8 HWC arrays of any size; output: `[n, 3, S, S]`
bf16 on the device, ready for `encode_images`.

Host work per image is O(side): the resize plan, i.e. PIL's resampling coefficient tables (`precompute_coeffs` +
`normalize_coeffs_8bpc` of Pillow's Resample.c, restated in double precision so that the device result equals
`Image.resize(..., BICUBIC)` bit for bit), cached per (input size, output size). The arithmetic on the pixels — two
fixed-point separable passes, the virtual padding, crop, rescale, normalise, bf16 cast — runs in csrc/preprocess.cu.
"""
import ctypes
import math
import threading

import numpy as np
import torch

from . import PreprocessPlan, check, init, ptr, stream_ptr, _vp

_PRECISION_BITS = 22  # Pillow: 32 - 8 - 2
_BICUBIC_SUPPORT = 2.0


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter: (bounds int32 [out,2], kk int32 [out,ksize])."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = _BICUBIC_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    one = float(1 << _PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            if ww != 0.0:
                v /= ww
            kk[xx, x] = int(-0.5 + v * one) if v < 0 else int(0.5 + v * one)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resize_geometry(H, W, out, pad):
    """(virtual source height/width, pad_top, pad_left, resized height/width, crop top/left) following expand2square
    (mm_utils.py:16-27) and CLIPImageProcessor's shortest-edge resize + centre crop."""
    if pad:
        Q = max(H, W)
        pad_top = (Q - H) // 2 if W > H else 0
        pad_left = (Q - W) // 2 if H > W else 0
        return Q, Q, pad_top, pad_left, out, out, 0, 0
    if H <= W:
        nh, nw = out, int(out * W / H)
    else:
        nh, nw = int(out * H / W), out
    return H, W, 0, 0, nh, nw, (nh - out) // 2, (nw - out) // 2


class ClipPreprocessor:
    """`process_images` on the device for a CLIP-style image processor configuration."""

    def __init__(self, image_processor, device="cuda", image_aspect_ratio=None):
        self.device = torch.device(device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.lib = init(self.index)
        ip = image_processor
        size = ip.size if isinstance(ip.size, dict) else dict(ip.size)
        side = size.get("shortest_edge") or size.get("height")
        crop = ip.crop_size if isinstance(ip.crop_size, dict) else dict(ip.crop_size)
        if not (getattr(ip, "do_resize", True) and getattr(ip, "do_center_crop", True) and getattr(ip, "do_normalize", True)
                and getattr(ip, "do_rescale", True)):
            raise NotImplementedError("the device path implements the CLIP preprocessing chain (resize, centre crop, rescale, normalise)")
        if crop.get("height") != side or crop.get("width") != side:
            raise NotImplementedError(f"crop size {crop} != resize edge {side}")
        if int(getattr(ip, "resample", 3)) != 3:
            raise NotImplementedError("only PIL BICUBIC (resample=3) has a kernel")
        self.out = int(side)
        self.mean = [float(x) for x in ip.image_mean]
        self.std = [float(x) for x in ip.image_std]
        self.rescale = float(getattr(ip, "rescale_factor", 1 / 255))
        self.pad = image_aspect_ratio == "pad"
        self.bg = [int(x * 255) for x in self.mean]  # mm_utils.py:36: tuple(int(x*255) for x in image_mean)
        self._tables = {}
        self._lock = threading.Lock()

    def _table(self, in_size, out_size):
        key = (in_size, out_size)
        with self._lock:
            t = self._tables.get(key)
            if t is None:
                if in_size == out_size:
                    t = (None, None, 0, None)
                else:
                    bounds, kk = resample_coeffs(in_size, out_size)
                    t = (torch.from_numpy(bounds).to(self.device), torch.from_numpy(kk).to(self.device), kk.shape[1], bounds)
                self._tables[key] = t
        return t

    def __call__(self, images, return_uint8=False):
        """images: list of PIL.Image / uint8 HWC arrays (or a single one). Returns bf16 [n,3,S,S] on the device
        (and, with return_uint8, the resized + cropped 8-bit images [n,S,S,3] the parity tests compare with PIL)."""
        if not isinstance(images, (list, tuple)):
            images = [images]
        n, S = len(images), self.out
        pixels = torch.empty(n, 3, S, S, dtype=torch.bfloat16, device=self.device)
        u8 = torch.empty(n, S, S, 3, dtype=torch.uint8, device=self.device) if return_uint8 else None
        with torch.cuda.device(self.index):
            for i, im in enumerate(images):
                a = np.asarray(im.convert("RGB") if hasattr(im, "convert") else im)
                if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8:
                    raise ValueError("images must be RGB uint8 (H, W, 3)")
                H, W = a.shape[:2]
                vh, vw, pad_top, pad_left, nh, nw, top, left = resize_geometry(H, W, S, self.pad)
                hb, hk, hks, _ = self._table(vw, nw)
                vb, vk, vks, vbounds = self._table(vh, nh)
                if vb is None:
                    y0, rows = top, S
                else:
                    lo = vbounds[top:top + S, 0]
                    hi = lo + vbounds[top:top + S, 1]
                    y0, rows = int(lo.min()), int(hi.max() - lo.min())
                src = torch.from_numpy(np.ascontiguousarray(a)).to(self.device, non_blocking=True)
                tmp = torch.empty(rows * S * 3, dtype=torch.uint8, device=self.device)
                pl = PreprocessPlan()
                pl.img, pl.H, pl.W, pl.pad_top, pl.pad_left = src.data_ptr(), H, W, pad_top, pad_left
                pl.bg = (ctypes.c_uint8 * 4)(*self.bg, 0)
                pl.h_bounds, pl.h_kk, pl.h_ksize, pl.h_identity = (hb.data_ptr() if hb is not None else None,
                                                                    hk.data_ptr() if hk is not None else None, hks, int(hb is None))
                pl.v_bounds, pl.v_kk, pl.v_ksize, pl.v_identity = (vb.data_ptr() if vb is not None else None,
                                                                    vk.data_ptr() if vk is not None else None, vks, int(vb is None))
                pl.y0, pl.rows, pl.x_lo, pl.y_lo, pl.out = y0, rows, left, top, S
                pl.tmp = tmp.data_ptr()
                pl.mean, pl.stdv, pl.rescale = (ctypes.c_float * 3)(*self.mean), (ctypes.c_float * 3)(*self.std), self.rescale
                pl.pixels = pixels[i].data_ptr()
                pl.u8_out = u8[i].data_ptr() if u8 is not None else None
                check(self.lib.b2_op_preprocess_clip(ctypes.byref(pl), stream_ptr()), "b2_op_preprocess_clip")
        # `src` / `tmp` are released here while the launches may still be queued: the caching allocator only hands their memory
        # to later work on this same stream, i.e. behind those launches
        return (pixels, u8) if return_uint8 else pixels


def process_images(images, image_processor, model_cfg, device="cuda", _cache={}):
    """Drop-in for `llava.mm_utils.process_images(images, image_processor, model_cfg)` that runs on the device and returns
    bf16 pixel_values already resident in HBM (same [n,3,S,S] layout; the reference returns fp32 on the host)."""
    ratio = getattr(model_cfg, "image_aspect_ratio", None)
    key = (id(image_processor), ratio, str(device))
    pre = _cache.get(key)
    if pre is None:
        pre = _cache[key] = ClipPreprocessor(image_processor, device=device, image_aspect_ratio=ratio)
    return pre(images)
