"""Image preprocessing on the device (VERDICT r1 missing #3): csrc/preprocess.cu + llava/_b2/preprocess.py against
the reference's arithmetic — `expand2square` (llava/mm_utils.py:16-27) and PIL's bicubic resize, which is what
CLIPImageProcessor.preprocess runs under the reference's pinned transformers 4.31. The 8-bit resized image must equal PIL's
BIT FOR BIT; the normalised bf16 output must be the bf16 rounding of the fp32 pipeline; the INSTALLED transformers 5.5
processor (torchvision backend, float antialiasing) is compared as well and differs from PIL itself by up to 2 grey levels.

CPU part: the host-side coefficient tables (Pillow's precompute_coeffs restated) drive a numpy version of the two passes."""
import json
import os
import tempfile

import numpy as np
import pytest
import torch
from PIL import Image

from llava._b2 import preprocess as PP

MEAN = [0.48145466, 0.4578275, 0.40821073]
STD = [0.26862954, 0.26130258, 0.27577711]
SIZES = [(480, 640), (1000, 700), (100, 80), (336, 336), (336, 500), (700, 336), (123, 457), (37, 41), (2000, 1500)]


def _images(seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i, (H, W) in enumerate(SIZES):
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        if i % 2:  # natural-like: low-pass noise
            a = np.asarray(Image.fromarray(a).resize((max(W // 8, 2), max(H // 8, 2))).resize((W, H), Image.BICUBIC))
        out.append(a)
    return out


def expand2square(pil_img, background_color):  # the reference's function (mm_utils.py:16-27), restated
    width, height = pil_img.size
    if width == height:
        return pil_img
    side = max(width, height)
    result = Image.new(pil_img.mode, (side, side), background_color)
    result.paste(pil_img, (0, (width - height) // 2) if width > height else ((height - width) // 2, 0))
    return result


def pil_reference(a, out, pad):
    """uint8 [out,out,3] through PIL, as the reference's process_images + transformers-4.31 CLIPImageProcessor compute it."""
    img = Image.fromarray(a)
    if pad:
        img = expand2square(img, tuple(int(x * 255) for x in MEAN))
    W, H = img.size
    nh, nw = (out, int(out * W / H)) if H <= W else (int(out * H / W), out)
    r = np.asarray(img.resize((nw, nh), resample=Image.BICUBIC))
    top, left = (nh - out) // 2, (nw - out) // 2
    return r[top:top + out, left:left + out]


def normalise(u8):
    x = u8.astype(np.float32) * np.float32(1 / 255)
    return ((x - np.array(MEAN, np.float32)) / np.array(STD, np.float32)).transpose(2, 0, 1)


def _numpy_passes(a, out, pad):
    """the device algorithm in numpy, driven by the PRODUCT's host tables (resample_coeffs / resize_geometry)."""
    H, W = a.shape[:2]
    vh, vw, pad_top, pad_left, nh, nw, top, left = PP.resize_geometry(H, W, out, pad)
    src = np.empty((vh, vw, 3), np.uint8)
    src[:] = np.array([int(x * 255) for x in MEAN], np.uint8)
    src[pad_top:pad_top + H, pad_left:pad_left + W] = a

    def axis_pass(img, axis, n_out):
        n_in = img.shape[axis]
        if n_in == n_out:
            return img
        b, kk = PP.resample_coeffs(n_in, n_out)
        x = np.moveaxis(img, axis, 0).astype(np.int64)
        o = np.zeros((n_out,) + x.shape[1:], np.int64)
        for i in range(n_out):
            lo, n = b[i]
            o[i] = np.clip(((1 << 21) + np.tensordot(kk[i, :n].astype(np.int64), x[lo:lo + n], axes=(0, 0))) >> 22, 0, 255)
        return np.moveaxis(o.astype(np.uint8), 0, axis)

    r = axis_pass(axis_pass(src, 1, nw), 0, nh)
    return r[top:top + out, left:left + out]


@pytest.mark.parametrize("pad", [False, True])
def test_host_tables_reproduce_pil_bit_for_bit(pad):
    for a in _images()[:7]:
        assert np.array_equal(_numpy_passes(a, 336, pad), pil_reference(a, 336, pad)), (a.shape, pad)
    small = _images(3)[5]
    assert np.array_equal(_numpy_passes(small, 56, pad), pil_reference(small, 56, pad))


def _processor(size=336):
    from transformers import CLIPImageProcessor

    d = tempfile.mkdtemp(prefix="b2pp_")
    with open(os.path.join(d, "preprocessor_config.json"), "w") as f:
        json.dump({"crop_size": size, "do_center_crop": True, "do_normalize": True, "do_resize": True, "image_mean": MEAN,
                   "image_std": STD, "resample": 3, "size": size, "image_processor_type": "CLIPImageProcessor"}, f)
    return CLIPImageProcessor.from_pretrained(d)


@pytest.mark.gpu
@pytest.mark.parametrize("pad", [False, True])
def test_device_preprocessing_equals_pil_and_the_fp32_pipeline(pad):
    ip = _processor()
    pre = PP.ClipPreprocessor(ip, device="cuda", image_aspect_ratio="pad" if pad else None)
    imgs = _images()
    pixels, u8 = pre([Image.fromarray(a) for a in imgs], return_uint8=True)
    assert pixels.shape == (len(imgs), 3, 336, 336) and pixels.dtype == torch.bfloat16
    for i, a in enumerate(imgs):
        ref8 = pil_reference(a, 336, pad)
        assert np.array_equal(u8[i].cpu().numpy(), ref8), (a.shape, pad)                      # PIL, bit for bit
        want = torch.from_numpy(normalise(ref8)).to(torch.bfloat16)
        assert torch.equal(pixels[i].cpu(), want), (a.shape, pad)                             # bf16(fp32 pipeline), exactly
    # the installed transformers processor (torchvision backend) differs from PIL itself by <= 2 grey levels
    if not pad:
        hf = ip.preprocess([Image.fromarray(a) for a in imgs], return_tensors="pt")["pixel_values"]
        tol = 2.0 / 255 / min(STD) + 2 ** -7 * 2.7
        assert float((hf - pixels.float().cpu()).abs().max()) <= tol


@pytest.mark.gpu
def test_process_images_drop_in_feeds_encode_images():
    from types import SimpleNamespace

    ip = _processor(56)
    cfg = SimpleNamespace(image_aspect_ratio="pad")
    imgs = [Image.fromarray(a) for a in _images(2)[:3]]
    got = PP.process_images(imgs, ip, cfg)
    assert got.shape == (3, 3, 56, 56) and got.is_cuda
    for i, im in enumerate(imgs):
        want = torch.from_numpy(normalise(pil_reference(np.asarray(im), 56, True))).to(torch.bfloat16)
        assert torch.equal(got[i].cpu(), want)
