"""CPU ORACLE for the video-frame preprocessing that feeds the hot path -- TEST INFRASTRUCTURE ONLY.

Restates, in numpy, what ``load_video`` does to the decoded uint8 frames (valley/util/data_util.py:271-281):

    TensorToNumpy -> Resize(256) -> CenterCrop(224) -> ClipToTensor(div 255) -> Normalize(CLIP mean/std)

* ``Resize(256)`` (valley/data/video_transform.py:269-277) keeps its default ``interpolation='nearest'``, and the PIL branch of
  ``resize_clip`` has the two names swapped (:63-66), so the frames are resized with **PIL.Image.BILINEAR**: Pillow's two-pass
  fixed-point convolution (the arithmetic lives in the third-party dependency Pillow, src/libImaging/Resample.c:
  ``precompute_coeffs``, ``normalize_coeffs_8bpc`` with PRECISION_BITS = 22, ``ImagingResampleHorizontal_8bpc`` then
  ``ImagingResampleVertical_8bpc``, each pass rounding to uint8).  A triangle filter whose support grows with the
  down-scaling factor -- NOT 2-tap bilinear interpolation.
* short side -> 256, long side ``int(256 * long / short)`` (video_transform.py:74-81); untouched when the short side is
  already 256 (:56-58).
* ``CenterCrop(224)``: ``x1 = int(round((w - 224) / 2.))`` with Python's round-half-to-even (:542-544).
* ``ClipToTensor``: uint8 -> float32 ``.div(255)`` (:139-163);  ``Normalize``: ``(x - mean) / std`` in float32 (:91-97).

Pinned how: ``oracle/make_golden_preprocess.py`` runs the reference's own transform classes (Pillow 12.2.0 underneath) on seeded
clips of several geometries and checks this restatement bit-for-bit (uint8 stage and float32 output); the reference's outputs
are committed under tests/golden/ and re-checked by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def resize_sizes(im_h: int, im_w: int, size: int = 256):
    """video_transform.py:56-60, :74-81 -> (new_h, new_w); unchanged when the short side already equals ``size``."""
    if (im_w <= im_h and im_w == size) or (im_h <= im_w and im_h == size):
        return im_h, im_w
    if im_w < im_h:
        return int(size * im_h / im_w), size
    return size, int(size * im_w / im_h)


def crop_origin(im_h: int, im_w: int, crop: int = 224):
    """video_transform.py:542-543 (Python round: half to even)."""
    return int(round((im_h - crop) / 2.)), int(round((im_w - crop) / 2.))


def bilinear_coeffs(in_size: int, out_size: int):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1.0).
    Returns (ksize, xmin[out], count[out], kk[out, ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            a = abs((x + lo - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        xmin[xx], cnt[xx] = lo, n
    return ksize, xmin, cnt, kk


def _pass(img: np.ndarray, axis: int, out_size: int) -> np.ndarray:
    """One 8-bit resampling pass along ``axis`` of an [H, W, C] uint8 image: sum(pixel * k) + half, >> 22, clip8."""
    in_size = img.shape[axis]
    ksize, xmin, cnt, kk = bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)                 # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], np.int64)
    for xx in range(out_size):
        taps = src[xmin[xx]: xmin[xx] + cnt[xx]]                      # [n, other, C]
        acc = (taps * kk[xx, : cnt[xx], None, None].astype(np.int64)).sum(0) + (1 << (PRECISION_BITS - 1))
        out[xx] = acc >> PRECISION_BITS
    return np.moveaxis(np.clip(out, 0, 255).astype(np.uint8), 0, axis)


def pil_bilinear_resize(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """Image.resize((new_w, new_h), PIL.Image.BILINEAR) on an [H, W, 3] uint8 array: horizontal pass over the rows the
    vertical pass will need, then the vertical pass (Resample.c ImagingResampleInner).  A pass whose size does not change
    is skipped by Pillow; its coefficients would be the identity, so running it gives the same bytes."""
    h, w, _ = img.shape
    out = img
    if new_w != w:
        out = _pass(out, 1, new_w)
    if new_h != h:
        out = _pass(out, 0, new_h)
    return out


def preprocess_frames(frames: np.ndarray, size: int = 256, crop: int = 224, mean=CLIP_MEAN, std=CLIP_STD,
                      return_uint8: bool = False) -> np.ndarray:
    """frames [T, H, W, 3] uint8 (what decord's get_batch returns, data_util.py:262-263) -> [T, 3, 224, 224] float32
    (the reference returns [3, T, 224, 224]; every caller permutes it to frames-first: model_worker.py:337, valley_model.py:430)."""
    T, H, W, _ = frames.shape
    nh, nw = resize_sizes(H, W, size)
    y1, x1 = crop_origin(nh, nw, crop)
    out = np.empty((T, 3, crop, crop), np.float32)
    u8 = np.empty((T, crop, crop, 3), np.uint8)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    for t in range(T):
        r = pil_bilinear_resize(frames[t], nh, nw) if (nh, nw) != (H, W) else frames[t]
        c = r[y1: y1 + crop, x1: x1 + crop]
        u8[t] = c
        x = c.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
        out[t] = (x - m) / s
    return u8 if return_uint8 else out


def fixed_frame_indices(video_len: int, n: int = 8) -> np.ndarray:
    """data_util.py:262: np.linspace(0, video_len - 1, n).astype(np.int_)."""
    return np.linspace(0, video_len - 1, n).astype(np.int_)


def fps_frame_indices(video_len: int, avg_fps: float, fps_number: float = 0.5) -> np.ndarray:
    """data_util.py:266-268: range(0, video_len, int(round(avg_fps) / fps_number))."""
    return np.arange(0, video_len, int(round(avg_fps) / fps_number))


def pil_pipeline(frames: np.ndarray, size: int = 256, crop: int = 224, mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """The same pipeline executed the way the reference executes it -- through Pillow itself (data_util.py:274-281:
    Image.fromarray -> Image.resize(PIL.Image.BILINEAR) -> Image.crop -> float32 / 255 -> (x - mean) / std), one frame
    at a time on one core.  Used (a) to re-check the numpy restatement above wherever Pillow is installed and (b) as the
    CPU baseline bench.py times next to the device kernels."""
    import torch
    from PIL import Image
    T, H, W, _ = frames.shape
    nh, nw = resize_sizes(H, W, size)
    y1, x1 = crop_origin(nh, nw, crop)
    clip = np.zeros([3, T, crop, crop])                               # ClipToTensor allocates float64 (video_transform.py:139)
    for t in range(T):
        img = Image.fromarray(np.uint8(frames[t])).convert("RGB")
        if (nh, nw) != (H, W):
            img = img.resize((nw, nh), Image.BILINEAR)
        img = img.crop((x1, y1, x1 + crop, y1 + crop))
        clip[:, t] = np.array(img).transpose(2, 0, 1)
    x = torch.from_numpy(clip).float().div(255)
    m = torch.as_tensor(mean, dtype=torch.float32)[:, None, None, None]
    s = torch.as_tensor(std, dtype=torch.float32)[:, None, None, None]
    return x.sub_(m).div_(s).permute(1, 0, 2, 3).contiguous().numpy()
