"""Frame selection + preprocessing in front of the vision tower: the device-side replacement of ``load_video``
(valley/util/data_util.py:249-303, file branch).

Decoding the container (decord) stays with the caller -- pass any reader with decord's surface (``len()``, ``get_batch(idx)``
returning [n,H,W,3] uint8, ``get_avg_fps()``) or the uint8 frames themselves.  Everything after decoding runs on the GPU through
``vly_preprocess_frames`` and is bit-exact with the reference's PIL pipeline (Resize(256) with PIL BILINEAR -> CenterCrop(224)
-> /255 -> CLIP mean/std).  The frame-index arithmetic is the reference's own numpy expressions (exact integer logic).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

_DT = {torch.float32: _lib.VLY_F32, torch.bfloat16: _lib.VLY_BF16, torch.float16: _lib.VLY_F16}


def fixed_frame_indices(video_len: int, fixed_frame_number: int = 8) -> np.ndarray:
    """data_util.py:262-263."""
    return np.linspace(0, video_len - 1, fixed_frame_number).astype(np.int_)


def fps_frame_indices(video_len: int, avg_fps: float, fps_number: float = 0.5) -> np.ndarray:
    """data_util.py:266-268."""
    return np.arange(0, video_len, int(round(avg_fps) / fps_number))


def preprocess_plan(height: int, width: int):
    """(new_h, new_w, crop_y, crop_x) of Resize(256) + CenterCrop(224) (video_transform.py:56-60, :74-81, :542-543).  No GPU."""
    v = [C.c_int() for _ in range(4)]
    check(_lib.load().vly_preprocess_plan(height, width, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)


def resample_coeffs(in_size: int, out_size: int):
    """Pillow's 8-bit triangle-filter tables (ksize, xmin[out], count[out], kk[out,ksize]) as the device uses them.  No GPU."""
    lib, k = _lib.load(), C.c_int()
    check(lib.vly_resample_coeffs(in_size, out_size, C.byref(k), None, None, None))
    xmin, cnt = np.zeros(out_size, np.int32), np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, k.value), np.int32)
    as_p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    check(lib.vly_resample_coeffs(in_size, out_size, C.byref(k), as_p(xmin), as_p(cnt), as_p(kk)))
    return k.value, xmin, cnt, kk


def preprocess_frames(model, frames: torch.Tensor, dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """frames [T,H,W,3] uint8 (host or device) -> [T,3,224,224] ``dtype`` on the model's device.
    fp16 is what the reference's callers cast to (model_worker.py:336, valley_model.py:430)."""
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError(f"frames must be [T,H,W,3] uint8, got {tuple(frames.shape)} {frames.dtype}")
    f = frames.to(model.device, non_blocking=True).contiguous()
    T, H, W, _ = f.shape
    out = torch.empty(T, 3, 224, 224, dtype=dtype, device=model.device)
    st = torch.cuda.current_stream(model.device).cuda_stream
    check(model._lib.vly_preprocess_frames(model._ctx, f.data_ptr(), T, H, W, _DT[dtype], out.data_ptr(), st))
    return out


def load_video(model, reader, frame_mode: str = "fixed", fixed_frame_number: int = 8, fps_number: float = 0.5,
               dtype: torch.dtype = torch.float16) -> torch.Tensor:
    """``load_video`` for an already-opened reader (decord.VideoReader surface) -> [T,3,224,224] on the device,
    i.e. ``load_video(path).permute(1,0,2,3).half()`` of the reference without the CPU round trip."""
    n = len(reader)
    if frame_mode == "fixed":
        idx = fixed_frame_indices(n, fixed_frame_number)
    elif frame_mode == "fps":
        idx = fps_frame_indices(n, reader.get_avg_fps(), fps_number)
    else:
        raise ValueError('Frame mode is only support "fps" or "fixed"')
    frames = reader.get_batch(idx)
    frames = torch.as_tensor(np.asarray(frames) if not torch.is_tensor(frames) else frames)
    return preprocess_frames(model, frames.to(torch.uint8), dtype)


# ---- the directory-of-images branch of load_video (data_util.py:282-302): CPU work, library code on both sides ----
def select_image_dir_frames(path, frame_mode: str = "fixed", fixed_frame_number: int = 8):
    """The files ``load_video`` opens for a directory: ``Path(path).rglob('*')`` in directory order (data_util.py:283), then
    ``np.linspace(0, n - 1, fixed_frame_number)`` of them (:284-286); 'fps' and anything else raise as the reference does (:287-290)."""
    from pathlib import Path
    files = list(Path(path).rglob('*'))
    if frame_mode == 'fixed':
        return [files[i] for i in np.linspace(0, len(files) - 1, fixed_frame_number).astype(np.int_)]
    if frame_mode == 'fps':
        raise ValueError('Input folder is not support this frame mode')
    raise ValueError('Frame mode is only support "fps" or "fixed"')


def load_image_dir(path, image_processor=None, frame_mode: str = "fixed", fixed_frame_number: int = 8,
                   frame_process_method: str = "centercrop") -> torch.Tensor:
    """``load_video(path_to_a_directory_of_frames, image_processer, ...)`` (data_util.py:282-302) -> [T,3,224,224] float32 on the
    host, FRAMES FIRST (the reference permutes to [3,T,...] and every caller permutes back).  Same calls as the reference: PIL
    ``Image.open``, the optional square resize of ``frame_process_method='resize'`` (torchvision's Resize on a PIL image ==
    ``Image.resize(..., BILINEAR)``), then the HF ``CLIPImageProcessor.preprocess`` (bicubic shortest-edge 224, centre crop,
    1/255, CLIP mean/std).  ``image_processor=None`` builds the default ``CLIPImageProcessor()`` -- the reference would fail with
    ``None.preprocess`` there (its ``completion`` never passes one)."""
    from PIL import Image
    frames = [Image.open(str(f)) for f in select_image_dir_frames(path, frame_mode, fixed_frame_number)]
    if frame_process_method == 'resize':
        m = min(frames[0].size)
        frames = [f.resize((m, m), Image.BILINEAR) for f in frames]
    if image_processor is None:
        from transformers import CLIPImageProcessor
        image_processor = CLIPImageProcessor()
    return image_processor.preprocess(frames, return_tensors='pt')['pixel_values']

