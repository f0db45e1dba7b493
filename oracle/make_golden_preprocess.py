"""Pin oracle/preprocess_oracle.py against the REFERENCE's own preprocessing classes and write tests/golden/ref_preprocess.pt.

Run in the build container only (needs /root/reference; cv2 / skimage / decord are stubbed -- the PIL branch is the one
``load_video`` takes).  Pipeline = data_util.py:271-281 verbatim.  The fixture stores, per geometry, the SHA-256 of the
reference's uint8 crop and float32 output plus a strided sample of both (the clips themselves are regenerated from the seed).
"""
import hashlib
import os
import sys
import types

for n in ("decord", "skimage", "skimage.transform", "cv2"):
    sys.modules.setdefault(n, types.ModuleType(n))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")

import numpy as np
import torch
from torchvision import transforms

from oracle import preprocess_oracle as P
from valley.data import video_transform as vt            # noqa: E402  (the reference)

GEOMETRIES = [(360, 640), (640, 360), (256, 340), (300, 256), (200, 150), (224, 224), (720, 1280), (255, 257), (481, 853)]


def make_clip(h, w, seed, T=2):
    """frame 0: uniform noise (worst case for rounding); frame 1: smooth ramps + noise."""
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    b = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) % 256)], -1)
    b = np.clip(b + rs.randint(-3, 4, b.shape), 0, 255).astype(np.uint8)
    return np.stack([a, b][:T])


def reference_pipeline(frames_u8):
    video = torch.from_numpy(frames_u8).permute(3, 0, 1, 2)         # 3 x T x H x W, as load_video builds it
    pre = transforms.Compose([vt.TensorToNumpy(), vt.Resize(256), vt.CenterCrop(224)])
    pil = pre(video)
    u8 = np.stack([np.array(im) for im in pil])
    full = transforms.Compose([vt.TensorToNumpy(), vt.Resize(256), vt.CenterCrop(224), vt.ClipToTensor(channel_nb=3),
                               vt.Normalize(mean=list(P.CLIP_MEAN), std=list(P.CLIP_STD))])
    return u8, full(video).permute(1, 0, 2, 3).contiguous().numpy()  # frames first


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    gold = {}
    for i, (h, w) in enumerate(GEOMETRIES):
        clip = make_clip(h, w, 100 + i)
        r_u8, r_f = reference_pipeline(clip)
        o_u8 = P.preprocess_frames(clip, return_uint8=True)
        o_f = P.preprocess_frames(clip)
        assert r_f.dtype == np.float32 and r_f.shape == (2, 3, 224, 224)
        assert np.array_equal(o_u8, r_u8), (h, w, int(np.abs(o_u8.astype(int) - r_u8).max()))
        assert np.array_equal(o_f.view(np.uint32), r_f.view(np.uint32)), (h, w)
        gold[(h, w)] = dict(seed=100 + i, sha_u8=sha(r_u8), sha_f32=sha(r_f), sub_f32=torch.from_numpy(r_f[:, :, ::9, ::7].copy()))
        print(f"  {h}x{w}: oracle == reference bit-for-bit (uint8 crop and float32 output)")
    torch.save(gold, os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_preprocess.pt"))
    print("wrote tests/golden/ref_preprocess.pt")


if __name__ == "__main__":
    main()
