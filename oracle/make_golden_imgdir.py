"""Pin valley_b200.video.load_image_dir against the REFERENCE's load_video on a directory of frame images (data_util.py:282-302)
and write tests/golden/ref_imgdir.pt.

Run in the build container only (needs /root/reference; decord / cv2 / skimage are stubbed -- this branch never touches them).
The images are regenerated from seeds by `make_images` (PNG, lossless), so the fixture stores only, per file name and per
`frame_process_method`, the float32 frame the reference produced ([3,224,224]) -- independent of the directory order."""
import os
import sys
import tempfile
import types

for n in ("decord", "skimage", "skimage.transform", "cv2"):
    sys.modules.setdefault(n, types.ModuleType(n))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")

import numpy as np
import torch
from PIL import Image
from transformers import CLIPImageProcessor

SIZES = [(240, 320), (320, 240), (224, 224), (300, 500), (257, 255), (480, 360), (200, 200)]     # (H, W) of frame_00 .. frame_06
SAME = [(300, 500)] * 3       # 'resize' squares every frame to min(size of the FIRST frame): same-size frames make that order-free


def make_images(d, sizes=SIZES, seed0=500):
    """frame_XX.png: smooth ramps + noise of a per-file seed (PNG is lossless: the test regenerates the same pixels)"""
    names = []
    for i, (h, w) in enumerate(sizes):
        rs = np.random.RandomState(seed0 + i)
        yy, xx = np.mgrid[0:h, 0:w]
        b = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) % 256)], -1)
        b = np.clip(b + rs.randint(-40, 41, b.shape), 0, 255).astype(np.uint8)
        name = f"frame_{i:02d}.png"
        Image.fromarray(b).save(os.path.join(d, name))
        names.append(name)
    return names


def sha(t):
    import hashlib
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    from valley.util import data_util as ref                        # the reference
    from valley_b200 import video as ours
    proc = CLIPImageProcessor()
    gold = {"sizes": SIZES, "same": SAME, "frames": {}}
    for method, sizes, seed0 in (("centercrop", SIZES, 500), ("resize", SAME, 700)):
        with tempfile.TemporaryDirectory() as d:
            names = make_images(d, sizes, seed0)
            for n_fixed in sorted({3, len(names)}):
                r = ref.load_video(d, image_processer=proc, frame_mode="fixed", fixed_frame_number=n_fixed, frame_process_method=method)
                r = r.permute(1, 0, 2, 3).contiguous()              # frames first
                o = ours.load_image_dir(d, proc, "fixed", n_fixed, method)
                assert r.shape == o.shape and torch.equal(r, o), (method, n_fixed, float((r - o).abs().max()))
                assert torch.equal(ours.load_image_dir(d, None, "fixed", n_fixed, method), r)      # default processor == CLIP's
                picked = [p.name for p in ours.select_image_dir_frames(d, "fixed", n_fixed)]
                for k, name in enumerate(picked):
                    gold["frames"][(method, name)] = {"sha256": sha(r[k]), "sample": r[k].flatten()[::997].clone()}
                print(f"{method:10s} fixed={n_fixed}: reference == load_image_dir (bit-exact), picked {picked}")
            for mode, msg in (("fps", "Input folder is not support this frame mode"), ("other", 'Frame mode is only support "fps" or "fixed"')):
                for fn in (lambda: ref.load_video(d, image_processer=proc, frame_mode=mode), lambda: ours.load_image_dir(d, proc, mode)):
                    try:
                        fn()
                        raise AssertionError("no error for frame_mode=" + mode)
                    except ValueError as e:
                        assert str(e) == msg, str(e)
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_imgdir.pt")
    torch.save(gold, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
