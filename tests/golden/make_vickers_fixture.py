"""Builds tests/golden/vickers_512.npz from the reference's own dataset (run HERE, where /root/reference exists).

The GPU box has no /root/reference, so the trained-weights parity tests (tests/test_gpu_trained.py) read this fixture.
Per image it holds what /root/reference/train.py's VickersDataset produces BEFORE the random augmentation and the
normalisation: `cv2.imread` (train.py:145) -> `A.LongestMaxSize(512, INTER_LINEAR)` (train.py:70-71 / 121; masks with
nearest-neighbour as Albumentations does), i.e. 410 x 512 pixels for the 1024 x 1280 micrographs.  Images are stored as
JPEG (quality 92, BGR as cv2 encodes them), masks (train.py:166-173, `> 0` -> 1) as PNG, both as byte blobs.
Padding to 512 x 512 (`A.PadIfNeeded`, centred, constant 0), BGR->RGB, /255 and (x - mean) / std (train.py:108-112)
happen in the test's loader.  The train / validation split is train.py:560-565 (`random.Random(42).shuffle`, first 10 %
= validation) over the images that have a mask.

    python tests/golden/make_vickers_fixture.py          # writes tests/golden/vickers_512.npz (~8 MB)
"""
import random
import sys
from pathlib import Path

import cv2
import numpy as np

REF = Path("/root/reference/data")
OUT = Path(__file__).resolve().parent / "vickers_512.npz"
IMG_EXTS = {".jpg", ".jpeg", ".png", ".bmp", ".tif", ".tiff"}
SIZE = 512


def longest_max_size(img, size, interp):
    h, w = img.shape[:2]
    s = size / max(h, w)
    nh, nw = int(round(h * s)), int(round(w * s))
    return cv2.resize(img, (nw, nh), interpolation=interp)


def main():
    imgs = sorted(str(p) for p in (REF / "images").glob("*") if p.suffix.lower() in IMG_EXTS)
    imgs = [p for p in imgs if (REF / "masks" / (Path(p).stem + ".png")).exists()]
    r = random.Random(42)
    order = imgs[:]
    r.shuffle(order)
    n_val = max(1, int(len(order) * 0.1))
    val = set(order[:n_val])
    names, is_val, jb, mb, shapes = [], [], [], [], []
    for p in order:
        im = cv2.imread(p, cv2.IMREAD_COLOR)
        m = cv2.imread(str(REF / "masks" / (Path(p).stem + ".png")), cv2.IMREAD_UNCHANGED)
        if m.ndim == 3:
            m = m[:, :, 0]
        m = (m > 0).astype(np.uint8) * 255
        im = longest_max_size(im, SIZE, cv2.INTER_LINEAR)
        m = longest_max_size(m, SIZE, cv2.INTER_NEAREST)
        ok1, j = cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 92])
        ok2, q = cv2.imencode(".png", m, [cv2.IMWRITE_PNG_COMPRESSION, 9])
        assert ok1 and ok2
        names.append(Path(p).stem)
        is_val.append(p in val)
        jb.append(j.tobytes())
        mb.append(q.tobytes())
        shapes.append(im.shape[:2])
    j_off = np.cumsum([0] + [len(b) for b in jb]).astype(np.int64)
    m_off = np.cumsum([0] + [len(b) for b in mb]).astype(np.int64)
    np.savez(OUT, names=np.array(names), is_val=np.array(is_val), shapes=np.array(shapes, dtype=np.int32),
             jpeg=np.frombuffer(b"".join(jb), dtype=np.uint8), jpeg_off=j_off,
             mask_png=np.frombuffer(b"".join(mb), dtype=np.uint8), mask_off=m_off)
    fg = []
    for b in mb:
        mm = cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_UNCHANGED)
        fg.append((mm > 0).mean())
    print(f"{OUT}: {len(names)} images ({sum(is_val)} validation), {OUT.stat().st_size / 1e6:.1f} MB, "
          f"foreground fraction mean {np.mean(fg):.3f} (min {np.min(fg):.4f}, max {np.max(fg):.3f})")


if __name__ == "__main__":
    sys.exit(main())
