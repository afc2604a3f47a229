"""Generates the committed fixtures under tests/golden/ from the reference's own outputs.

Run HERE (the container that has /root/reference); the GPU box only sees the results.
  cornell_ref_75.npy — /root/reference/cornel_box.png (the reference's shipped render of the
      deterministic Cornell-box scene, 600x600 RGBA, ~200 spp), RGB, 8x8 box-downsampled to
      75x75 float32 (8-bit units). Noise drops 8x, which makes region-wise comparison with a
      render of ours meaningful (two independent 200-spp renders differ by ~14 dB per pixel).
  final_ref_100.npy — /root/reference/image.png (final scene, 800x800) 8x8-downsampled. The
      scene's geometry is drawn from an unseeded RNG, so this is only a coarse (layout,
      brightness) reference.
"""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def down(path, f):
    a = np.asarray(Image.open(path).convert("RGB"), dtype=np.float64)
    h, w, _ = a.shape
    return a.reshape(h // f, f, w // f, f, 3).mean(axis=(1, 3)).astype(np.float32)


if __name__ == "__main__":
    np.save(os.path.join(HERE, "cornell_ref_75.npy"), down(os.path.join(REF, "cornel_box.png"), 8))
    np.save(os.path.join(HERE, "final_ref_100.npy"), down(os.path.join(REF, "image.png"), 8))
    print("written")
