#!/usr/bin/env python
"""Compares a rendered image.png of scene 9 (800x800) or scene 7 (600x600) with the reference's shipped render
(8x8 box-filtered copies under tests/golden/, made by tests/golden/make_golden.py). Prints the same statistics
tests/test_gpu_parity.py::test_*_matches_the_reference_shipped_image assert on.

usage: tools/shipped_compare.py IMAGE.png
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib  # noqa: E402

R = importlib.import_module("rttnw_b200.render")  # (the package also exports a function called render)


def main():
    rgb = R.png_read_rgba8(sys.argv[1])[..., :3].astype(np.float64)
    h, w, _ = rgb.shape
    name = {800: "final_ref_100.npy", 600: "cornell_ref_75.npy"}[h]
    ref = np.load(os.path.join(ROOT, "tests", "golden", name)).astype(np.float64)
    ours = rgb.reshape(h // 8, 8, w // 8, 8, 3).mean(axis=(1, 3))
    diff = ours - ref
    print(f"{sys.argv[1]}: {w}x{h} vs tests/golden/{name}")
    print("  whole-frame mean colour difference (8-bit units, R G B):", np.round(diff.mean(axis=(0, 1)), 3))
    n = 10 if h == 800 else 15
    blocks = np.abs(diff.reshape(n, ours.shape[0] // n, n, ours.shape[1] // n, 3).mean(axis=(1, 3))).mean(axis=2)
    print(f"  |difference| of {n}x{n} regions: median {np.median(blocks):.2f}, max {blocks.max():.2f}"
          + (f"; upper half (no unseeded geometry): median {np.median(blocks[:n // 2]):.2f}, max {blocks[:n // 2].max():.2f}" if h == 800 else ""))
    mse = (diff ** 2).mean()
    print(f"  8x8-block PSNR {10 * np.log10(255.0 ** 2 / mse):.1f} dB; correlation of luminance {np.corrcoef(ours.mean(axis=2).ravel(), ref.mean(axis=2).ravel())[0, 1]:.4f}")


if __name__ == "__main__":
    main()
