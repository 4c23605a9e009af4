#!/usr/bin/env python
"""Writes N synthetic 200x200 JPEGs + input.csv: ``python tests/tools/make_synth_dataset.py <dir> <N>``."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main(out_dir, n):
    from PIL import Image

    from oracle.preprocess import synth_image

    os.makedirs(out_dir, exist_ok=True)
    names = []
    for i in range(n):
        rng = np.random.default_rng(20221000 + i)
        img = synth_image(i, 200, 200)
        name = f"{i:05d}.jpg"
        Image.fromarray(img).save(os.path.join(out_dir, name), quality=int(rng.integers(65, 100)), subsampling=2)
        names.append(name)
    with open(os.path.join(out_dir, "input.csv"), "w") as f:
        f.write("filename\n" + "\n".join(names) + "\n")
    print(f"wrote {n} images to {out_dir}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
