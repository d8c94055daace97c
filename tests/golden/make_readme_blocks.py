#!/usr/bin/env python3
"""Reduce the reference's README images (the only known-answer outputs it ships,
reference images/*.jpg, README.md:32-38) to 16x16-pixel block means of the 8-bit
tonemapped values: tests/golden/readme_blocks.npz, 49x49x3 float32 per image.
Run in the development container (needs /root/reference and PIL)."""
import os

import numpy as np
from PIL import Image

REF = os.environ.get("TPT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {"Standard": "standard", "SmoothDieletric": "smooth", "SilverBackground": "silver",
         "RefractiveBall": "refractive", "Occlusion": "occlusion"}

out = {}
for ref_name, scene in NAMES.items():
    for tag in ("PT-64", "BDPT-16"):
        img = np.asarray(Image.open(os.path.join(REF, "images", "Cornell-%s-%s.jpg" % (ref_name, tag))).convert("RGB"),
                         dtype=np.float32)
        assert img.shape == (784, 784, 3)
        out["%s_%s" % (scene, tag.split("-")[0].lower())] = img.reshape(49, 16, 49, 16, 3).mean((1, 3))
np.savez_compressed(os.path.join(HERE, "readme_blocks.npz"), **out)
print("wrote", sorted(out))
