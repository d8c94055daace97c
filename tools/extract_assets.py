#!/usr/bin/env python3
"""Build the geometry fixtures under assets/ from the reference's model files.

Run once in the development container (needs /root/reference); the outputs are
committed so nothing at test/bench time reads /root/reference.

  assets/cornellbox.json   the nine Cornell-box meshes (reference models/cornellbox/*.obj),
                           vertices and 0-based triangle indices, in file order
  assets/bunny_x1500.npz   reference models/bunny/bunny.obj moved into the box:
                           x' = 278 + (x + 0.0167) * 1500, y' = (y - 0.0333) * 1500,
                           z' = 280 + (z + 0.0015) * 1500   (SURVEY.md F7 / §8d C4 —
                           the reference has no mesh transform, Triangle.cpp:32-75)
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("TPT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = os.path.join(os.path.dirname(HERE), "assets")


def read_obj(path):
    verts, faces = [], []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = [int(t.split("/")[0]) for t in tok[1:]]
                assert len(idx) == 3, "only triangles in the shipped assets"
                faces.append([i - 1 if i > 0 else len(verts) + i for i in idx])
    return verts, faces


def main():
    box = {}
    d = os.path.join(REF, "models", "cornellbox")
    for name in sorted(os.listdir(d)):
        if name.endswith(".obj"):
            v, f = read_obj(os.path.join(d, name))
            box[name[:-4]] = {"vertices": v, "faces": f}
    with open(os.path.join(ASSETS, "cornellbox.json"), "w") as out:
        json.dump(box, out, indent=1)
    v, f = read_obj(os.path.join(REF, "models", "bunny", "bunny.obj"))
    v = np.asarray(v, dtype=np.float64)
    t = np.empty_like(v)
    t[:, 0] = 278.0 + (v[:, 0] + 0.0167) * 1500.0
    t[:, 1] = (v[:, 1] - 0.0333) * 1500.0
    t[:, 2] = 280.0 + (v[:, 2] + 0.0015) * 1500.0
    np.savez_compressed(os.path.join(ASSETS, "bunny_x1500.npz"),
                        vertices=t.astype(np.float32), faces=np.asarray(f, dtype=np.int32))
    print("wrote", len(box), "cornell meshes and the bunny:", t.shape, len(f), "faces")
    print("bunny bounds", t.min(0), t.max(0))


if __name__ == "__main__":
    sys.exit(main())
