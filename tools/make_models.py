#!/usr/bin/env python3
"""Write Wavefront OBJ files from the committed geometry fixtures (assets/).

The host API keeps the reference's constructor MeshTriangle(const std::string& objPath,
Material*) (reference Triangle.hpp:56), so scenes are built from .obj paths.  This
writes them to assets/_models/ (git-ignored, rebuilt by __graft_entry__.build()):
    assets/_models/cornellbox/<name>.obj     from assets/cornellbox.json
    assets/_models/bunny/bunny_x1500.obj     from assets/bunny_x1500.npz
Floats are written with repr() so strtof() returns the float the fixture holds.
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "assets")
OUT = os.path.join(ASSETS, "_models")


def write_obj(path, verts, faces):
    with open(path, "w") as f:
        for v in verts:
            f.write("v %s %s %s\n" % tuple(repr(float(c)) for c in v))
        for tri in faces:
            f.write("f %d %d %d\n" % tuple(int(i) + 1 for i in tri))


def make_models(out=OUT):
    os.makedirs(os.path.join(out, "cornellbox"), exist_ok=True)
    os.makedirs(os.path.join(out, "bunny"), exist_ok=True)
    with open(os.path.join(ASSETS, "cornellbox.json")) as f:
        box = json.load(f)
    for name, mesh in box.items():
        write_obj(os.path.join(out, "cornellbox", name + ".obj"), mesh["vertices"], mesh["faces"])
    b = np.load(os.path.join(ASSETS, "bunny_x1500.npz"))
    # float32 -> shortest repr that round-trips through strtof
    verts = [[np.float32(c).item() for c in v] for v in b["vertices"]]
    verts = [[float(np.format_float_scientific(np.float32(c), unique=True)) for c in v] for v in verts]
    write_obj(os.path.join(out, "bunny", "bunny_x1500.obj"), verts, b["faces"])
    return out


if __name__ == "__main__":
    print(make_models())
