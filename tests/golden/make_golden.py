#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libtptref.so).

Run in the development container (where /root/reference exists and oracle/build_ref.sh
has been run).  The outputs are committed; tests compare the restatement (oracle/), the
host API and the CUDA path against them, so nothing at test time needs /root/reference.

  kat.npz           RNG streams, helper / material known answers, pixel rays
  rays_<scene>.npz  ray batches P (primary, strided), S (first BDPT bounce), R (random),
                    A (adversarial) with the reference's prim id / t / hit point / normal
  render.npz        64x64 reference renders (all scenes x 3 modes) as float images
  bdpt_<scene>.npz  subpaths + every strategy weight of single BDPT samples
  flat_<scene>.npz  the reference's own trees, flattened (byte images of the TptSceneDesc arrays)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

SCENES = ["standard", "smooth", "silver", "refractive", "occlusion", "bunny"]
W = H = 784


def xorshift_stream(seed, n):
    out = np.empty(n, np.float64)
    s = np.uint32(seed)
    for i in range(n):
        s ^= np.uint32((int(s) << 13) & 0xFFFFFFFF)
        s ^= np.uint32(int(s) >> 17)
        s ^= np.uint32((int(s) << 15) & 0xFFFFFFFF)
        out[i] = float(np.float32(float(int(s)) / 4294967295.0))
    return out


def random_rays(n, seed0=0xC0FFEE):
    """Batch R of SURVEY.md 8(d): origin uniform in the box, normalised direction, cull = k mod 3."""
    rs = np.random.RandomState(seed0 & 0x7FFFFFFF)
    org = (rs.rand(n, 3) * np.array([556.0, 548.8, 559.2])).astype(np.float32)
    d = (rs.rand(n, 3) * 2 - 1).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6).astype(np.float32)
    cull = (np.arange(n) % 3).astype(np.uint8)
    return org, d.astype(np.float32), cull


def adversarial_rays():
    """Batch A: zero direction components (+-0), origins on wall planes, quad diagonals, shared
    vertices, rays starting on a surface, sphere tangents."""
    o, d, c = [], [], []

    def add(org, dr, cull):
        o.append(org); d.append(dr); c.append(cull)

    nz = np.float32(-0.0)
    for cull in (0, 1, 2):
        add([0, 100, -10], [0.0, 0, 1], cull)            # along the x=0 wall plane (0*inf = NaN slab)
        add([0, 100, -10], [nz, 0, 1], cull)
        add([556, 100, -10], [0.0, 0, 1], cull)
        add([278, 0, -10], [0, 0.0, 1], cull)            # along the floor plane
        add([278, 0, -10], [0, nz, 1], cull)
        add([278, 548.8, -10], [0, nz, 1], cull)
        add([278, 274.4, -800], [0, 0, 1], cull)         # axis-aligned through the box centre
        add([278, 274.4, 280], [1, 0, 0], cull)
        add([278, 274.4, 280], [-1, 0, 0], cull)
        add([278, 274.4, 280], [0, 1, 0], cull)
        add([278, 274.4, 280], [0, -1, 0], cull)
        add([278, 500, 279.5], [0, 1, 0], cull)          # up through the light's shared diagonal region
        add([213 + 65, 400, 227 + 52.5], [0, 1, 0], cull)
        add([278, 10, 279.6], [0, -1, 0], cull)          # down onto the floor's quad diagonal
        add([556 / 2, 100, 559.2 / 2], [0, -1, 0], cull)
        add([0, 0, 0], [1, 1, 1], cull)                  # from a shared corner vertex
        add([100, 100, 100], [-1, -1, -1], cull)         # into a corner vertex
        add([278, 0, 280], [0, 1, 0], cull)              # starts on the floor (t = 0 candidates)
        add([278, 0, 280], [0.3, 1, 0.2], cull)
        add([0, 274, 280], [1, 0.1, 0.1], cull)          # starts on the left wall
        add([278, 278, -800], [0, 0, -1], cull)          # away from everything
        add([228, 278, -100], [0, 0, 1], cull)           # tangent to the glass ball (x = 278 - 50)
        add([328, 278, -100], [0, 0, 1], cull)
        add([278, 278, 200], [0, 0, 1], cull)            # from the ball's centre
        add([278, 278, 150], [0, 0, 1], cull)            # from the ball's surface
        add([300, 0, 300], [0, 1, 0], cull)              # light occluder duplicate faces (exact t ties)
        add([200, 0, 250], [0, 1, 0], cull)
        add([350, 0, 200], [0, 1, 0], cull)
    org = np.array(o, np.float32)
    dr = np.array(d, np.float32)
    n = np.linalg.norm(dr.astype(np.float64), axis=1, keepdims=True)
    dr = np.where(dr == 0, dr, (dr / n).astype(np.float32)).astype(np.float32)   # keep the signed zeros
    return org, dr, np.array(c, np.uint8)


def primary_rays(chk, w, h, stride):
    scale = chk.calculate_scale(40.0)
    idx = np.arange(0, w * h, stride)
    dirs = np.stack([chk.pixel_ray(int(i % w), int(i // w), w, h, scale) for i in idx]).astype(np.float32)
    org = np.tile(np.array([278, 278, -800], np.float32), (len(idx), 1))
    return idx, org, dirs


def main():
    assert B.have_ref(), "build oracle/_ref first (oracle/build_ref.sh)"
    os.makedirs(HERE, exist_ok=True)

    # ---- known answers -------------------------------------------------------------
    ref, _ = B.ref_scene("refractive", W, H)
    kat = {}
    for seed in (1, 2, 614656, 400275, 12345):
        st, fl = ref.rng(seed, 64)
        kat["rng_state_%d" % seed] = st
        kat["rng_float_%d" % seed] = fl
    n = np.array([0, 1, 0], np.float32)
    wo = np.array([0.3, 0.8, -0.2], np.float32); wo /= np.float32(np.sqrt((wo * wo).sum()))
    wi = np.array([-0.5, 0.6, 0.1], np.float32); wi /= np.float32(np.sqrt((wi * wi).sum()))
    wt = np.array([-0.2, -0.9, 0.1], np.float32); wt /= np.float32(np.sqrt((wt * wt).sum()))
    kat["vec_n"], kat["vec_wo"], kat["vec_wi"], kat["vec_wt"] = n, wo, wi, wt
    r1, r2, r3 = ref.helpers(wo, n, 1.5)
    kat["reflect_wo_n"], kat["refract_wo_n"], kat["perp_wo"] = r1, r2, r3
    kat["refract_wt_n"] = ref.helpers(wt, n, 1.5)[1]
    kat["scale40"] = np.float32(ref.calculate_scale(40.0))
    px = np.array([[392, 392], [0, 0], [100, 700], [392, 120], [600, 600], [434, 510], [783, 783]], np.int32)
    kat["pixels"] = px
    kat["pixel_rays"] = np.stack([ref.pixel_ray(int(x), int(y), W, H, float(kat["scale40"])) for x, y in px])
    # random material inputs: every material of the refractive scene (white, red, green, light, glass) + silver
    rs = np.random.RandomState(7)
    m = 4096

    def unit(v):
        v = v.astype(np.float32)
        return (v / np.linalg.norm(v.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
    mwo, mwi, mn = unit(rs.randn(m, 3)), unit(rs.randn(m, 3)), unit(rs.randn(m, 3))
    mn[: m // 4] = np.array([0, 1, 0], np.float32)          # axis-aligned normals as in the box
    mn[m // 4: m // 2] = np.array([0, 0, -1], np.float32)
    kat["mat_wo"], kat["mat_wi"], kat["mat_n"] = mwo, mwi, mn
    seeds = (np.arange(m) * 7919 + 12345).astype(np.uint32)
    kat["mat_seeds"] = seeds
    sref, _ = B.ref_scene("silver", W, H)
    for tag, chk, mat in (("white", ref, 0), ("red", ref, 1), ("light", ref, 3), ("glass", ref, 4), ("silver", sref, 0)):
        kat["eval1_" + tag] = chk.mat_eval(mat, mwo, mwi, mn, True)
        kat["eval0_" + tag] = chk.mat_eval(mat, mwo, mwi, mn, False)
        kat["pdf_" + tag] = chk.mat_pdf(mat, mwo, mn, mwi)
        kat["fresnel_" + tag] = chk.mat_fresnel(mat, mwi, mn)
        swi, spdf, sst = chk.mat_sample(mat, mwo, mn, seeds)
        kat["sample_wi_" + tag], kat["sample_pdf_" + tag], kat["sample_state_" + tag] = swi, spdf, sst
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)

    # ---- ray batches + flattened trees ---------------------------------------------
    for name in SCENES:
        chk, desc = B.ref_scene(name, W, H)
        arrs = B.desc_arrays(desc)
        hdr = arrs.pop("header")
        np.savez_compressed(os.path.join(HERE, "flat_%s.npz" % name), width=hdr[0], height=hdr[1], fov=hdr[2],
                            eye=np.array(hdr[3], np.float32), background=np.array(hdr[4], np.float32), **arrs)
        out = {}
        idx, org, dirs = primary_rays(chk, W, H, 37)
        batches = {"P": (org, dirs, np.zeros(len(org), np.uint8))}
        # S: the first BDPT bounce of those pixels, generated by the reference itself
        so, sd, sc = [], [], []
        for i in idx[::4]:
            cam, nc, light, nl, w, st = chk.bdpt_sample(int(i), int(i) + 1)
            if nc >= 3 or (nc == 2 and cam[1]["type"] == 1):
                pass
            if cam[1]["type"] != 1:
                continue
            # re-derive the bounce ray from the vertices the reference produced (float bits preserved)
            if nc >= 3 and cam[2]["type"] == 1:
                o = cam[1]["x"]; dd = (cam[2]["x"].astype(np.float64) - o.astype(np.float64))
                dd = (dd / np.linalg.norm(dd)).astype(np.float32)
                so.append(o); sd.append(dd); sc.append(0 if float(np.dot(cam[1]["N"], dd)) > 0 else 1)
        if so:
            batches["S"] = (np.array(so, np.float32), np.array(sd, np.float32), np.array(sc, np.uint8))
        batches["R"] = random_rays(20000 if name != "bunny" else 8000)
        batches["A"] = adversarial_rays()
        for tag, (o, d, c) in batches.items():
            prim, t, coords, normal = chk.intersect(o, d, c)
            out[tag + "_org"], out[tag + "_dir"], out[tag + "_cull"] = o, d, c
            out[tag + "_prim"], out[tag + "_t"], out[tag + "_coords"], out[tag + "_normal"] = prim, t, coords, normal
        # shadow queries between random pairs of hit points
        prim, t, coords, normal = chk.intersect(*batches["R"])
        hit = coords[prim >= 0]
        k = min(len(hit) // 2, 4000)
        a, b = hit[:k], hit[k:2 * k]
        cull = (np.arange(k) % 2).astype(np.uint8)
        out["shadow_from"], out["shadow_to"], out["shadow_cull"] = a, b, cull
        out["shadow"] = chk.shadow(a, b, cull)
        np.savez_compressed(os.path.join(HERE, "rays_%s.npz" % name), **out)
        print(name, {k: len(v[0]) for k, v in batches.items()}, "hit rate R", float((prim >= 0).mean()))

    # ---- renders ----------------------------------------------------------------------
    ren = {}
    for name in SCENES:
        chk, _ = B.ref_scene(name, 64, 64)
        for mode, spp in ((0, 16), (1, 16), (2, 8)):
            img, rays, sec = chk.render(mode, spp, 8 if mode != 2 else 1, 64, 64)
            ren["%s_m%d" % (name, mode)] = img
            ren["%s_m%d_rays" % (name, mode)] = np.int64(rays)
    np.savez_compressed(os.path.join(HERE, "render.npz"), **ren)

    # ---- BDPT samples -------------------------------------------------------------------
    for name in ("standard", "refractive", "silver"):
        chk, _ = B.ref_scene(name, W, H)
        pix = np.arange(1000, W * H, 2503)[:240]
        cams, lights, ncs, nls, ws, sts = [], [], [], [], [], []
        for p in pix:
            cam, nc, light, nl, w, st = chk.bdpt_sample(int(p), int(p) + 1)
            cams.append(cam); lights.append(light); ncs.append(nc); nls.append(nl); ws.append(w); sts.append(st)
        np.savez_compressed(os.path.join(HERE, "bdpt_%s.npz" % name), pixels=pix.astype(np.int32),
                            cam=np.stack(cams), light=np.stack(lights), cam_count=np.array(ncs, np.int32),
                            light_count=np.array(nls, np.int32), weights=np.stack(ws),
                            state=np.array(sts, np.uint32))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
