"""Ray-visible geometry of a compiled model: what the lidar (<rangefinder>, row S2) and the
camera (row C3) can hit.

[upstream] mj_ray semantics restated in SURVEY.md Appendix B: every geom is a candidate regardless
of group or contype, geoms whose rgba (or material) alpha is 0 are skipped, meshes are intersected
as TRIANGLE meshes (not hulls).  A collision-class geom that is an exact copy of a visual geom
(same body, mesh and pose -- most of `stretch.xml`'s collision meshes) produces identical hits and
is dropped here so that it is not traversed twice.

Arrays added to the model (consumed by csrc/rays.cu and by the oracle):
  raygeom_id[n]            geom ids that rays test
  geom_shade[ngeom, 8]     rgb, alpha, specular, shininess, emission, reflectance
  geom_tex[ngeom, 4]       2-D texture of the geom's material: texture index (-1 = none), texrepeat x / y, mode (0 = planar x-y
                           projection scaled to the geom, 1 = texuniform, 2 = the mesh's UV coordinates)
  rmesh_uv[*, 6] f32       per triangle of rmesh_face: the three UV pairs (zeros for meshes without UVs)
  tex_adr/tex_w/tex_h[ntex], tex_rgb u8 [sum h*w*3]   texture images (row 0 = top row of the file)
  rmesh_vertadr/faceadr/facenum[nmesh], rmesh_vert[*,3] f32, rmesh_face[*,3] i32 (mesh frame)
Acceleration structures (BVHs) are built at load time by each consumer.
"""
from __future__ import annotations

import numpy as np

from .mjcf import _floats


def build_ray_geometry(m, verbose: bool = False) -> None:
    A = m.arrays
    sc = m.scene
    ngeom = len(A["geom_type"])
    # per-geom shading parameters (geom rgba overrides the material colour when given explicitly)
    shade = np.zeros((ngeom, 8))
    flat_geoms = [g for b in sc.bodies for g in b.geoms]
    for gi, g in enumerate(flat_geoms):
        rgba = A["geom_rgba"][gi]
        spec, shin, emis, refl = 0.5, 0.5, 0.0, 0.0
        mat = g.get("material")
        if mat:
            mt = sc.materials[mat]
            spec, shin, emis, refl = float(mt["specular"]), float(mt["shininess"]), float(mt["emission"]), float(mt["reflectance"])
        shade[gi] = [rgba[0], rgba[1], rgba[2], rgba[3], spec, shin, emis, refl]
    A["geom_shade"] = shade
    # 2-D textures (`<texture type="2d" file=...>` behind a material, models/scene.xml:15-16): sampled on boxes and planes through
    # the planar x-y projection of the geom frame [upstream: glTexGen OBJECT_LINEAR]; meshes would need their UV sets and keep
    # the material colour
    gtex = np.zeros((ngeom, 4)); gtex[:, 0] = -1
    tex_ids, tex_imgs = {}, []
    for gi, g in enumerate(flat_geoms):
        mat = g.get("material")
        if not mat or A["geom_type"][gi] not in (0, 6, 7):
            continue
        if A["geom_type"][gi] == 7 and m.mesh_assets[int(A["geom_dataid"][gi])].face_tc is None:
            continue   # a textured mesh without UV coordinates keeps the material colour
        mt = sc.materials[mat]
        tname = mt.get("texture")
        tx = sc.textures.get(tname) if tname else None
        if not tx or tx.get("type", "cube") != "2d" or "path" not in tx:
            continue
        if tname not in tex_ids:
            try:
                from PIL import Image
                img = np.asarray(Image.open(tx["path"]).convert("RGB"), dtype=np.uint8)
            except Exception:
                continue
            tex_ids[tname] = len(tex_imgs); tex_imgs.append(img)
        rep = _floats(mt.get("texrepeat", "1 1"), 2)
        mode = 2.0 if A["geom_type"][gi] == 7 else (1.0 if mt.get("texuniform", "false") == "true" else 0.0)   # 2 = the mesh's UV set
        gtex[gi] = [tex_ids[tname], rep[0], rep[1], mode]
    A["geom_tex"] = gtex
    adr = np.cumsum([0] + [im.size for im in tex_imgs])[:-1] if tex_imgs else np.zeros(0)
    A["tex_adr"] = np.asarray(adr, dtype=np.int32)
    A["tex_w"] = np.asarray([im.shape[1] for im in tex_imgs], dtype=np.int32)
    A["tex_h"] = np.asarray([im.shape[0] for im in tex_imgs], dtype=np.int32)
    A["tex_rgb"] = np.concatenate([im.reshape(-1) for im in tex_imgs]) if tex_imgs else np.zeros(0, np.uint8)
    # candidate list with duplicate suppression
    keep = []
    seen = {}
    order = sorted(range(ngeom), key=lambda g: (A["geom_group"][g] > 2, g))  # visual copies win over collision copies
    for g in order:
        if shade[g, 3] == 0:
            continue
        if A["geom_type"][g] == 7:
            key = (int(A["geom_bodyid"][g]), int(A["geom_dataid"][g]), tuple(np.round(A["geom_pos"][g], 9)),
                   tuple(np.round(A["geom_quat"][g], 9)))
            if key in seen:
                continue
            seen[key] = g
        keep.append(g)
    keep.sort()
    A["raygeom_id"] = np.asarray(keep, dtype=np.int32)
    # triangle soups of the meshes referenced by kept geoms
    nmesh = len(m.mesh_assets)
    used = sorted({int(A["geom_dataid"][g]) for g in keep if A["geom_type"][g] == 7})
    vertadr = -np.ones(nmesh, np.int32); faceadr = -np.ones(nmesh, np.int32); facenum = np.zeros(nmesh, np.int32)
    verts, faces, uvs = [], [], []
    nv = nf = 0
    for mid in used:
        ma = m.mesh_assets[mid]
        vertadr[mid], faceadr[mid], facenum[mid] = nv, nf, len(ma.faces)
        verts.append(ma.verts.astype(np.float32)); faces.append(ma.faces.astype(np.int32))
        if ma.face_tc is not None and ma.texcoord is not None:
            uvs.append(np.asarray(ma.texcoord)[np.asarray(ma.face_tc)].reshape(-1, 6).astype(np.float32))
        else:
            uvs.append(np.zeros((len(ma.faces), 6), np.float32))
        nv += len(ma.verts); nf += len(ma.faces)
    A["rmesh_vertadr"], A["rmesh_faceadr"], A["rmesh_facenum"] = vertadr, faceadr, facenum
    A["rmesh_vert"] = np.concatenate(verts) if verts else np.zeros((0, 3), np.float32)
    A["rmesh_face"] = np.concatenate(faces) if faces else np.zeros((0, 3), np.int32)
    A["rmesh_uv"] = np.concatenate(uvs) if uvs else np.zeros((0, 6), np.float32)
    if verbose:
        print(f"ray geometry: {len(keep)} of {ngeom} geoms, {len(used)} meshes, {nf} triangles, {nv} vertices")
