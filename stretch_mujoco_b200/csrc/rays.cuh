// Device-side ray-geometry description (lidar + camera kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct RayModel {
  int present;
  int nraygeom, nmesh, ngeom, nbody, ncam, nsite, nlight, nsky, headlight_active;
  float extent, znear, zfar;
  const int *rg_geom, *rg_type, *rg_body, *rg_mesh, *rg_group;   // ray-visible geoms (alpha>0, deduplicated)
  const float *rg_pos, *rg_quat, *rg_size, *rg_rbound, *rg_shade; // shade: rgb, specular, shininess, emission (6)
  const int *rmesh_vertadr, *rmesh_faceadr, *rmesh_facenum, *rmesh_bvhadr;
  const float4* tri;        // 3 float4 per triangle: v0, e1, e2 (mesh frame), BVH leaf order
  const float4* bvh;        // 4 float4 per internal node: children boxes + (refL, refR, cntL, cntR), see rays.cu
  const float4* rg_rec;     // 4 float4 per ray-geom: (type, root ref, root count, body|group<<16), (rbound, size), lo, hi
  const int *cam_bodyid, *site_bodyid;
  const float *cam_pos, *cam_quat, *cam_fovy, *site_pos, *site_quat;
  const int *range_site;    // site id per rangefinder sensor
  const int *range_adr;     // sensordata address per rangefinder
  const float* range_cutoff;
  int nrange;
  float headlight[9], sky[6];
  // raster camera path: work chunks of <= RCHUNK consecutive triangles (BVH leaf order, spatially compact) of the
  // camera-visible mesh geoms: (ray-geom index, first triangle, count, 0) and the chunk's bounds in the mesh frame
  int nchunk;
  const int4* rchunk;
  const float4* rchunk_box;   // 2 per chunk: lo, hi
  // 2-D textures of box / plane geoms (planar x-y projection of the geom frame): per ray-geom (texture index or -1, texrepeat x, y,
  // texuniform), per texture (first byte, width, height), RGB8 texels (row 0 = top of the image)
  const float4* rg_tex;
  const int4* tex_info;
  const unsigned char* tex_rgb;
  const float2* tri_uv;     // 3 per triangle of `tri` (uv0, uv1 - uv0, uv2 - uv0) or nullptr: UV sets of textured meshes
  int ntex;
  const int *light_bodyid, *light_directional;
  const float *light_pos, *light_dir, *light_ambient, *light_diffuse, *light_specular;
};
