// mmz_render.cuh - batched top-down RGB rasteriser: the stand-in for MazeEnv.render(mode="rgb_array")
// (reference maze_env.py:389-420, which reads pixels back from MuJoCo's OpenGL off-screen context, one
// environment per call). There is no OpenGL here and no pixel-parity target: the image is an orthographic view from
// +z of what the reference's scene contains - floor, maze boxes, platforms / chasms, goal sites, the agent's geoms,
// movable blocks and object balls - drawn for `count` environments in one launch, one block per environment.
//
// Per block: thread 0 walks the kinematic tree of its environment from qpos (body poses), the first ngeom threads
// place the geoms, then every thread shades pixels: the static layer comes from the maze grid, the moving layer from
// a ray cast down the z axis against every moving geom (sphere / capsule silhouettes in closed form, boxes by the
// slab test in the box frame); the highest hit wins.
#pragma once
#include "mmz_layout.h"
#include "mmz_math.cuh"

namespace mmz {

struct RenderArgs {
  const mmz_model* model;
  const float* state;  // [nstate][npad]
  uint8_t* rgb;        // [count][height][width][3], row 0 = largest y
  int npad, first_env, count, width, height;
  float x0, y0, x1, y1;  // world window
};

struct RenderGeom {
  float pos[3], R[9];
};

__device__ __forceinline__ void shade(float* c, float r, float g, float b, float k) {
  c[0] = r * k; c[1] = g * k; c[2] = b * k;
}

__global__ void __launch_bounds__(256) maze_render_kernel(const __grid_constant__ RenderArgs A) {
  const mmz_model* __restrict__ m = A.model;
  __shared__ float xpos[MMZ_MAXBODY][3], xquat[MMZ_MAXBODY][4];
  __shared__ RenderGeom geom[MMZ_MAXGEOM];
  const int env = A.first_env + blockIdx.x;
  auto Q = [&](int i) { return A.state[(size_t)i * A.npad + env]; };

  if (threadIdx.x == 0) {  // body poses, parents first (the blob orders bodies that way)
    for (int b = 0; b < m->nbody; b++) {
      const int p = m->body_parent[b];
      float pos[3], quat[4], R[9];
      if (p < 0) {
        for (int k = 0; k < 3; k++) pos[k] = m->body_pos[b][k];
        for (int k = 0; k < 4; k++) quat[k] = m->body_quat[b][k];
      } else {
        quat2mat(R, xquat[p]);
        mat_vec(pos, R, m->body_pos[b]);
        for (int k = 0; k < 3; k++) pos[k] += xpos[p][k];
        quat_mul(quat, xquat[p], m->body_quat[b]);
      }
      for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
        const int qa = m->jnt_qadr[j], type = m->jnt_type[j];
        if (type == MMZ_JNT_FREE) {
          for (int k = 0; k < 3; k++) pos[k] = Q(qa + k);
          for (int k = 0; k < 4; k++) quat[k] = Q(qa + 3 + k);
          quat_norm(quat);
          continue;
        }
        quat2mat(R, quat);
        float anchor[3], axis[3];
        mat_vec(anchor, R, m->jnt_pos[j]);
        for (int k = 0; k < 3; k++) anchor[k] += pos[k];
        mat_vec(axis, R, m->jnt_axis[j]);
        const float dq = Q(qa) - m->qpos0[qa];
        if (type == MMZ_JNT_SLIDE) {
          for (int k = 0; k < 3; k++) pos[k] += axis[k] * dq;
        } else {  // hinge: rotate about the anchor
          float qr[4], q2[4], off[3];
          axisangle2quat(qr, m->jnt_axis[j], dq);
          quat_mul(q2, quat, qr);
          for (int k = 0; k < 4; k++) quat[k] = q2[k];
          quat2mat(R, quat);
          mat_vec(off, R, m->jnt_pos[j]);
          for (int k = 0; k < 3; k++) pos[k] = anchor[k] - off[k];
        }
      }
      quat_norm(quat);
      for (int k = 0; k < 3; k++) xpos[b][k] = pos[k];
      for (int k = 0; k < 4; k++) xquat[b][k] = quat[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < m->ngeom) {
    const int g = threadIdx.x, b = m->geom_body[g];
    float R[9], q[4];
    quat2mat(R, xquat[b]);
    mat_vec(geom[g].pos, R, m->geom_pos[g]);
    for (int k = 0; k < 3; k++) geom[g].pos[k] += xpos[b][k];
    quat_mul(q, xquat[b], m->geom_quat[g]);
    quat2mat(geom[g].R, q);
  }
  __syncthreads();

  const float s = m->cell_size;
  const int agent_root = m->body_root[0];
  uint8_t* out = A.rgb + (size_t)blockIdx.x * A.width * A.height * 3;
  for (int p = threadIdx.x; p < A.width * A.height; p += blockDim.x) {
    const int py = p / A.width, px = p - py * A.width;
    const float x = A.x0 + (px + 0.5f) * (A.x1 - A.x0) / A.width;
    const float y = A.y1 - (py + 0.5f) * (A.y1 - A.y0) / A.height;
    float c[3], top;
    // ---- static layer: floor / platform checker, chasms, maze boxes (maze_env.py:116-150)
    const int j = (int)floorf((x + m->origin[0]) / s + 0.5f), i = (int)floorf((y + m->origin[1]) / s + 0.5f);
    const bool inside = i >= 0 && i < m->grid_h && j >= 0 && j < m->grid_w;
    const int cell = inside ? m->grid[i * m->grid_w + j] : 0;
    const bool dark = (((int)floorf(x) + (int)floorf(y)) & 1) != 0;
    if (cell & MMZ_CELL_WALL) {
      top = m->wall_z + m->wall_half[2];
      shade(c, 0.4f, 0.4f, 0.4f, 1.f);  // rgba of the maze blocks, maze_env.py:133,148
    } else if (m->elevated && !(cell & MMZ_CELL_PLATFORM)) {
      top = m->floor_z;
      shade(c, 0.05f, 0.05f, 0.08f, 1.f);  // chasm: the floor far below the platforms
    } else {
      top = m->elevated ? m->plat_z + m->wall_half[2] : m->floor_z;
      if (m->elevated) shade(c, 0.9f, 0.9f, 0.9f, dark ? 0.85f : 1.f);  // platform, maze_env.py:131-137
      else shade(c, 0.8f, 0.9f, 0.8f, dark ? 0.85f : 1.f);  // floor rgba of ant.xml:20 / point.xml:17
    }
    // ---- goal sites: flat discs on the ground (maze_env.py:199-210), drawn under the moving geoms
    for (int g = 0; g < m->ngoal; g++) {
      const float dx = x - m->goal_pos[g][0], dy = y - m->goal_pos[g][1], r = 0.1f * s;  // site size, maze_env.py:203
      if (dx * dx + dy * dy <= r * r && !(cell & MMZ_CELL_WALL)) shade(c, 0.9f, 0.15f, 0.15f, 1.f);
    }
    // ---- moving geoms: highest hit of the ray (x, y, +inf) -> -z
    for (int g = 0; g < m->ngeom; g++) {
      const RenderGeom& G = geom[g];
      const int type = m->geom_type[g];
      float hit = -1e30f;
      if (type == MMZ_GEOM_SPHERE || type == MMZ_GEOM_CAPSULE) {
        const float r = m->geom_size[g][0];
        float cx = G.pos[0], cy = G.pos[1], cz = G.pos[2];
        if (type == MMZ_GEOM_CAPSULE) {  // nearest point of the projected axis segment
          const float h = m->geom_size[g][1], ax = G.R[2], ay = G.R[5], az = G.R[8];
          const float den = ax * ax + ay * ay;
          float t = den > 1e-12f ? ((x - cx) * ax + (y - cy) * ay) / den : (az > 0.f ? h : -h);
          t = fminf(h, fmaxf(-h, t));
          cx += t * ax; cy += t * ay; cz += t * az;
        }
        const float d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
        if (d2 <= r * r) hit = cz + sqrtf(r * r - d2);
      } else if (type == MMZ_GEOM_BOX) {  // slab test in the box frame
        const float rel[3] = {x - G.pos[0], y - G.pos[1], 1e3f - G.pos[2]}, down[3] = {0.f, 0.f, -1.f};
        float o[3], d[3], t0 = -1e30f, t1 = 1e30f;
        matT_vec(o, G.R, rel);
        matT_vec(d, G.R, down);
        for (int k = 0; k < 3; k++) {
          const float half = m->geom_size[g][k];
          if (fabsf(d[k]) < 1e-9f) {
            if (fabsf(o[k]) > half) t0 = 1e30f;
          } else {
            const float a = (-half - o[k]) / d[k], b2 = (half - o[k]) / d[k];
            t0 = fmaxf(t0, fminf(a, b2)); t1 = fminf(t1, fmaxf(a, b2));
          }
        }
        if (t0 <= t1) hit = 1e3f - t0;
      }
      if (hit > top) {
        top = hit;
        const bool agent = m->body_root[m->geom_body[g]] == agent_root;
        const float k = 0.8f + 0.2f * fminf(1.f, fmaxf(0.f, hit / (2.f * m->wall_half[2] + 1e-6f)));
        if (agent) shade(c, 0.8f, 0.6f, 0.4f, k);                         // ant.xml / point.xml body colour
        else if (type == MMZ_GEOM_BOX) shade(c, 0.9f, 0.1f, 0.1f, k);     // movable blocks, maze_env.py:600
        else shade(c, 0.1f, 0.1f, 0.7f, k);                               // object balls, maze_env.py:500
      }
    }
    for (int k = 0; k < 3; k++) out[(size_t)p * 3 + k] = (uint8_t)fminf(255.f, c[k] * 255.f + 0.5f);
  }
}

}  // namespace mmz
