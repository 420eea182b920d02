// Physics kernels of libstretchsim: one mj_step for a batch of independent envs.
//
// Replaces the reference's per-step call `mujoco.mj_step(model, data)`
// (stretch_mujoco/mujoco_server.py:378) -- stages S1..S10 of SURVEY.md §8(a) -- with a pipeline of
// three sm_100a kernels per step (csrc/api.cu:launch_physics):
//   ss_smooth_kernel  one warp per env: kinematics, CRBA, broadphase, velocity / RNE, actuation (S1, S3, S5)
//   ss_narrow_kernel  one warp per candidate geom pair of any env: narrowphase (S1c), balanced over the chip
//   ss_solve_kernel   one warp per env: constraint rows, Newton solver, sensors, implicitfast (S1m, S4, S6..S10)
// An env's working set lives in its warp's slice of shared memory inside a kernel and travels between
// kernels as one contiguous "persistent block" per env in global memory (L2 resident).  The model's small
// tables are pulled into shared memory once per CTA by a TMA bulk copy.
//
// Lane mapping: lane = dof for joint-space vectors and matrix rows, lane = body inside one tree
// level for the kinematic passes, lane = constraint row / contact for the solver's row passes.
// Cross-lane traffic goes through shared memory + __syncwarp or through shuffles for reductions.
//
// Code-size discipline (the first profile was instruction-fetch bound): every model-dependent
// stage has exactly one call site and is inlined so that DevModel fields stay constant-bank
// operands; the generic dense-algebra / row-pass helpers are __noinline__ single copies.
#include "model.cuh"
#include "batch.cuh"
#include <math_constants.h>

#define FULL 0xffffffffu
#define MINVAL 1e-15f
#define MAXVAL 1e10f
#define MINIMP 0.0001f
#define MAXIMP 0.9999f

extern __shared__ __align__(16) float smem[];
// [model pack | env slice 0 | env slice 1 | ...]
#define PKF(name) (smem + m.pk.name)
#define PKI(name) (reinterpret_cast<const int*>(smem) + m.pk.name)
#define GPI(name) (reinterpret_cast<const int*>(m.pack) + m.pk.name)   // the same table in global memory

// ----------------------------------------------------------------------------- small math
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(float* r, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void quat_mul(float* r, const float* a, const float* b) {
  float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
__device__ __forceinline__ void quat_normalize(float* q) {
  float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  float inv = 1.0f / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
__device__ __forceinline__ void quat2mat(float* R, const float* q) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ void mat_vec(float* r, const float* M, const float* v) {
  float x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2],
        z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void matT_vec(float* r, const float* M, const float* v) {
  float x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2],
        z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void mul_inert_vec(float* r, const float* i, const float* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
__device__ __forceinline__ void cross_motion(float* r, const float* vel, const float* v) {
  float a[3], b[3];
  cross3(r, vel, v);
  cross3(a, vel, v + 3);
  cross3(b, vel + 3, v);
  r[3] = a[0] + b[0]; r[4] = a[1] + b[1]; r[5] = a[2] + b[2];
}
__device__ __forceinline__ void cross_force(float* r, const float* vel, const float* f) {
  float a[3], b[3];
  cross3(a, vel, f);
  cross3(b, vel + 3, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  cross3(r + 3, vel, f + 3);
}
__device__ __forceinline__ void normalize3(float* a) {
  float n = sqrtf(dot3(a, a));
  if (n < 1e-20f) { a[0] = 1; a[1] = a[2] = 0; return; }
  float inv = 1.0f / n;
  a[0] *= inv; a[1] *= inv; a[2] *= inv;
}

// Named barriers for a group of warps inside the CTA (bar = id | thread_count << 8; 0 = no barrier).
__device__ __forceinline__ void group_sync(int bar) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar & 255), "r"(bar >> 8) : "memory");
}
__device__ __forceinline__ bool group_sync_or(int bar, bool pred) {
  unsigned out, in = pred;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 p, %1, 0;\n"
      "bar.red.or.pred q, %2, %3, p;\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n"
      : "=r"(out)
      : "r"(in), "r"(bar & 255), "r"(bar >> 8)
      : "memory");
  return out != 0;
}

// ----------------------------------------------------------------------------- dense algebra (single copies)
// y = A x for the symmetric dense matrix in shared memory (full storage).  Row stride a multiple of 4
// (the n <= 32 layout): lane i reads row i and x as float4 (conflict-free: 7 * lane mod 8 is a
// permutation), the tail below n scalar.
__device__ __noinline__ void symv(float* y, const float* A, const float* x, int n, int ld, int lane) {
  _Pragma("unroll 1") for (int i = lane; i < n; i += 32) {
    float s = 0;
    const float* r = A + i * ld;
    int k = 0;
    if ((ld & 3) == 0) {
      const float4 *r4 = reinterpret_cast<const float4*>(r), *x4 = reinterpret_cast<const float4*>(x);
      for (; k + 4 <= n; k += 4) {
        float4 a = r4[k >> 2], b = x4[k >> 2];
        s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      }
    }
    for (; k < n; k++) s += r[k] * x[k];
    y[i] = s;
  }
  __syncwarp();
}

// dst = src for a whole n x ld matrix (ld a multiple of 4: float4 copy) or its lower triangle (odd ld)
__device__ __noinline__ void copy_matrix(float* dst, const float* src, int n, int ld, int lane) {
  if ((ld & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    _Pragma("unroll 1") for (int i = lane; i < (n * ld) >> 2; i += 32) d4[i] = s4[i];
  } else {
    _Pragma("unroll 1") for (int i = lane; i < n; i += 32)
      for (int k = 0; k <= i; k++) dst[i * ld + k] = src[i * ld + k];
  }
  __syncwarp();
}

// ---- n <= 32: one matrix row per lane held in REGISTERS, columns exchanged by warp shuffles ------------
// a += s * v[0..NT) (v in shared memory, 16-byte aligned); cmask: the 4-column chunks of v that hold anything
template <int NT>
__device__ __forceinline__ void rank1_row(float (&a)[NT], float s, const float* v, unsigned cmask) {
  const float4* v4 = reinterpret_cast<const float4*>(v);
#pragma unroll
  for (int k = 0; k < NT / 4; k++) {
    if ((cmask >> k) & 1) {
      float4 t = v4[k];
      a[4 * k] = fmaf(s, t.x, a[4 * k]); a[4 * k + 1] = fmaf(s, t.y, a[4 * k + 1]);
      a[4 * k + 2] = fmaf(s, t.z, a[4 * k + 2]); a[4 * k + 3] = fmaf(s, t.w, a[4 * k + 3]);
    }
  }
}
// packed Jacobian rows: position of dof i among the set bits of the contact's dof mask
__device__ __forceinline__ int mask_pos(unsigned lo, unsigned hi, int i) {
  return i < 32 ? __popc(lo & ((1u << i) - 1)) : __popc(lo) + __popc(hi & ((1u << (i - 32)) - 1));
}
__device__ __forceinline__ bool mask_has(unsigned lo, unsigned hi, int i) { return ((i < 32 ? lo >> i : hi >> (i - 32)) & 1) != 0; }
// dense[0..ldj) = the packed row expanded (zeros elsewhere); every lane takes part
__device__ __forceinline__ void expand_row(float* dense, const float* Jr, unsigned lo, unsigned hi, int ldj, int lane) {
  _Pragma("unroll 1") for (int i = lane; i < ldj; i += 32) dense[i] = mask_has(lo, hi, i) ? Jr[mask_pos(lo, hi, i)] : 0.f;
  __syncwarp();
}
// Right-looking Cholesky of the register rows a[0..NT) (lane i = row i; entries k > i are don't-care,
// rows >= n are identity rows, lanes >= NT idle) FUSED with the solve of L L^T x = b:
//  * forward substitution rides along the factorisation (y_j is final once column j is scaled);
//  * L is then parked in shared memory (row stride ld, float4 stores) and the backward substitution reads
//    row j across the lanes (one conflict-free load per column, all issued up front).
// 2 instructions per (row, column) pair + ~11 per column; straight-line convergent code.
template <int NT>
__device__ __forceinline__ void chol_solve_rows(float (&a)[NT], float* Ls, int n, int ld, float* xs, int lane) {
  float x = lane < n ? xs[lane] : 0.f, rinv_own = 1.f;
#pragma unroll
  for (int j = 0; j < NT; j++) {
    float piv = __shfl_sync(FULL, a[j], j);
    float rinv = rsqrtf(fmaxf(piv, MINVAL));
    float l = a[j] * rinv;
    a[j] = l;
    float yj = __shfl_sync(FULL, x, j) * rinv;
    if (lane == j) { rinv_own = rinv; x = yj; }
    if (lane > j) x = fmaf(-l, yj, x);
#pragma unroll
    for (int k = j + 1; k < NT; k++) a[k] = fmaf(-l, __shfl_sync(FULL, l, k), a[k]);
  }
  if (lane < n) {
    float4* L4 = reinterpret_cast<float4*>(Ls + lane * ld);
#pragma unroll
    for (int k = 0; k < NT / 4; k++)
      if (4 * k < ld) L4[k] = make_float4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]);
  }
  __syncwarp();
  float lj[NT];
#pragma unroll
  for (int j = 0; j < NT; j++) lj[j] = (j < n && lane < j) ? Ls[j * ld + lane] : 0.f;
#pragma unroll
  for (int j = NT - 1; j >= 0; j--) {
    float xj = __shfl_sync(FULL, x * rinv_own, j);
    x = (lane == j) ? xj : fmaf(-lj[j], xj, x);
  }
  if (lane < n) xs[lane] = x;
  __syncwarp();
}

// ----------------------------------------------------------------------------- constraint rows (single copies)
struct Rows {
  const int *sd1, *sd2;
  int* info;                       // type | state << 4 | id << 8
  const float *sc1, *sc2, *J, *eD, *eR, *efl, *con;
  float *jar, *jv, *force;
  int ns, nefc, ncon, ldj, nv;
};
#define INFO_TYPE(x) ((x) & 15)
#define INFO_STATE(x) (((x) >> 4) & 15)
#define INFO_ID(x) ((x) >> 8)

// Row-parallel evaluation of the constraint cost at jar (+ alpha*jv).
// WRITE=true : alpha ignored; writes force/state, returns cost in c_out.
// WRITE=false: line-search probe; returns cost, derivative g, curvature h (constraint part only).
template <bool WRITE>
__device__ __noinline__ float4 eval_constraints(const Rows R, float alpha, int lane) {
  const float *jar = R.jar, *jv = R.jv, *eD = R.eD, *eR = R.eR, *efl = R.efl;
  float c = 0, g = 0, h = 0;
  _Pragma("unroll 1") for (int r = lane; r < R.nefc; r += 32) {
    int inf = R.info[r], tp = INFO_TYPE(inf);
    if (tp == CNSTR_CONTACT_ELLIPTIC) continue;
    float D = eD[r], v = WRITE ? 0.f : jv[r], x = WRITE ? jar[r] : jar[r] + alpha * v;
    float f = 0; int st = ST_SATISFIED;
    if (tp == CNSTR_EQUALITY) {
      c += 0.5f * D * x * x; g += D * x * v; h += D * v * v; f = -D * x; st = ST_QUADRATIC;
    } else if (tp == CNSTR_FRICTION) {
      float fl = efl[r], Rr = eR[r];
      if (x <= -Rr * fl) { c += -0.5f * Rr * fl * fl - fl * x; g += -fl * v; f = fl; st = ST_LINEARNEG; }
      else if (x >= Rr * fl) { c += -0.5f * Rr * fl * fl + fl * x; g += fl * v; f = -fl; st = ST_LINEARPOS; }
      else { c += 0.5f * D * x * x; g += D * x * v; h += D * v * v; f = -D * x; st = ST_QUADRATIC; }
    } else {  // limit or frictionless contact
      if (x < 0) { c += 0.5f * D * x * x; g += D * x * v; h += D * v * v; f = -D * x; st = ST_QUADRATIC; }
    }
    if (WRITE) { R.force[r] = f; R.info[r] = (inf & ~0xF0) | (st << 4); }
  }
  _Pragma("unroll 1") for (int k = lane; k < R.ncon; k += 32) {
    const float* con = R.con + k * CON_STRIDE;
    int dim = __float_as_int(con[C_DIM]);
    if (dim == 1) continue;
    int i = __float_as_int(con[C_EFC]);
    if (i < 0) continue;
    float mu = con[C_MU];
    float x0 = WRITE ? jar[i] : jar[i] + alpha * jv[i];
    float N = x0 * mu, N1 = WRITE ? 0.f : jv[i] * mu, TT = 0, UV = 0, VV = 0;
    for (int j = 1; j < dim; j++) {
      float fj = con[C_FRICTION + j - 1];
      float v = WRITE ? 0.f : jv[i + j] * fj, u = (WRITE ? jar[i + j] : jar[i + j] + alpha * jv[i + j]) * fj;
      TT += u * u; UV += u * v; VV += v * v;
    }
    float T = sqrtf(TT);
    int st;
    if (N >= mu * T || (T <= 0 && N >= 0)) {
      st = ST_SATISFIED;
      if (WRITE) for (int j = 0; j < dim; j++) R.force[i + j] = 0;
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
      st = ST_QUADRATIC;
      for (int j = 0; j < dim; j++) {
        float D = eD[i + j], v = WRITE ? 0.f : jv[i + j], x = WRITE ? jar[i + j] : jar[i + j] + alpha * v;
        c += 0.5f * D * x * x; g += D * x * v; h += D * v * v;
        if (WRITE) R.force[i + j] = -D * x;
      }
    } else {
      st = ST_CONE;
      float Dm = eD[i] / fmaxf(mu * mu * (1 + mu * mu), MINVAL);
      float NT = N - mu * T;
      c += 0.5f * Dm * NT * NT;
      if (WRITE) {
        float f0 = -Dm * NT * mu;
        R.force[i] = f0;
        for (int j = 1; j < dim; j++) {
          float fj = con[C_FRICTION + j - 1];
          R.force[i + j] = -f0 / T * (jar[i + j] * fj) * fj;
        }
      } else {
        float T1 = UV / T, T2d = (VV - T1 * T1) / T, NT1 = N1 - mu * T1;
        g += Dm * NT * NT1; h += Dm * (NT1 * NT1 - NT * mu * T2d);
      }
    }
    if (WRITE) for (int j = 0; j < dim; j++) R.info[i + j] = (R.info[i + j] & ~0xF0) | (st << 4);
  }
  float4 res = make_float4(warp_sum(c), 0.f, 0.f, 0.f);
  if (!WRITE) { res.y = warp_sum(g); res.z = warp_sum(h); }
  if (WRITE) __syncwarp();
  return res;   // cost, derivative, curvature
}

// y[r] = J[r,:] . x for all rows (simple rows + packed contact rows)
__device__ __noinline__ void mul_J(const Rows R, float* y, const float* x, int lane) {
  _Pragma("unroll 1") for (int r = lane; r < R.nefc; r += 32) {
    float s;
    if (r < R.ns) {
      s = R.sc1[r] * x[R.sd1[r]];
      int d2 = R.sd2[r];
      if (d2 >= 0) s += R.sc2[r] * x[d2];
    } else {
      const float* con = R.con + INFO_ID(R.info[r]) * CON_STRIDE;
      unsigned lo = __float_as_uint(con[C_MASKLO]), hi = __float_as_uint(con[C_MASKHI]);
      int w = __popc(lo) + __popc(hi);
      const float* Jr = R.J + __float_as_int(con[C_JOFS]) + (r - __float_as_int(con[C_EFC])) * w;
      s = 0;
      int k = 0;
      while (lo) { int d = __ffs(lo) - 1; lo &= lo - 1; s += Jr[k++] * x[d]; }
      while (hi) { int d = __ffs(hi) - 1; hi &= hi - 1; s += Jr[k++] * x[32 + d]; }
    }
    y[r] = s;
  }
  __syncwarp();
}

// y[i] = sum_r J[r,i] f[r]
__device__ __noinline__ void mul_JT(const Rows R, float* y, const float* f, int lane) {
  _Pragma("unroll 1") for (int i = lane; i < R.nv; i += 32) {
    float s = 0;
    for (int r = 0; r < R.ns; r++) {
      if (R.sd1[r] == i) s += R.sc1[r] * f[r];
      if (R.sd2[r] == i) s += R.sc2[r] * f[r];
    }
    for (int c = 0; c < R.ncon; c++) {
      const float* con = R.con + c * CON_STRIDE;
      unsigned lo = __float_as_uint(con[C_MASKLO]), hi = __float_as_uint(con[C_MASKHI]);
      if (!mask_has(lo, hi, i)) continue;
      int w = __popc(lo) + __popc(hi), dim = __float_as_int(con[C_DIM]), r0 = __float_as_int(con[C_EFC]);
      const float* Jc = R.J + __float_as_int(con[C_JOFS]) + mask_pos(lo, hi, i);
      for (int j = 0; j < dim; j++) s += Jc[j * w] * f[r0 + j];
    }
    y[i] = s;
  }
  __syncwarp();
}

// n <= 32 variant: the sparse rows are folded into a shared-memory copy of M, then lane i pulls row i
// into registers (float4 loads), accumulates the dense contact rows as rank-1 updates (one unrolled body
// shared by quadratic rows and cone blocks), factors in registers and solves H x = b for the vector xs in
// place (chol_solve_rows).  With R.nefc == 0 this is a plain Cholesky solve with M (H may alias M).
// `alive` = this warp is still iterating (vote returned to the caller); `work` = it has a system to solve;
// `sync` = barrier in front of the ~1000-instruction straight-line factorisation so that all warps of the
// group stream it through the instruction cache together (every warp of the group must then make this
// call, working or not).
template <int NT>
__device__ __noinline__ bool hessian_solve(const Rows R, float* H, const float* M, float* tmpJ, int ld, float* xs, int lane,
                                           bool alive, bool work, int sync) {
  int nv = R.nv, ldj = R.ldj, ns = R.ns;
  float h[NT];
  if (work) {
  if (H != M) copy_matrix(H, M, nv, ld, lane);
  if (lane < nv) {
    float* Hi = H + lane * ld;
#pragma unroll 1
    for (int r = 0; r < ns; r++) {
      if (INFO_STATE(R.info[r]) != ST_QUADRATIC) continue;
      int d1 = R.sd1[r], d2 = R.sd2[r];
      float D = R.eD[r], c1 = R.sc1[r], c2 = R.sc2[r];
      if (d1 == lane) Hi[lane] += D * c1 * c1;
      if (d2 == lane) Hi[lane] += D * c2 * c2;
      if (d2 >= 0 && max(d1, d2) == lane) Hi[min(d1, d2)] += D * c1 * c2;
    }
  }
  __syncwarp();
  {
    const float4* H4 = reinterpret_cast<const float4*>(H + min(lane, nv - 1) * ld);
#pragma unroll
    for (int k = 0; k < NT / 4; k++) {
      float4 t = (4 * k < ld) ? H4[k] : make_float4(0.f, 0.f, 0.f, 0.f);
      h[4 * k] = t.x; h[4 * k + 1] = t.y; h[4 * k + 2] = t.z; h[4 * k + 3] = t.w;
    }
    if (lane >= nv) {
#pragma unroll
      for (int k = 0; k < NT; k++) h[k] = (k == lane) ? 1.f : 0.f;
    }
  }
#pragma unroll 1
  for (int r = ns; r < R.nefc; r++) {
    int inf = R.info[r], st = INFO_STATE(inf);
    if (st != ST_QUADRATIC && st != ST_CONE) continue;
    // packed rows of this contact: lane i's entry sits at position pos (if dof i is in the mask); the rank-1 operands are
    // expanded into dense scratch rows, and only the 4-column chunks that hold a dof of the contact are touched
    const float* con = R.con + INFO_ID(inf) * CON_STRIDE;
    const unsigned lo = __float_as_uint(con[C_MASKLO]), hi = __float_as_uint(con[C_MASKHI]), cm = __float_as_uint(con[C_CMASK]);
    const int w = __popc(lo) + __popc(hi), r0 = __float_as_int(con[C_EFC]);
    const float* Jc = R.J + __float_as_int(con[C_JOFS]);
    const bool mine = lane < nv && mask_has(lo, hi, lane);
    const int pos = mask_pos(lo, hi, lane);
    float* dense = tmpJ + 2 * ldj;
    if (st == ST_QUADRATIC) {
      float jl = mine ? Jc[(r - r0) * w + pos] : 0.f;
      __syncwarp();
      if (lane < ldj) dense[lane] = jl;
      __syncwarp();
      rank1_row<NT>(h, R.eD[r] * jl, dense, cm);
    } else {
      int dim = __float_as_int(con[C_DIM]);
      float mu = con[C_MU], U[6], sc[6], T2 = 0;
      sc[0] = mu; U[0] = R.jar[r] * mu;
#pragma unroll
      for (int j = 1; j < 6; j++) {   // fixed bound: U / sc stay in registers
        sc[j] = j < dim ? con[C_FRICTION + j - 1] : 0.f; U[j] = j < dim ? R.jar[r + j] * sc[j] : 0.f; T2 += U[j] * U[j];
      }
      float N = U[0], T = sqrtf(T2), Dm = R.eD[r] / fmaxf(mu * mu * (1 + mu * mu), MINVAL);
      float iT = 1.0f / T;
      // Exact cone Hessian in the scaled coordinates U (s = Dm/2 (N - mu T)^2):
      //   Hc = Dm g g^T + c (I_t - u u^T),  g = (1, -mu u),  u = U_t / T,  c = -Dm mu (N - mu T) / T  (> 0)
      // so J^T Hc J is dim+1 rank-1 updates: Dm vg vg^T - c vu vu^T + sum_j c sc_j^2 J_j J_j^T with
      // vg = sum_j sc_j g_j J_j and vu = sum_{j>=1} sc_j u_j J_j.
      float c = -Dm * mu * (N - mu * T) * iT;
      float vg = 0.f, vu = 0.f;
      if (mine) {
        vg = mu * Jc[pos];
#pragma unroll
        for (int j = 1; j < 6; j++) {
          if (j < dim) {
            float ww = sc[j] * U[j] * iT, Jj = Jc[j * w + pos];
            vu = fmaf(ww, Jj, vu); vg = fmaf(-mu * ww, Jj, vg);
          }
        }
      }
      __syncwarp();
      if (lane < ldj) { tmpJ[lane] = vg; tmpJ[ldj + lane] = vu; }
      __syncwarp();
      rank1_row<NT>(h, Dm * vg, tmpJ, cm);
      rank1_row<NT>(h, -c * vu, tmpJ + ldj, cm);
#pragma unroll 1
      for (int j = 1; j < dim; j++) {
        float fj = con[C_FRICTION + j - 1], jl = mine ? Jc[j * w + pos] : 0.f;
        __syncwarp();
        if (lane < ldj) dense[lane] = jl;
        __syncwarp();
        rank1_row<NT>(h, c * fj * fj * jl, dense, cm);
      }
      r += dim - 1;
    }
  }
  }
  bool any = sync ? group_sync_or(sync, alive) : alive;
  if (work) chol_solve_rows<NT>(h, H, nv, ld, xs, lane);
  return any;
}
// ---- 32 < n <= 64: TWO matrix rows per lane (row `lane` with its 32 leading columns in a0, row 32 + lane with
// NT2 columns in a1).  Same right-looking scheme: the column shuffles of the first tile serve both rows, the
// trailing columns 32.. only live in the second tile.  Rows in [n, NT2) are identity rows, rows >= NT2 are zero.
template <int NT2>
__device__ __forceinline__ void rank1_row2(float (&a0)[32], float (&a1)[NT2], float s0, float s1, const float* v, unsigned cmask) {
  const float4* v4 = reinterpret_cast<const float4*>(v);
#pragma unroll
  for (int k = 0; k < NT2 / 4; k++) {
    if ((cmask >> k) & 1) {
      float4 t = v4[k];
      if (k < 8) {
        a0[4 * k] = fmaf(s0, t.x, a0[4 * k]); a0[4 * k + 1] = fmaf(s0, t.y, a0[4 * k + 1]);
        a0[4 * k + 2] = fmaf(s0, t.z, a0[4 * k + 2]); a0[4 * k + 3] = fmaf(s0, t.w, a0[4 * k + 3]);
      }
      a1[4 * k] = fmaf(s1, t.x, a1[4 * k]); a1[4 * k + 1] = fmaf(s1, t.y, a1[4 * k + 1]);
      a1[4 * k + 2] = fmaf(s1, t.z, a1[4 * k + 2]); a1[4 * k + 3] = fmaf(s1, t.w, a1[4 * k + 3]);
    }
  }
}
template <int NT2>
__device__ __forceinline__ void chol_solve_rows2(float (&a0)[32], float (&a1)[NT2], float* Ls, int n, int ld, float* xs, int lane) {
  constexpr int N1 = NT2 - 32;
  const bool has1 = 32 + lane < n;
  float x0 = xs[lane], x1 = has1 ? xs[32 + lane] : 0.f, rinv0 = 1.f, rinv1 = 1.f;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    float piv = __shfl_sync(FULL, a0[j], j);
    float rinv = rsqrtf(fmaxf(piv, MINVAL));
    float l0 = a0[j] * rinv, l1 = a1[j] * rinv;
    a0[j] = l0; a1[j] = l1;
    float yj = __shfl_sync(FULL, x0, j) * rinv;
    if (lane == j) { rinv0 = rinv; x0 = yj; }
    if (lane > j) x0 = fmaf(-l0, yj, x0);
    x1 = fmaf(-l1, yj, x1);
#pragma unroll
    for (int k = j + 1; k < 32; k++) {
      float t = __shfl_sync(FULL, l0, k);
      a0[k] = fmaf(-l0, t, a0[k]); a1[k] = fmaf(-l1, t, a1[k]);
    }
#pragma unroll
    for (int k = 0; k < N1; k++) a1[32 + k] = fmaf(-l1, __shfl_sync(FULL, l1, k), a1[32 + k]);
  }
#pragma unroll
  for (int j = 32; j < NT2; j++) {
    float piv = __shfl_sync(FULL, a1[j], j - 32);
    float rinv = rsqrtf(fmaxf(piv, MINVAL));
    float l1 = a1[j] * rinv;
    a1[j] = l1;
    float yj = __shfl_sync(FULL, x1, j - 32) * rinv;
    if (lane == j - 32) { rinv1 = rinv; x1 = yj; }
    if (lane > j - 32) x1 = fmaf(-l1, yj, x1);
#pragma unroll
    for (int k = j + 1; k < NT2; k++) a1[k] = fmaf(-l1, __shfl_sync(FULL, l1, k - 32), a1[k]);
  }
  {
    float4* L4 = reinterpret_cast<float4*>(Ls + lane * ld);
#pragma unroll
    for (int k = 0; k < 8; k++) L4[k] = make_float4(a0[4 * k], a0[4 * k + 1], a0[4 * k + 2], a0[4 * k + 3]);
    if (has1) {
      float4* L41 = reinterpret_cast<float4*>(Ls + (32 + lane) * ld);
#pragma unroll
      for (int k = 0; k < NT2 / 4; k++)
        if (4 * k < ld) L41[k] = make_float4(a1[4 * k], a1[4 * k + 1], a1[4 * k + 2], a1[4 * k + 3]);
    }
  }
  __syncwarp();
  // backward substitution: column `lane` (and 32 + lane) of L below the diagonal, one conflict-free load per row
  float lj0[NT2], lj1[N1];
#pragma unroll
  for (int j = 0; j < NT2; j++) lj0[j] = (j < n && lane < j) ? Ls[j * ld + lane] : 0.f;
#pragma unroll
  for (int j = 32; j < NT2; j++) lj1[j - 32] = (j < n && 32 + lane < j) ? Ls[j * ld + 32 + lane] : 0.f;
#pragma unroll
  for (int j = NT2 - 1; j >= 32; j--) {
    float xj = __shfl_sync(FULL, x1 * rinv1, j - 32);
    x1 = (lane == j - 32) ? xj : fmaf(-lj1[j - 32], xj, x1);
    x0 = fmaf(-lj0[j], xj, x0);
  }
#pragma unroll
  for (int j = 31; j >= 0; j--) {
    float xj = __shfl_sync(FULL, x0 * rinv0, j);
    x0 = (lane == j) ? xj : fmaf(-lj0[j], xj, x0);
  }
  xs[lane] = x0;
  if (has1) xs[32 + lane] = x1;
  __syncwarp();
}

// H = M + J^T D J, factor, solve: the 32 < n <= NT2 counterpart of hessian_solve (same contract)
template <int NT2>
__device__ __noinline__ bool hessian_solve2(const Rows R, float* H, const float* M, float* tmpJ, int ld, float* xs, int lane,
                                            bool alive, bool work, int sync) {
  int nv = R.nv, ldj = R.ldj, ns = R.ns;
  const bool has1 = 32 + lane < nv;
  float h0[32], h1[NT2];
  if (work) {
    if (H != M) copy_matrix(H, M, nv, ld, lane);
#pragma unroll 1
    for (int i = lane; i < nv; i += 32) {
      float* Hi = H + i * ld;
#pragma unroll 1
      for (int r = 0; r < ns; r++) {
        if (INFO_STATE(R.info[r]) != ST_QUADRATIC) continue;
        int d1 = R.sd1[r], d2 = R.sd2[r];
        float D = R.eD[r], c1 = R.sc1[r], c2 = R.sc2[r];
        if (d1 == i) Hi[i] += D * c1 * c1;
        if (d2 == i) Hi[i] += D * c2 * c2;
        if (d2 >= 0 && max(d1, d2) == i) Hi[min(d1, d2)] += D * c1 * c2;
      }
    }
    __syncwarp();
    {
      const float4* H4 = reinterpret_cast<const float4*>(H + lane * ld);
#pragma unroll
      for (int k = 0; k < 8; k++) { float4 t = H4[k]; h0[4 * k] = t.x; h0[4 * k + 1] = t.y; h0[4 * k + 2] = t.z; h0[4 * k + 3] = t.w; }
      const float4* H41 = reinterpret_cast<const float4*>(H + (has1 ? 32 + lane : 0) * ld);
#pragma unroll
      for (int k = 0; k < NT2 / 4; k++) {
        float4 t = (has1 && 4 * k < ld) ? H41[k] : make_float4(0.f, 0.f, 0.f, 0.f);
        h1[4 * k] = t.x; h1[4 * k + 1] = t.y; h1[4 * k + 2] = t.z; h1[4 * k + 3] = t.w;
      }
      if (!has1) {
#pragma unroll
        for (int k = 32; k < NT2; k++) h1[k] = (k == 32 + lane) ? 1.f : 0.f;
      } else {
#pragma unroll
        for (int k = 32; k < NT2; k++) if (k >= nv) h1[k] = 0.f;   // row padding behind column nv
      }
    }
#pragma unroll 1
    for (int r = ns; r < R.nefc; r++) {
      int inf = R.info[r], st = INFO_STATE(inf);
      if (st != ST_QUADRATIC && st != ST_CONE) continue;
      const float* con = R.con + INFO_ID(inf) * CON_STRIDE;   // packed rows, dense scratch operands, chunk mask: see hessian_solve
      const unsigned lo = __float_as_uint(con[C_MASKLO]), hi = __float_as_uint(con[C_MASKHI]), cm = __float_as_uint(con[C_CMASK]);
      const int w = __popc(lo) + __popc(hi), r0 = __float_as_int(con[C_EFC]);
      const float* Jc = R.J + __float_as_int(con[C_JOFS]);
      const bool mine0 = ((lo >> lane) & 1) != 0, mine1 = has1 && ((hi >> lane) & 1) != 0;
      const int pos0 = __popc(lo & ((1u << lane) - 1)), pos1 = __popc(lo) + __popc(hi & ((1u << lane) - 1));
      float* dense = tmpJ + 2 * ldj;
      if (st == ST_QUADRATIC) {
        const float* Jr = Jc + (r - r0) * w;
        float j0 = mine0 ? Jr[pos0] : 0.f, j1 = mine1 ? Jr[pos1] : 0.f, D = R.eD[r];
        __syncwarp();
        dense[lane] = j0;
        if (32 + lane < ldj) dense[32 + lane] = j1;
        __syncwarp();
        rank1_row2<NT2>(h0, h1, D * j0, D * j1, dense, cm);
      } else {
        int dim = __float_as_int(con[C_DIM]);
        float mu = con[C_MU], U[6], sc[6], T2 = 0;
        sc[0] = mu; U[0] = R.jar[r] * mu;
#pragma unroll
        for (int j = 1; j < 6; j++) {
          sc[j] = j < dim ? con[C_FRICTION + j - 1] : 0.f; U[j] = j < dim ? R.jar[r + j] * sc[j] : 0.f; T2 += U[j] * U[j];
        }
        float N = U[0], T = sqrtf(T2), Dm = R.eD[r] / fmaxf(mu * mu * (1 + mu * mu), MINVAL);
        float iT = 1.0f / T;
        float c = -Dm * mu * (N - mu * T) * iT;   // see hessian_solve for the rank-1 form of the cone Hessian
        float vg0 = mine0 ? mu * Jc[pos0] : 0.f, vu0 = 0, vg1 = mine1 ? mu * Jc[pos1] : 0.f, vu1 = 0;
#pragma unroll
        for (int j = 1; j < 6; j++) {
          if (j < dim) {
            float ww = sc[j] * U[j] * iT, Ja = mine0 ? Jc[j * w + pos0] : 0.f, Jb = mine1 ? Jc[j * w + pos1] : 0.f;
            vu0 = fmaf(ww, Ja, vu0); vg0 = fmaf(-mu * ww, Ja, vg0);
            vu1 = fmaf(ww, Jb, vu1); vg1 = fmaf(-mu * ww, Jb, vg1);
          }
        }
        __syncwarp();
        tmpJ[lane] = vg0; tmpJ[ldj + lane] = vu0;
        if (32 + lane < ldj) { tmpJ[32 + lane] = vg1; tmpJ[ldj + 32 + lane] = vu1; }
        __syncwarp();
        rank1_row2<NT2>(h0, h1, Dm * vg0, Dm * vg1, tmpJ, cm);
        rank1_row2<NT2>(h0, h1, -c * vu0, -c * vu1, tmpJ + ldj, cm);
#pragma unroll 1
        for (int j = 1; j < dim; j++) {
          float fj = con[C_FRICTION + j - 1], cf = c * fj * fj;
          float j0 = mine0 ? Jc[j * w + pos0] : 0.f, j1 = mine1 ? Jc[j * w + pos1] : 0.f;
          __syncwarp();
          dense[lane] = j0;
          if (32 + lane < ldj) dense[32 + lane] = j1;
          __syncwarp();
          rank1_row2<NT2>(h0, h1, cf * j0, cf * j1, dense, cm);
        }
        r += dim - 1;
      }
    }
  }
  bool any = sync ? group_sync_or(sync, alive) : alive;
  if (work) chol_solve_rows2<NT2>(h0, h1, H, nv, ld, xs, lane);
  return any;
}

// One register tile per solve-kernel instantiation (TILE <= 32: one row per lane, TILE columns; TILE > 32: two rows
// per lane).  28 columns cover the Stretch robot's nv = 26 with 24 % fewer pair updates than 32; 44 is the
// reference's default scene (robot + three free bodies).  The host picks the kernel (ss_solve_tile).
template <int TILE>
__device__ __forceinline__ bool hessian_solve_any(const Rows& R, float* H, const float* M, float* tmpJ, int ld, float* xs, int lane,
                                                  bool alive, bool work, int sync) {
  if constexpr (TILE <= 32) return hessian_solve<TILE>(R, H, M, tmpJ, ld, xs, lane, alive, work, sync);
  else return hessian_solve2<TILE>(R, H, M, tmpJ, ld, xs, lane, alive, work, sync);
}
// x = A^-1 b for an SPD matrix in shared memory (factor left in A)
template <int TILE>
__device__ __forceinline__ void chol_solve_any(float* A, int n, int ld, float* xs, int lane, bool work, int sync) {
  Rows R;
  R.nv = n; R.ns = 0; R.nefc = 0; R.ldj = (n + 3) & ~3;
  R.sd1 = R.sd2 = nullptr; R.info = nullptr; R.sc1 = R.sc2 = R.J = R.eD = R.eR = R.efl = R.con = nullptr;
  R.jar = R.jv = R.force = nullptr; R.ncon = 0;
  hessian_solve_any<TILE>(R, A, A, nullptr, ld, xs, lane, work, work, sync);
}

// ----------------------------------------------------------------------------- S1: kinematics + inertias + dof axes
// One pass over the tree levels: body frames, spatial inertia about the tree's reference point
// (the root body's origin) and the motion axis of every dof [upstream mj_kinematics + mj_comPos;
// MuJoCo uses the subtree COM as reference point, any common point gives the same dynamics].
//
// fp32 note.  Spatial quantities about a far reference point are fine for the bias forces, but the
// joint-space inertia of a light distal link is a difference of terms ~ m d^2 (d = distance to the
// reference point, ~0.7 m for the gripper) that are 300x larger than the result.  The mass matrix is
// therefore built from LOCAL quantities that never see world-magnitude offsets: every body keeps its inertia
// about its own COM with the COM relative to its own origin (crb, COM form), its origin relative to the
// parent's origin (dpos, accumulated from the local joint displacements) and the anchor of each hinge
// relative to the body origin (danchor); see crb_mass_matrix.
__device__ __forceinline__ void kinematics(const DevModel& m, float* S, int lane) {
  const EnvLayout& o = m.L;
  float *xpos = S + o.xpos, *xquat = S + o.xquat, *xmat = S + o.xmat, *cinert = S + o.cinert, *cdof = S + o.cdof;
  float *crb = S + o.crb, *dpos = S + o.dpos, *danchor = S + o.danchor;
  const float* qpos = S + o.qpos;
  if (lane == 0) {
    xpos[0] = xpos[1] = xpos[2] = 0; xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
    for (int k = 0; k < 9; k++) xmat[k] = (k % 4 == 0) ? 1.f : 0.f;
    for (int k = 0; k < 10; k++) { cinert[k] = 0; crb[k] = 0; }
    dpos[0] = dpos[1] = dpos[2] = 0;
  }
  __syncwarp();
  for (int lv = 0; lv < m.nlevel; lv++) {
    for (int idx = PKI(lvl_adr)[lv] + lane; idx < PKI(lvl_adr)[lv + 1]; idx += 32) {
      int b = PKI(lvl_body)[idx], p = PKI(body_parentid)[b], jn = PKI(body_jntnum)[b], ja = PKI(body_jntadr)[b];
      int rb = PKI(root_list)[PKI(body_rootidx)[b]];
      float pos[3], q[4], ref[3], d[3];   // d = pos - parent origin, accumulated from local displacements only
      bool isfree = (jn == 1 && PKI(jnt_type)[ja] == JNT_FREE);
      if (isfree) {
        int qa = PKI(jnt_qposadr)[ja];
        pos[0] = qpos[qa]; pos[1] = qpos[qa + 1]; pos[2] = qpos[qa + 2];
        q[0] = qpos[qa + 3]; q[1] = qpos[qa + 4]; q[2] = qpos[qa + 5]; q[3] = qpos[qa + 6];
        quat_normalize(q);
        d[0] = pos[0]; d[1] = pos[1]; d[2] = pos[2];
      } else {
        const float *bq = PKF(body_quat) + 4 * b, *bp = PKF(body_pos) + 3 * b;
        mat_vec(d, xmat + 9 * p, bp);
        pos[0] = xpos[3 * p] + d[0]; pos[1] = xpos[3 * p + 1] + d[1]; pos[2] = xpos[3 * p + 2] + d[2];
        quat_mul(q, xquat + 4 * p, bq);
      }
      if (lv == 0) { ref[0] = pos[0]; ref[1] = pos[1]; ref[2] = pos[2]; }
      else { ref[0] = xpos[3 * rb]; ref[1] = xpos[3 * rb + 1]; ref[2] = xpos[3 * rb + 2]; }
      if (isfree) {
        int da = PKI(jnt_dofadr)[ja];
        float R[9];
        quat2mat(R, q);
        float off[3] = {ref[0] - pos[0], ref[1] - pos[1], ref[2] - pos[2]};
        for (int k = 0; k < 3; k++) {
          float* c = cdof + 6 * (da + k);
          c[0] = c[1] = c[2] = 0; c[3] = (k == 0); c[4] = (k == 1); c[5] = (k == 2);
          float a3[3] = {R[k], R[3 + k], R[6 + k]};
          float* cr = cdof + 6 * (da + 3 + k);
          cr[0] = a3[0]; cr[1] = a3[1]; cr[2] = a3[2];
          cross3(cr + 3, a3, off);
          for (int t = 0; t < 3; t++) { danchor[3 * (da + k) + t] = 0; danchor[3 * (da + 3 + k) + t] = 0; }
        }
      } else {
        for (int j = ja; j < ja + jn; j++) {
          const float *jp = PKF(jnt_pos) + 3 * j, *jax = PKF(jnt_axis) + 3 * j;
          float R[9], anchor[3], axis[3], r[3];
          quat2mat(R, q);
          mat_vec(r, R, jp); anchor[0] = pos[0] + r[0]; anchor[1] = pos[1] + r[1]; anchor[2] = pos[2] + r[2];
          mat_vec(axis, R, jax);
          int qa = PKI(jnt_qposadr)[j], dj = PKI(jnt_dofadr)[j];
          float dq = qpos[qa] - PKF(qpos0)[qa];
          float* c = cdof + 6 * dj;
          // anchor relative to the parent origin for now; made relative to the final body origin below
          danchor[3 * dj] = d[0] + r[0]; danchor[3 * dj + 1] = d[1] + r[1]; danchor[3 * dj + 2] = d[2] + r[2];
          if (PKI(jnt_type)[j] == JNT_SLIDE) {
            pos[0] += axis[0] * dq; pos[1] += axis[1] * dq; pos[2] += axis[2] * dq;
            d[0] += axis[0] * dq; d[1] += axis[1] * dq; d[2] += axis[2] * dq;
            c[0] = c[1] = c[2] = 0; c[3] = axis[0]; c[4] = axis[1]; c[5] = axis[2];
          } else {
            float sn, cs;
            sincosf(0.5f * dq, &sn, &cs);
            float qr[4] = {cs, jax[0] * sn, jax[1] * sn, jax[2] * sn}, qn[4], r2[3];
            quat_mul(qn, q, qr);
            q[0] = qn[0]; q[1] = qn[1]; q[2] = qn[2]; q[3] = qn[3];
            quat2mat(R, q);
            mat_vec(r2, R, jp);
            pos[0] = anchor[0] - r2[0]; pos[1] = anchor[1] - r2[1]; pos[2] = anchor[2] - r2[2];
            d[0] += r[0] - r2[0]; d[1] += r[1] - r2[1]; d[2] += r[2] - r2[2];
            float off[3] = {ref[0] - anchor[0], ref[1] - anchor[1], ref[2] - anchor[2]};
            c[0] = axis[0]; c[1] = axis[1]; c[2] = axis[2];
            cross3(c + 3, axis, off);
          }
        }
        for (int j = ja; j < ja + jn; j++) {
          int dj = PKI(jnt_dofadr)[j];
          danchor[3 * dj] -= d[0]; danchor[3 * dj + 1] -= d[1]; danchor[3 * dj + 2] -= d[2];
        }
        quat_normalize(q);
      }
      xpos[3 * b] = pos[0]; xpos[3 * b + 1] = pos[1]; xpos[3 * b + 2] = pos[2];
      xquat[4 * b] = q[0]; xquat[4 * b + 1] = q[1]; xquat[4 * b + 2] = q[2]; xquat[4 * b + 3] = q[3];
      dpos[3 * b] = d[0]; dpos[3 * b + 1] = d[1]; dpos[3 * b + 2] = d[2];
      float R[9], t[3], qi[4];
      quat2mat(R, q);
#pragma unroll
      for (int k = 0; k < 9; k++) xmat[9 * b + k] = R[k];
      // spatial inertia about the reference point, world axes (cinert) and about the body's own COM (crb, COM form)
      mat_vec(t, R, PKF(body_ipos) + 3 * b);
      float off[3] = {pos[0] + t[0] - ref[0], pos[1] + t[1] - ref[1], pos[2] + t[2] - ref[2]};
      quat_mul(qi, q, PKF(body_iquat) + 4 * b);
      quat2mat(R, qi);
      float I0 = PKF(body_inertia)[3 * b], I1 = PKF(body_inertia)[3 * b + 1], I2 = PKF(body_inertia)[3 * b + 2];
      float mass = PKF(body_mass)[b], o2 = dot3(off, off);
      float* ci = cinert + 10 * b;
      float* cb = crb + 10 * b;
#define IW(r, c) (R[3 * r] * I0 * R[3 * c] + R[3 * r + 1] * I1 * R[3 * c + 1] + R[3 * r + 2] * I2 * R[3 * c + 2])
      cb[0] = IW(0, 0); cb[1] = IW(1, 1); cb[2] = IW(2, 2); cb[3] = IW(0, 1); cb[4] = IW(0, 2); cb[5] = IW(1, 2);
#undef IW
      cb[6] = t[0]; cb[7] = t[1]; cb[8] = t[2]; cb[9] = mass;
      ci[0] = cb[0] + mass * (o2 - off[0] * off[0]);
      ci[1] = cb[1] + mass * (o2 - off[1] * off[1]);
      ci[2] = cb[2] + mass * (o2 - off[2] * off[2]);
      ci[3] = cb[3] - mass * off[0] * off[1];
      ci[4] = cb[4] - mass * off[0] * off[2];
      ci[5] = cb[5] - mass * off[1] * off[2];
      ci[6] = mass * off[0]; ci[7] = mass * off[1]; ci[8] = mass * off[2]; ci[9] = mass;
    }
    __syncwarp();
  }
}

// Composite rigid bodies in COM form -- crb[b] = {inertia about the composite's COM (xx yy zz xy xz yz, world axes),
// COM relative to body b's origin, mass} -- gathered level by level from the leaves, and the dense joint-space
// inertia  M_ij = w_i . Ic w_j + mass v_i . v_j  (j = i or an ancestor dof of i; w, v = angular velocity and
// velocity of the composite's COM per unit rate of the dof).  All lever arms are sums of local offsets.
__device__ __forceinline__ void crb_mass_matrix(const DevModel& m, float* S, int lane) {
  const EnvLayout& o = m.L;
  const float *cdof = S + o.cdof, *dpos = S + o.dpos, *danchor = S + o.danchor;
  float *crb = S + o.crb, *M = S + o.M;
  int nv = m.nv;
  _Pragma("unroll 1") for (int k = lane; k < nv * o.ldm; k += 32) M[k] = 0;
  for (int lv = m.nlevel - 1; lv >= 0; lv--) {
    for (int idx = PKI(lvl_adr)[lv] + lane; idx < PKI(lvl_adr)[lv + 1]; idx += 32) {
      int b = PKI(lvl_body)[idx];
      int c0 = PKI(child_adr)[b], c1 = PKI(child_adr)[b + 1];
      if (c0 == c1) continue;
      float* cb = crb + 10 * b;
      // total mass and COM (relative to b's origin)
      float mass = cb[9], mc[3] = {mass * cb[6], mass * cb[7], mass * cb[8]};
      for (int c = c0; c < c1; c++) {
        int ch = PKI(child_list)[c];
        const float* cc = crb + 10 * ch;
        float mk = cc[9];
        mass += mk;
        for (int t = 0; t < 3; t++) mc[t] += mk * (dpos[3 * ch + t] + cc[6 + t]);
      }
      float inv = mass > 0 ? 1.0f / mass : 0.f, com[3] = {mc[0] * inv, mc[1] * inv, mc[2] * inv};
      // inertia about the new COM: parallel-axis shift of every part by its own (local) offset
      float I[6];
      {
        float dd[3] = {cb[6] - com[0], cb[7] - com[1], cb[8] - com[2]}, mk = cb[9], d2 = dot3(dd, dd);
        I[0] = cb[0] + mk * (d2 - dd[0] * dd[0]); I[1] = cb[1] + mk * (d2 - dd[1] * dd[1]); I[2] = cb[2] + mk * (d2 - dd[2] * dd[2]);
        I[3] = cb[3] - mk * dd[0] * dd[1]; I[4] = cb[4] - mk * dd[0] * dd[2]; I[5] = cb[5] - mk * dd[1] * dd[2];
      }
      for (int c = c0; c < c1; c++) {
        int ch = PKI(child_list)[c];
        const float* cc = crb + 10 * ch;
        float dd[3] = {dpos[3 * ch] + cc[6] - com[0], dpos[3 * ch + 1] + cc[7] - com[1], dpos[3 * ch + 2] + cc[8] - com[2]};
        float mk = cc[9], d2 = dot3(dd, dd);
        I[0] += cc[0] + mk * (d2 - dd[0] * dd[0]); I[1] += cc[1] + mk * (d2 - dd[1] * dd[1]); I[2] += cc[2] + mk * (d2 - dd[2] * dd[2]);
        I[3] += cc[3] - mk * dd[0] * dd[1]; I[4] += cc[4] - mk * dd[0] * dd[2]; I[5] += cc[5] - mk * dd[1] * dd[2];
      }
#pragma unroll
      for (int k = 0; k < 6; k++) cb[k] = I[k];
      cb[6] = com[0]; cb[7] = com[1]; cb[8] = com[2]; cb[9] = mass;
    }
    __syncwarp();
  }
  _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) {
    int cur = PKI(dof_bodyid)[i];
    const float* cb = crb + 10 * cur;
    const float I[6] = {cb[0], cb[1], cb[2], cb[3], cb[4], cb[5]}, mass = cb[9];
    float c[3] = {cb[6], cb[7], cb[8]};   // composite COM relative to the origin of body `cur`
    float Iw[3] = {0, 0, 0}, mv[3] = {0, 0, 0};
    for (int j = i; j >= 0; j = PKI(dof_parentid)[j]) {
      int bj = PKI(dof_bodyid)[j];
      while (cur != bj) { c[0] += dpos[3 * cur]; c[1] += dpos[3 * cur + 1]; c[2] += dpos[3 * cur + 2]; cur = PKI(body_parentid)[cur]; }
      const float* cd = cdof + 6 * j;
      float w[3] = {cd[0], cd[1], cd[2]}, v[3];
      if (w[0] == 0.f && w[1] == 0.f && w[2] == 0.f) { v[0] = cd[3]; v[1] = cd[4]; v[2] = cd[5]; }   // slide / free translation: the axis
      else {
        float arm[3] = {c[0] - danchor[3 * j], c[1] - danchor[3 * j + 1], c[2] - danchor[3 * j + 2]};
        cross3(v, w, arm);
      }
      if (j == i) {
        Iw[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2];
        Iw[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2];
        Iw[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2];
        mv[0] = mass * v[0]; mv[1] = mass * v[1]; mv[2] = mass * v[2];
      }
      float val = dot3(w, Iw) + dot3(v, mv);
      if (j == i) val += PKF(dof_armature)[i];
      M[i * o.ldm + j] = val;
      M[j * o.ldm + i] = val;
    }
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------- S1c: collision
struct Cvx {
  int type, nvert;
  int hadr;                      // first hull vertex of the mesh in hull_vert / hull_edgeadr
  float pos[3], mat[9], size[3];
  const float4* verts;
};

// support point of an analytic primitive in its local frame (direction l, local)
__device__ __forceinline__ void support_prim(const Cvx& g, const float* l, float* r) {
  r[0] = r[1] = r[2] = 0;
  if (g.type == GEOM_SPHERE) {
    float n = sqrtf(dot3(l, l));
    if (n > MINVAL) { float s = g.size[0] / n; r[0] = l[0] * s; r[1] = l[1] * s; r[2] = l[2] * s; }
  } else if (g.type == GEOM_BOX) {
    r[0] = l[0] > 0 ? g.size[0] : -g.size[0]; r[1] = l[1] > 0 ? g.size[1] : -g.size[1];
    r[2] = l[2] > 0 ? g.size[2] : -g.size[2];
  } else if (g.type == GEOM_CYLINDER) {
    float n = sqrtf(l[0] * l[0] + l[1] * l[1]);
    if (n > MINVAL) { r[0] = l[0] / n * g.size[0]; r[1] = l[1] / n * g.size[0]; }
    r[2] = l[2] > 0 ? g.size[1] : -g.size[1];
  }
}
// warp arg-max over per-lane candidates (best, bi): lowest vertex index wins ties.  Two REDUX
// instructions on an order-preserving integer image of the float instead of a 5-round shuffle butterfly.
__device__ __forceinline__ int warp_argmax(float best, int bi) {
  unsigned u = __float_as_uint(best);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  unsigned mx = __reduce_max_sync(FULL, u);
  return (int)__reduce_min_sync(FULL, (u == mx) ? (unsigned)bi : 0x7fffffffu);
}

// support point of one convex geom in world coordinates (mesh hulls: lane-parallel vertex scan)
__device__ __forceinline__ int support(const Cvx& g, const float* dir, float* out, int lane) {
  float l[3], r[3];
  int index = -1;
  matT_vec(l, g.mat, dir);
  if (g.type == GEOM_MESH) {
    float best = -CUDART_INF_F; int bi = 0x7fffffff;
#pragma unroll 4
    for (int i = lane; i < g.nvert; i += 32) {
      float4 v = g.verts[i];
      float s = v.x * l[0] + v.y * l[1] + v.z * l[2];
      if (s > best) { best = s; bi = i; }
    }
    index = warp_argmax(best, bi);
    float4 v = g.verts[index];
    r[0] = v.x; r[1] = v.y; r[2] = v.z;
  } else support_prim(g, l, r);
  mat_vec(out, g.mat, r);
  out[0] += g.pos[0]; out[1] += g.pos[1]; out[2] += g.pos[2];
  return index;
}

struct Spt { float v[3], v1[3], v2[3]; };

// Minkowski-difference support: both geoms are scanned in ONE loop so that the two vertex streams and
// the two reductions overlap (the MPR iteration is a dependent chain of these calls).  Hull vertices are
// read-only float4 loads that live in L1 / L2 (229 KB for the robot's 65 hulls).
__device__ __forceinline__ void msupport_scan(const Cvx& a, const Cvx& b, int na, int nb, const float* la, const float* lb, float* ra,
                                              float* rb, int lane) {
  float besta = -CUDART_INF_F, bestb = -CUDART_INF_F; int bia = 0x7fffffff, bib = 0x7fffffff;
#pragma unroll 2
  for (int i = lane; i < max(na, nb); i += 32) {
    if (i < na) {
      float4 v = __ldg(a.verts + i);
      float t = v.x * la[0] + v.y * la[1] + v.z * la[2];
      if (t > besta) { besta = t; bia = i; }
    }
    if (i < nb) {
      float4 v = __ldg(b.verts + i);
      float t = v.x * lb[0] + v.y * lb[1] + v.z * lb[2];
      if (t > bestb) { bestb = t; bib = i; }
    }
  }
  if (na) { int k = warp_argmax(besta, bia); float4 v = __ldg(a.verts + k); ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; }
  if (nb) { int k = warp_argmax(bestb, bib); float4 v = __ldg(b.verts + k); rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; }
}
__device__ __forceinline__ void msupport(const Cvx& a, const Cvx& b, const float* dir, Spt& s, int lane) {
  float nd[3] = {-dir[0], -dir[1], -dir[2]}, la[3], lb[3], ra[3], rb[3];
  matT_vec(la, a.mat, dir);
  matT_vec(lb, b.mat, nd);
  int na = a.type == GEOM_MESH ? a.nvert : 0, nb = b.type == GEOM_MESH ? b.nvert : 0;
  if (na | nb) msupport_scan(a, b, na, nb, la, lb, ra, rb, lane);
  if (!na) support_prim(a, la, ra);
  if (!nb) support_prim(b, lb, rb);
  mat_vec(s.v1, a.mat, ra);
  mat_vec(s.v2, b.mat, rb);
#pragma unroll
  for (int k = 0; k < 3; k++) { s.v1[k] += a.pos[k]; s.v2[k] += b.pos[k]; s.v[k] = s.v1[k] - s.v2[k]; }
}

#define MPR_TOL 1e-6f
#define MPR_MAXIT 50
#define MPR_EPS 1e-10f

__device__ __forceinline__ void portal_dir(const Spt* p, float* dir) {
  float a[3] = {p[2].v[0] - p[1].v[0], p[2].v[1] - p[1].v[1], p[2].v[2] - p[1].v[2]};
  float b[3] = {p[3].v[0] - p[1].v[0], p[3].v[1] - p[1].v[1], p[3].v[2] - p[1].v[2]};
  cross3(dir, a, b);
  normalize3(dir);
}
__device__ __forceinline__ void expand_portal(Spt* p, const Spt& v4) {
  float v4v0[3];
  cross3(v4v0, v4.v, p[0].v);
  if (dot3(p[1].v, v4v0) > 0) {
    if (dot3(p[2].v, v4v0) > 0) p[1] = v4; else p[3] = v4;
  } else {
    if (dot3(p[3].v, v4v0) > 0) p[2] = v4; else p[1] = v4;
  }
}
__device__ __forceinline__ bool reach_tolerance(const Spt* p, const Spt& v4, const float* dir) {
  float dv4 = dot3(v4.v, dir);
  float m1 = dv4 - dot3(p[1].v, dir), m2 = dv4 - dot3(p[2].v, dir), m3 = dv4 - dot3(p[3].v, dir);
  return fminf(fminf(m1, m2), m3) <= MPR_TOL;
}
__device__ __forceinline__ float origin_tri_dist2(const float* a, const float* b, const float* c, float* w) {
  float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  float ap[3] = {-a[0], -a[1], -a[2]};
  float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { w[0] = a[0]; w[1] = a[1]; w[2] = a[2]; return dot3(w, w); }
  float bp[3] = {-b[0], -b[1], -b[2]};
  float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { w[0] = b[0]; w[1] = b[1]; w[2] = b[2]; return dot3(w, w); }
  float vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) {
    float t = d1 / (d1 - d3);
    w[0] = a[0] + ab[0] * t; w[1] = a[1] + ab[1] * t; w[2] = a[2] + ab[2] * t;
    return dot3(w, w);
  }
  float cp[3] = {-c[0], -c[1], -c[2]};
  float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { w[0] = c[0]; w[1] = c[1]; w[2] = c[2]; return dot3(w, w); }
  float vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) {
    float t = d2 / (d2 - d6);
    w[0] = a[0] + ac[0] * t; w[1] = a[1] + ac[1] * t; w[2] = a[2] + ac[2] * t;
    return dot3(w, w);
  }
  float va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    float t = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    w[0] = b[0] + (c[0] - b[0]) * t; w[1] = b[1] + (c[1] - b[1]) * t; w[2] = b[2] + (c[2] - b[2]) * t;
    return dot3(w, w);
  }
  float den = 1.0f / (va + vb + vc), v = vb * den, u = vc * den;
  w[0] = a[0] + ab[0] * v + ac[0] * u; w[1] = a[1] + ab[1] * v + ac[1] * u; w[2] = a[2] + ab[2] * v + ac[2] * u;
  return dot3(w, w);
}
__device__ __forceinline__ void find_pos(const Spt* p, float* pos) {
  float dir[3], b[4], t[3], sum;
  portal_dir(p, dir);
  cross3(t, p[1].v, p[2].v); b[0] = dot3(t, p[3].v);
  cross3(t, p[3].v, p[2].v); b[1] = dot3(t, p[0].v);
  cross3(t, p[0].v, p[1].v); b[2] = dot3(t, p[3].v);
  cross3(t, p[2].v, p[1].v); b[3] = dot3(t, p[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (sum <= 0) {
    b[0] = 0;
    cross3(t, p[2].v, p[3].v); b[1] = dot3(t, dir);
    cross3(t, p[3].v, p[1].v); b[2] = dot3(t, dir);
    cross3(t, p[1].v, p[2].v); b[3] = dot3(t, dir);
    sum = b[1] + b[2] + b[3];
  }
  float inv = 1.0f / sum;
  for (int k = 0; k < 3; k++) {
    float p1 = 0, p2 = 0;
    for (int i = 0; i < 4; i++) { p1 += p[i].v1[k] * b[i]; p2 += p[i].v2[k] * b[i]; }
    pos[k] = 0.5f * (p1 + p2) * inv;
  }
}

// Minkowski Portal Refinement; warp-uniform control flow, mesh support is lane-parallel.
// One support call site (phase machine) keeps the code small.
__device__ __forceinline__ bool mpr_penetration(const Cvx& A, const Cvx& B, float* depth, float* dir_out, float* pos, int lane) {
  Spt p[4], v4;
  float dir[3], va[3], vb[3];
  for (int k = 0; k < 3; k++) { p[0].v[k] = A.pos[k] - B.pos[k]; p[0].v1[k] = A.pos[k]; p[0].v2[k] = B.pos[k]; }
  if (fabsf(p[0].v[0]) < MPR_EPS && fabsf(p[0].v[1]) < MPR_EPS && fabsf(p[0].v[2]) < MPR_EPS) p[0].v[0] = 1e-5f;
  dir[0] = -p[0].v[0]; dir[1] = -p[0].v[1]; dir[2] = -p[0].v[2];
  normalize3(dir);
  // phases: 0 -> v1, 1 -> v2, 2 -> v3 (portal discovery), 3 -> refinement, 4 -> penetration
  int phase = 0, it = 0;
  for (int guard = 0; guard < 3 * MPR_MAXIT + 120; guard++) {
    msupport(A, B, dir, v4, lane);
    if (phase == 0) {
      p[1] = v4;
      if (dot3(p[1].v, dir) <= 0) return false;
      cross3(dir, p[0].v, p[1].v);
      if (dot3(dir, dir) < MPR_EPS * MPR_EPS) {
        if (dot3(p[1].v, p[1].v) < MPR_EPS * MPR_EPS) { *depth = 0; dir_out[0] = dir_out[1] = dir_out[2] = 0; }
        else { *depth = sqrtf(dot3(p[1].v, p[1].v)); dir_out[0] = p[1].v[0]; dir_out[1] = p[1].v[1]; dir_out[2] = p[1].v[2]; normalize3(dir_out); }
        for (int k = 0; k < 3; k++) pos[k] = 0.5f * (p[1].v1[k] + p[1].v2[k]);
        return true;
      }
      normalize3(dir);
      phase = 1;
    } else if (phase == 1) {
      p[2] = v4;
      if (dot3(p[2].v, dir) <= 0) return false;
      for (int k = 0; k < 3; k++) { va[k] = p[1].v[k] - p[0].v[k]; vb[k] = p[2].v[k] - p[0].v[k]; }
      cross3(dir, va, vb); normalize3(dir);
      if (dot3(dir, p[0].v) > 0) { Spt t = p[1]; p[1] = p[2]; p[2] = t; dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2]; }
      phase = 2; it = 0;
    } else if (phase == 2) {
      p[3] = v4;
      if (dot3(p[3].v, dir) <= 0 || ++it > 100) return false;
      bool cont = false;
      cross3(va, p[1].v, p[3].v);
      if (dot3(va, p[0].v) < -MPR_EPS) { p[2] = p[3]; cont = true; }
      if (!cont) {
        cross3(va, p[3].v, p[2].v);
        if (dot3(va, p[0].v) < -MPR_EPS) { p[1] = p[3]; cont = true; }
      }
      if (cont) {
        for (int k = 0; k < 3; k++) { va[k] = p[1].v[k] - p[0].v[k]; vb[k] = p[2].v[k] - p[0].v[k]; }
        cross3(dir, va, vb); normalize3(dir);
      } else {
        portal_dir(p, dir);
        it = 0;
        phase = (dot3(dir, p[1].v) >= 0) ? 4 : 3;
      }
    } else if (phase == 3) {
      if (dot3(v4.v, dir) < 0 || reach_tolerance(p, v4, dir) || it++ > MPR_MAXIT) return false;
      expand_portal(p, v4);
      portal_dir(p, dir);
      if (dot3(dir, p[1].v) >= 0) { phase = 4; it = 0; }
    } else {
      if (reach_tolerance(p, v4, dir) || it++ > MPR_MAXIT) {
        float w[3];
        float d2 = origin_tri_dist2(p[1].v, p[2].v, p[3].v, w);
        *depth = sqrtf(d2);
        if (*depth < MPR_EPS) { dir_out[0] = dir_out[1] = dir_out[2] = 0; }
        else { dir_out[0] = w[0]; dir_out[1] = w[1]; dir_out[2] = w[2]; normalize3(dir_out); }
        find_pos(p, pos);
        return true;
      }
      expand_portal(p, v4);
      portal_dir(p, dir);
    }
  }
  return false;
}

__device__ __forceinline__ void make_cvx(const DevModel& m, const float* __restrict__ S, int cg, Cvx& c) {
  const EnvLayout& o = m.L;
  int b = PKI(cg_bodyid)[cg];
  c.type = PKI(cg_type)[cg];
  const float* gp = S + o.gpos + 3 * cg;
  c.pos[0] = gp[0]; c.pos[1] = gp[1]; c.pos[2] = gp[2];
  float q[4];
  quat_mul(q, S + o.xquat + 4 * b, PKF(cg_quat) + 4 * cg);
  quat_normalize(q);
  quat2mat(c.mat, q);
  c.size[0] = PKF(cg_size)[3 * cg]; c.size[1] = PKF(cg_size)[3 * cg + 1]; c.size[2] = PKF(cg_size)[3 * cg + 2];
  c.verts = nullptr; c.nvert = 0; c.hadr = 0;
  if (c.type == GEOM_MESH) {
    int mid = PKI(cg_dataid)[cg];
    c.hadr = PKI(mesh_hulladr)[mid];
    c.verts = m.hull_vert + c.hadr; c.nvert = PKI(mesh_hullnum)[mid];
  }
}


// Narrowphase result record in shared memory (written by lane 0; the computation is warp-uniform):
//   [so] count, then per contact c at so + 4 + 8 c: normal[3], dist, pos[3], -.  Up to NP_MAXC contacts.
// [so + 72, so + 96) is scratch for the unperturbed poses of the multiccd queries.
#define NP_MAXC 8
#define NP_SMEM 96   // floats of shared memory per warp of the narrowphase kernel (record + multiccd scratch)
#define NP_EMIT(c, nx, ny, nz, dd, px, py, pz) do { if (w0) { float* o_ = smem + so + 4 + 8 * (c); \
  o_[0] = (nx); o_[1] = (ny); o_[2] = (nz); o_[3] = (dd); o_[4] = (px); o_[5] = (py); o_[6] = (pz); } } while (0)

// box-box [upstream mjc_BoxBox]: 15-axis separating-axis test, then the incident face clipped against the
// reference face (up to 8 points) or one edge-edge point.  Same restatement as oracle/ss_oracle_collision.c:box_box.
__device__ __noinline__ int box_box(const Cvx& A, const Cvx& B, float margin, int so, int lane) {
  const bool w0 = lane == 0;
  float a1[3][3], a2[3][3], pp[3];
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { a1[i][k] = A.mat[3 * k + i]; a2[i][k] = B.mat[3 * k + i]; }
  for (int k = 0; k < 3; k++) pp[k] = B.pos[k] - A.pos[k];
  const float *s1 = A.size, *s2 = B.size;
  float bestf = -CUDART_INF_F, beste = -CUDART_INF_F, nf[3] = {0, 0, 0}, ne[3] = {0, 0, 0};
  int codef = -1, codee = -1;
  for (int w = 0; w < 2; w++)
    for (int i = 0; i < 3; i++) {
      const float* L = w ? a2[i] : a1[i];
      float t = dot3(pp, L), ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) { ra += s1[k] * fabsf(dot3(a1[k], L)); rb += s2[k] * fabsf(dot3(a2[k], L)); }
      float sep = fabsf(t) - ra - rb;
      if (sep > margin) return 0;
      if (sep > bestf) { bestf = sep; codef = 3 * w + i; float sg = t >= 0 ? 1.f : -1.f; nf[0] = L[0] * sg; nf[1] = L[1] * sg; nf[2] = L[2] * sg; }
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float L[3];
      cross3(L, a1[i], a2[j]);
      float len = sqrtf(dot3(L, L));
      if (len < 1e-6f) continue;
      float il = 1.0f / len;
      L[0] *= il; L[1] *= il; L[2] *= il;
      float t = dot3(pp, L), ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) { ra += s1[k] * fabsf(dot3(a1[k], L)); rb += s2[k] * fabsf(dot3(a2[k], L)); }
      float sep = fabsf(t) - ra - rb;
      if (sep > margin) return 0;
      if (sep > beste) { beste = sep; codee = 3 * i + j; float sg = t >= 0 ? 1.f : -1.f; ne[0] = L[0] * sg; ne[1] = L[1] * sg; ne[2] = L[2] * sg; }
    }
  if (codee >= 0 && beste > bestf + 0.05f * fabsf(bestf) + 1e-9f) {
    int i = codee / 3, j = codee % 3;
    float c1[3] = {A.pos[0], A.pos[1], A.pos[2]}, c2[3] = {B.pos[0], B.pos[1], B.pos[2]};
    for (int k = 0; k < 3; k++) {
      if (k != i) { float sg = (dot3(ne, a1[k]) > 0 ? 1.f : -1.f) * s1[k]; c1[0] += a1[k][0] * sg; c1[1] += a1[k][1] * sg; c1[2] += a1[k][2] * sg; }
      if (k != j) { float sg = (dot3(ne, a2[k]) > 0 ? -1.f : 1.f) * s2[k]; c2[0] += a2[k][0] * sg; c2[1] += a2[k][1] * sg; c2[2] += a2[k][2] * sg; }
    }
    float r[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]}, bb = dot3(a1[i], a2[j]), den = 1 - bb * bb;
    float d1 = dot3(r, a1[i]), d2 = dot3(r, a2[j]);
    float sA = den > 1e-12f ? (d1 - bb * d2) / den : 0.f, tB = den > 1e-12f ? (bb * d1 - d2) / den : 0.f;
    NP_EMIT(0, ne[0], ne[1], ne[2], beste, 0.5f * (c1[0] + a1[i][0] * sA + c2[0] + a2[j][0] * tB),
            0.5f * (c1[1] + a1[i][1] * sA + c2[1] + a2[j][1] * tB), 0.5f * (c1[2] + a1[i][2] * sA + c2[2] + a2[j][2] * tB));
    return 1;
  }
  bool refis2 = codef >= 3;
  int ax = codef % 3;
  const float *pr = refis2 ? B.pos : A.pos, *sr = refis2 ? s2 : s1, *pc = refis2 ? A.pos : B.pos, *sc = refis2 ? s1 : s2;
  float(*ar)[3] = refis2 ? a2 : a1;
  float(*ac)[3] = refis2 ? a1 : a2;
  float nr[3] = {refis2 ? -nf[0] : nf[0], refis2 ? -nf[1] : nf[1], refis2 ? -nf[2] : nf[2]};
  int jc = 0; float bd = -1;
  for (int k = 0; k < 3; k++) { float t = fabsf(dot3(ac[k], nr)); if (t > bd) { bd = t; jc = k; } }
  float sgn = dot3(ac[jc], nr) > 0 ? -1.f : 1.f;
  float fc[3] = {pc[0] + ac[jc][0] * sgn * sc[jc], pc[1] + ac[jc][1] * sgn * sc[jc], pc[2] + ac[jc][2] * sgn * sc[jc]};
  int k1 = (jc + 1) % 3, k2 = (jc + 2) % 3, u = (ax + 1) % 3, v = (ax + 2) % 3;
  float poly[16][2], tmp[16][2], hgt[4];
  for (int c = 0; c < 4; c++) {
    float e1 = (c == 0 || c == 3) ? -sc[k1] : sc[k1], e2 = c < 2 ? -sc[k2] : sc[k2];
    float rel[3];
    for (int k = 0; k < 3; k++) rel[k] = fc[k] + ac[k1][k] * e1 + ac[k2][k] * e2 - pr[k];
    poly[c][0] = dot3(rel, ar[u]); poly[c][1] = dot3(rel, ar[v]); hgt[c] = dot3(rel, nr);
  }
  float e1x = poly[1][0] - poly[0][0], e1y = poly[1][1] - poly[0][1], e2x = poly[3][0] - poly[0][0], e2y = poly[3][1] - poly[0][1];
  float dh1 = hgt[1] - hgt[0], dh2 = hgt[3] - hgt[0], det = e1x * e2y - e1y * e2x;
  if (fabsf(det) < 1e-14f) return 0;
  float gu = (dh1 * e2y - dh2 * e1y) / det, gv = (e1x * dh2 - e2x * dh1) / det;
  float h0 = hgt[0] - gu * poly[0][0] - gv * poly[0][1];
  int n = 4;
  for (int pass = 0; pass < 4 && n > 0; pass++) {
    int axis = pass >> 1;
    float sign = (pass & 1) ? -1.f : 1.f, lim = axis ? sr[v] : sr[u];
    int no = 0;
    for (int i = 0; i < n; i++) {
      const float *pa = poly[i], *pb = poly[(i + 1) % n];
      float da = sign * pa[axis] - lim, db = sign * pb[axis] - lim;
      if (da <= 0) { tmp[no][0] = pa[0]; tmp[no][1] = pa[1]; no++; }
      if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
        float t = da / (da - db);
        tmp[no][0] = pa[0] + t * (pb[0] - pa[0]); tmp[no][1] = pa[1] + t * (pb[1] - pa[1]); no++;
      }
    }
    n = min(no, 8);
    for (int i = 0; i < n; i++) { poly[i][0] = tmp[i][0]; poly[i][1] = tmp[i][1]; }
  }
  int cnt = 0;
  for (int c = 0; c < n && cnt < NP_MAXC; c++) {
    float h = h0 + gu * poly[c][0] + gv * poly[c][1], dist = h - sr[ax];
    if (dist > margin) continue;
    float hh = h - 0.5f * dist;
    NP_EMIT(cnt, nf[0], nf[1], nf[2], dist, pr[0] + ar[u][0] * poly[c][0] + ar[v][0] * poly[c][1] + nr[0] * hh,
            pr[1] + ar[u][1] * poly[c][0] + ar[v][1] * poly[c][1] + nr[1] * hh, pr[2] + ar[u][2] * poly[c][0] + ar[v][2] * poly[c][1] + nr[2] * hh);
    cnt++;
  }
  return cnt;
}

// analytic plane-vs-primitive / box routines and MPR (+ multiccd) for the rest; fills up to NP_MAXC contacts.
// mtol > 0: multiccd is on and mtol is the distinct-contact tolerance 1e-3 * min(rbound) [upstream mjc_Convex].
__device__ __forceinline__ void narrow_pair_body(const DevModel& m, Cvx& A, Cvx& B, float margin, float mtol, int so, int lane) {
  const bool w0 = lane == 0;
  int& OC = reinterpret_cast<int&>(smem[so]);
  if (w0) OC = 0;
  if (A.type == GEOM_PLANE) {
    float n[3] = {A.mat[2], A.mat[5], A.mat[8]};
    float dif[3] = {B.pos[0] - A.pos[0], B.pos[1] - A.pos[1], B.pos[2] - A.pos[2]};
    if (B.type == GEOM_SPHERE) {
      float r = B.size[0], dist = dot3(dif, n) - r;
      if (dist > margin) return;
      float t = r + 0.5f * dist;
      NP_EMIT(0, n[0], n[1], n[2], dist, B.pos[0] - n[0] * t, B.pos[1] - n[1] * t, B.pos[2] - n[2] * t);
      if (w0) OC = 1;
    } else if (B.type == GEOM_CYLINDER) {
      float axis[3] = {B.mat[2], B.mat[5], B.mat[8]}, vec[3];
      float r = B.size[0], h = B.size[1];
      float prjaxis = dot3(n, axis);
      if (prjaxis > 0) { axis[0] = -axis[0]; axis[1] = -axis[1]; axis[2] = -axis[2]; prjaxis = -prjaxis; }
      float dist0 = dot3(dif, n);
      for (int k = 0; k < 3; k++) vec[k] = axis[k] * prjaxis - n[k];
      float len = sqrtf(dot3(vec, vec));
      if (len < 1e-6f) { vec[0] = B.mat[0] * r; vec[1] = B.mat[3] * r; vec[2] = B.mat[6] * r; }
      else { float s = r / len; vec[0] *= s; vec[1] *= s; vec[2] *= s; }
      float prjvec = dot3(vec, n);
      axis[0] *= h; axis[1] *= h; axis[2] *= h; prjaxis *= h;
      float dist = dist0 + prjaxis + prjvec;
      if (dist > margin) return;
      int c = 0;
      NP_EMIT(c, n[0], n[1], n[2], dist, B.pos[0] + vec[0] + axis[0] - n[0] * dist * 0.5f, B.pos[1] + vec[1] + axis[1] - n[1] * dist * 0.5f,
              B.pos[2] + vec[2] + axis[2] - n[2] * dist * 0.5f);
      c++;
      dist = dist0 - prjaxis + prjvec;
      if (dist <= margin) {
        NP_EMIT(c, n[0], n[1], n[2], dist, B.pos[0] + vec[0] - axis[0] - n[0] * dist * 0.5f, B.pos[1] + vec[1] - axis[1] - n[1] * dist * 0.5f,
                B.pos[2] + vec[2] - axis[2] - n[2] * dist * 0.5f);
        c++;
      }
      float prjvec1 = -prjvec * 0.5f;
      dist = dist0 + prjaxis + prjvec1;
      if (dist <= margin) {
        float vec1[3];
        cross3(vec1, vec, axis); normalize3(vec1);
        float s = r * 0.8660254037844386f;
        vec1[0] *= s; vec1[1] *= s; vec1[2] *= s;
        NP_EMIT(c, n[0], n[1], n[2], dist, B.pos[0] + vec1[0] + axis[0] - vec[0] * 0.5f - n[0] * dist * 0.5f,
                B.pos[1] + vec1[1] + axis[1] - vec[1] * 0.5f - n[1] * dist * 0.5f, B.pos[2] + vec1[2] + axis[2] - vec[2] * 0.5f - n[2] * dist * 0.5f);
        c++;
        NP_EMIT(c, n[0], n[1], n[2], dist, B.pos[0] - vec1[0] + axis[0] - vec[0] * 0.5f - n[0] * dist * 0.5f,
                B.pos[1] - vec1[1] + axis[1] - vec[1] * 0.5f - n[1] * dist * 0.5f, B.pos[2] - vec1[2] + axis[2] - vec[2] * 0.5f - n[2] * dist * 0.5f);
        c++;
      }
      if (w0) OC = c;
    } else if (B.type == GEOM_BOX) {
      float dist = dot3(dif, n);
      int c = 0;
      for (int i = 0; i < 8 && c < 4; i++) {
        float l[3] = {(i & 1) ? B.size[0] : -B.size[0], (i & 2) ? B.size[1] : -B.size[1], (i & 4) ? B.size[2] : -B.size[2]};
        float vec[3];
        mat_vec(vec, B.mat, l);
        float ldist = dot3(n, vec);
        if (dist + ldist > margin || ldist > 0) continue;
        float cd = dist + ldist;
        NP_EMIT(c, n[0], n[1], n[2], cd, B.pos[0] + vec[0] - n[0] * cd * 0.5f, B.pos[1] + vec[1] - n[1] * cd * 0.5f, B.pos[2] + vec[2] - n[2] * cd * 0.5f);
        c++;
      }
      if (w0) OC = c;
    } else if (B.type == GEOM_MESH) {
      // [upstream mjc_PlaneConvex] support vertex + up to three of its hull neighbours within the margin
      float nd[3] = {-n[0], -n[1], -n[2]}, s[3];
      int v0 = support(B, nd, s, lane);
      float d3[3] = {s[0] - A.pos[0], s[1] - A.pos[1], s[2] - A.pos[2]};
      float dist = dot3(d3, n);
      if (dist > margin) return;
      NP_EMIT(0, n[0], n[1], n[2], dist, s[0] - n[0] * 0.5f * dist, s[1] - n[1] * 0.5f * dist, s[2] - n[2] * 0.5f * dist);
      int c = 1;
      if (m.hull_edgeadr) {
        int e0 = m.hull_edgeadr[B.hadr + v0], e1 = m.hull_edgeadr[B.hadr + v0 + 1];
        for (int e = e0; e < e1 && c < 4; e++) {
          float4 vl = B.verts[m.hull_edge[e]];
          float l[3] = {vl.x, vl.y, vl.z}, w[3];
          mat_vec(w, B.mat, l);
          w[0] += B.pos[0]; w[1] += B.pos[1]; w[2] += B.pos[2];
          float dd = (w[0] - A.pos[0]) * n[0] + (w[1] - A.pos[1]) * n[1] + (w[2] - A.pos[2]) * n[2];
          if (dd > margin) continue;
          NP_EMIT(c, n[0], n[1], n[2], dd, w[0] - n[0] * 0.5f * dd, w[1] - n[1] * 0.5f * dd, w[2] - n[2] * 0.5f * dd);
          c++;
        }
      }
      if (w0) OC = c;
    }
  } else if (A.type == GEOM_BOX && B.type == GEOM_BOX) {
    int c = box_box(A, B, margin, so, lane);
    if (w0) OC = c;
  } else if (A.type == GEOM_SPHERE && B.type == GEOM_BOX) {
    // [upstream mjc_SphereBox] closest point of the box to the sphere centre
    float dif[3] = {A.pos[0] - B.pos[0], A.pos[1] - B.pos[1], A.pos[2] - B.pos[2]}, cl[3], q[3], nl[3] = {0, 0, 0}, dist;
    matT_vec(cl, B.mat, dif);
    bool inside = true;
    for (int k = 0; k < 3; k++) { q[k] = fminf(fmaxf(cl[k], -B.size[k]), B.size[k]); if (q[k] != cl[k]) inside = false; }
    float r = A.size[0];
    if (!inside) {
      float dl[3] = {cl[0] - q[0], cl[1] - q[1], cl[2] - q[2]}, len = sqrtf(dot3(dl, dl));
      dist = len - r;
      if (dist > margin) return;
      nl[0] = dl[0] / len; nl[1] = dl[1] / len; nl[2] = dl[2] / len;
    } else {
      int best = 0; float bd = CUDART_INF_F;
      for (int k = 0; k < 3; k++) { float t = B.size[k] - fabsf(cl[k]); if (t < bd) { bd = t; best = k; } }
      for (int k = 0; k < 3; k++) if (k == best) { nl[k] = cl[k] >= 0 ? 1.f : -1.f; q[k] = nl[k] * B.size[k]; }
      dist = -bd - r;
    }
    float nw[3], qw[3];
    mat_vec(nw, B.mat, nl); mat_vec(qw, B.mat, q);
    NP_EMIT(0, -nw[0], -nw[1], -nw[2], dist, qw[0] + B.pos[0] + nw[0] * 0.5f * dist, qw[1] + B.pos[1] + nw[1] * 0.5f * dist,
            qw[2] + B.pos[2] + nw[2] * 0.5f * dist);
    if (w0) OC = 1;
  } else {
    // [upstream mjc_Convex] one MPR query, then (multiccd, stretch.xml:8) four more with both geoms counter-rotated
    // by +-1e-3 rad about the tangent axes of the first contact; a contact is kept when its position is farther
    // than mtol from all contacts found so far.  One call site of mpr_penetration (code size).
    bool multi = mtol > 0 && A.type != GEOM_SPHERE && A.type != GEOM_ELLIPSOID && B.type != GEOM_SPHERE && B.type != GEOM_ELLIPSOID;
    float* P0 = smem + so + 72;   // unperturbed poses: A.pos, A.mat, B.pos, B.mat
    float frame[9];
    int cnt = 0;
    for (int q = 0; q < (multi ? 5 : 1); q++) {
      if (q == 1) {
        __syncwarp();
        if (w0) {
          for (int k = 0; k < 3; k++) { P0[k] = A.pos[k]; P0[12 + k] = B.pos[k]; }
          for (int k = 0; k < 9; k++) { P0[3 + k] = A.mat[k]; P0[15 + k] = B.mat[k]; }
        }
        __syncwarp();
      }
      if (q >= 1) {
        const float* ax = frame + 3 + 3 * ((q - 1) >> 1);
        const float sn = ((q - 1) & 1) ? 4.99999979e-4f : -4.99999979e-4f, cs = 0.999999875f;   // sin, cos of 1e-3 / 2
        float qr[4] = {cs, ax[0] * sn, ax[1] * sn, ax[2] * sn}, R[9];
        quat2mat(R, qr);
        const float* org = smem + so + 4 + 4;   // position of the first contact
        float o3[3] = {org[0], org[1], org[2]};
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) {
            A.mat[3 * r + c] = R[3 * r] * P0[3 + c] + R[3 * r + 1] * P0[6 + c] + R[3 * r + 2] * P0[9 + c];
            B.mat[3 * r + c] = R[r] * P0[15 + c] + R[3 + r] * P0[18 + c] + R[6 + r] * P0[21 + c];
          }
        float ra[3] = {P0[0] - o3[0], P0[1] - o3[1], P0[2] - o3[2]}, rb[3] = {P0[12] - o3[0], P0[13] - o3[1], P0[14] - o3[2]}, t[3];
        mat_vec(t, R, ra); A.pos[0] = o3[0] + t[0]; A.pos[1] = o3[1] + t[1]; A.pos[2] = o3[2] + t[2];
        matT_vec(t, R, rb); B.pos[0] = o3[0] + t[0]; B.pos[1] = o3[1] + t[1]; B.pos[2] = o3[2] + t[2];
      }
      float depth, dir[3], pos[3];
      bool hit = mpr_penetration(A, B, &depth, dir, pos, lane) && dot3(dir, dir) >= 0.5f;
      if (q == 0) {
        if (!hit) return;
        frame[0] = dir[0]; frame[1] = dir[1]; frame[2] = dir[2];
        float* y = frame + 3;
        y[0] = 0; y[1] = 1; y[2] = 0;
        if (dir[1] > 0.5f || dir[1] < -0.5f) { y[1] = 0; y[2] = 1; }
        float d = dot3(dir, y);
        y[0] -= dir[0] * d; y[1] -= dir[1] * d; y[2] -= dir[2] * d;
        normalize3(y);
        cross3(frame + 6, dir, y);
      } else {
        if (!hit) continue;
        bool distinct = true;
        for (int k = 0; k < cnt; k++) {
          const float* pk = smem + so + 4 + 8 * k + 4;
          float dx = pk[0] - pos[0], dy = pk[1] - pos[1], dz = pk[2] - pos[2];
          if (dx * dx + dy * dy + dz * dz < mtol * mtol) distinct = false;
        }
        if (!distinct) continue;
      }
      NP_EMIT(cnt, dir[0], dir[1], dir[2], margin - depth, pos[0], pos[1], pos[2]);
      cnt++;
      if (w0) OC = cnt;
      __syncwarp();
    }
  }
}

// Single noinline body of the narrowphase; the operands are copied into registers once (they arrive through
// local memory) and everything below (MPR, support scans) is inlined so that they stay there.
__device__ __noinline__ void narrow_pair(const DevModel& m, const Cvx& A_, const Cvx& B_, float margin, float mtol, int so, int lane) {
  Cvx A = A_, B = B_;
  narrow_pair_body(m, A, B, margin, mtol, so, lane);
  __syncwarp();
}

// Broadphase (runs in the smooth kernel): world positions of the collision geoms, then bounding spheres and
// oriented bounding boxes over the compile-time pair list.  Passing pairs keep the reference's pair order:
// slot s of env e holds the s-th passing pair (a.slot_pair) and becomes one work item of the narrowphase kernel.
__device__ __forceinline__ void broadphase(const DevModel& m, const StepArgs& a, float* S, int env, int& flags, int lane) {
  const EnvLayout& o = m.L;
  float* gpos = S + o.gpos;
  _Pragma("unroll 1") for (int g = lane; g < m.ncgeom; g += 32) {
    int b = PKI(cg_bodyid)[g];
    float t[3];
    mat_vec(t, S + o.xmat + 9 * b, PKF(cg_pos) + 3 * g);
    gpos[3 * g] = S[o.xpos + 3 * b] + t[0]; gpos[3 * g + 1] = S[o.xpos + 3 * b + 1] + t[1]; gpos[3 * g + 2] = S[o.xpos + 3 * b + 2] + t[2];
  }
  __syncwarp();
  int npass = 0;
  for (int base = 0; base < m.npair; base += 32) {
    int p = base + lane;
    bool pass = false;
    if (p < m.npair) {
      int pc = PKI(pair_cg)[p], c1 = pc & 0xffff, c2 = pc >> 16;
      float margin = m.max_margin;
      float dif[3] = {gpos[3 * c2] - gpos[3 * c1], gpos[3 * c2 + 1] - gpos[3 * c1 + 1], gpos[3 * c2 + 2] - gpos[3 * c1 + 2]};
      // stage 1: bounding spheres; stage 2: oriented bounding boxes (6 face axes / lowest box corner
      // against the plane).  Both are conservative: a rejected pair cannot produce a contact.
      float q2[4], R2[9];
      const float *ab2 = PKF(cg_aabb) + 6 * c2;
      if (PKI(cg_type)[c1] == GEOM_PLANE) {
        int b = PKI(cg_bodyid)[c1];
        float q[4], R[9];
        quat_mul(q, S + o.xquat + 4 * b, PKF(cg_quat) + 4 * c1);
        quat2mat(R, q);
        float n[3] = {R[2], R[5], R[8]};
        pass = dot3(dif, n) <= PKF(cg_rbound)[c2] + margin;
        if (pass) {
          quat_mul(q2, S + o.xquat + 4 * PKI(cg_bodyid)[c2], PKF(cg_quat) + 4 * c2);
          quat_normalize(q2);
          quat2mat(R2, q2);
          float cw[3];
          mat_vec(cw, R2, ab2);
          float low = dot3(dif, n) + dot3(cw, n);
          for (int k = 0; k < 3; k++) low -= fabsf(R2[k] * n[0] + R2[3 + k] * n[1] + R2[6 + k] * n[2]) * ab2[3 + k];
          pass = low <= margin;
        }
      } else {
        float bound = PKF(cg_rbound)[c1] + PKF(cg_rbound)[c2] + margin;
        pass = dot3(dif, dif) <= bound * bound;
        if (pass) {
          float q1[4], R1[9], c1w[3], c2w[3];
          const float* ab1 = PKF(cg_aabb) + 6 * c1;
          quat_mul(q1, S + o.xquat + 4 * PKI(cg_bodyid)[c1], PKF(cg_quat) + 4 * c1);
          quat_normalize(q1);
          quat2mat(R1, q1);
          quat_mul(q2, S + o.xquat + 4 * PKI(cg_bodyid)[c2], PKF(cg_quat) + 4 * c2);
          quat_normalize(q2);
          quat2mat(R2, q2);
          mat_vec(c1w, R1, ab1);
          mat_vec(c2w, R2, ab2);
          float t[3] = {dif[0] + c2w[0] - c1w[0], dif[1] + c2w[1] - c1w[1], dif[2] + c2w[2] - c1w[2]};
          for (int i = 0; i < 3 && pass; i++) {
            float a1[3] = {R1[i], R1[3 + i], R1[6 + i]}, a2[3] = {R2[i], R2[3 + i], R2[6 + i]};
            float r1 = ab1[3 + i] + margin, r2 = ab2[3 + i] + margin;
            for (int k = 0; k < 3; k++) {
              r1 += fabsf(R2[k] * a1[0] + R2[3 + k] * a1[1] + R2[6 + k] * a1[2]) * ab2[3 + k];
              r2 += fabsf(R1[k] * a2[0] + R1[3 + k] * a2[1] + R1[6 + k] * a2[2]) * ab1[3 + k];
            }
            if (fabsf(dot3(t, a1)) > r1 || fabsf(dot3(t, a2)) > r2) pass = false;
          }
          if (pass) {
            // the nine edge-edge axes L = a1_i x a2_j of the separating-axis test (in box 1's frame:
            // C = R1^T R2, T = R1^T t); the margin inflates the projection by margin * |L| <= margin
            float C[9], AC[9], T[3];
            for (int i = 0; i < 3; i++) {
              T[i] = R1[i] * t[0] + R1[3 + i] * t[1] + R1[6 + i] * t[2];
              for (int j = 0; j < 3; j++) {
                C[3 * i + j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
                AC[3 * i + j] = fabsf(C[3 * i + j]) + 1e-6f;
              }
            }
            const float *h1 = ab1 + 3, *h2 = ab2 + 3;
            for (int i = 0; i < 3 && pass; i++) {
              int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
              for (int j = 0; j < 3; j++) {
                int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                float ra = h1[i1] * AC[3 * i2 + j] + h1[i2] * AC[3 * i1 + j];
                float rb = h2[j1] * AC[3 * i + j2] + h2[j2] * AC[3 * i + j1];
                if (fabsf(T[i2] * C[3 * i1 + j] - T[i1] * C[3 * i2 + j]) > ra + rb + margin) pass = false;
              }
            }
          }
        }
      }
    }
    unsigned mask = __ballot_sync(FULL, pass);
    int cnt = __popc(mask);
    if (cnt) {
      int slot = npass + __popc(mask & ((1u << lane) - 1));
      int qbase = 0;
      int room = max(0, min(cnt, a.maxslot - npass));
      if (cnt > room) flags |= 2;
      if (lane == 0 && room) qbase = atomicAdd(a.item_count, room);
      qbase = __shfl_sync(FULL, qbase, 0);
      if (pass && slot < a.maxslot) {
        int item = env * a.maxslot + slot;
        a.slot_pair[item] = p;
        a.items[qbase + slot - npass] = item;
      }
      npass += room;
    }
  }
  if (lane == 0) a.npass[env] = npass;
  __syncwarp();
}

// ----------------------------------------------------------------------------- S3/S5: velocity, bias, actuation
__device__ __forceinline__ void velocity_stage(const DevModel& m, float* S, int lane) {
  const EnvLayout& o = m.L;
  const float *cdof = S + o.cdof, *qvel = S + o.qvel, *cinert = S + o.cinert;
  float *cvel = S + o.cvel, *cdofdot = S + o.cdofdot, *cacc = S + o.cacc, *cfrc = S + o.cfrc;
  if (lane < 6) { cvel[lane] = 0; cfrc[lane] = 0; cacc[lane] = (lane < 3) ? 0.f : -m.gravity[lane - 3]; }
  __syncwarp();
  for (int lv = 0; lv < m.nlevel; lv++) {
    for (int idx = PKI(lvl_adr)[lv] + lane; idx < PKI(lvl_adr)[lv + 1]; idx += 32) {
      int b = PKI(lvl_body)[idx], p = PKI(body_parentid)[b];
      float cv[6], ca[6];
#pragma unroll
      for (int k = 0; k < 6; k++) { cv[k] = cvel[6 * p + k]; ca[k] = cacc[6 * p + k]; }
      int ja = PKI(body_jntadr)[b];
      for (int j = ja; j < ja + PKI(body_jntnum)[b]; j++) {
        int da = PKI(jnt_dofadr)[j];
        if (PKI(jnt_type)[j] == JNT_FREE) {
          for (int k = 0; k < 3; k++)
            for (int c = 0; c < 6; c++) { cdofdot[6 * (da + k) + c] = 0; cv[c] += cdof[6 * (da + k) + c] * qvel[da + k]; }
          for (int k = 3; k < 6; k++) {
            float cd[6];
            cross_motion(cd, cv, cdof + 6 * (da + k));
            for (int c = 0; c < 6; c++) { cdofdot[6 * (da + k) + c] = cd[c]; ca[c] += cd[c] * qvel[da + k]; }
          }
          for (int k = 3; k < 6; k++)
            for (int c = 0; c < 6; c++) cv[c] += cdof[6 * (da + k) + c] * qvel[da + k];
        } else {
          float cd[6];
          cross_motion(cd, cv, cdof + 6 * da);
          for (int c = 0; c < 6; c++) { cdofdot[6 * da + c] = cd[c]; ca[c] += cd[c] * qvel[da]; cv[c] += cdof[6 * da + c] * qvel[da]; }
        }
      }
#pragma unroll
      for (int k = 0; k < 6; k++) { cvel[6 * b + k] = cv[k]; cacc[6 * b + k] = ca[k]; }
      // body force; gravity compensation folded in by scaling this body's gravity term
      float gc = PKF(body_gravcomp)[b];
      float cae[6] = {ca[0], ca[1], ca[2], ca[3] + gc * m.gravity[0], ca[4] + gc * m.gravity[1], ca[5] + gc * m.gravity[2]};
      float f[6], t1[6], t2[6];
      mul_inert_vec(f, cinert + 10 * b, cae);
      mul_inert_vec(t1, cinert + 10 * b, cv);
      cross_force(t2, cv, t1);
#pragma unroll
      for (int k = 0; k < 6; k++) cfrc[6 * b + k] = f[k] + t2[k];
    }
    __syncwarp();
  }
  for (int lv = m.nlevel - 2; lv >= 0; lv--) {
    for (int idx = PKI(lvl_adr)[lv] + lane; idx < PKI(lvl_adr)[lv + 1]; idx += 32) {
      int b = PKI(lvl_body)[idx];
      float acc[6];
#pragma unroll
      for (int k = 0; k < 6; k++) acc[k] = cfrc[6 * b + k];
      for (int c = PKI(child_adr)[b]; c < PKI(child_adr)[b + 1]; c++) {
        const float* cc = cfrc + 6 * PKI(child_list)[c];
#pragma unroll
        for (int k = 0; k < 6; k++) acc[k] += cc[k];
      }
#pragma unroll
      for (int k = 0; k < 6; k++) cfrc[6 * b + k] = acc[k];
    }
    __syncwarp();
  }
}

// qfrc_smooth = passive(springs, dampers) - bias(+gravcomp) + actuator ; also actuator length/velocity/force
__device__ __forceinline__ void smooth_forces(const DevModel& m, float* S, int lane) {
  const EnvLayout& o = m.L;
  const float *qpos = S + o.qpos, *qvel = S + o.qvel, *cdof = S + o.cdof, *cfrc = S + o.cfrc, *ctrl = S + o.ctrl;
  float *actforce = S + o.actforce, *actlen = S + o.actlen, *actvel = S + o.actvel, *qs = S + o.qfrc_smooth;
  int nv = m.nv;
  _Pragma("unroll 1") for (int a = lane; a < m.nu; a += 32) {
    float len = 0, vel = 0;
    const float* mom = PKF(act_moment) + a * nv;
    for (int d = 0; d < nv; d++) {
      float mm = mom[d];
      if (mm != 0) { len += mm * qpos[PKI(dof_qposadr)[d]]; vel += mm * qvel[d]; }
    }
    float c = ctrl[a];
    if (PKI(actuator_ctrllimited)[a]) c = fminf(fmaxf(c, PKF(actuator_ctrlrange)[2 * a]), PKF(actuator_ctrlrange)[2 * a + 1]);
    const float *gp = PKF(actuator_gainprm) + 3 * a, *bp = PKF(actuator_biasprm) + 3 * a;
    float f = gp[0] * c + bp[0] + bp[1] * len + bp[2] * vel;
    if (PKI(actuator_forcelimited)[a]) f = fminf(fmaxf(f, PKF(actuator_forcerange)[2 * a]), PKF(actuator_forcerange)[2 * a + 1]);
    actlen[a] = len; actvel[a] = vel; actforce[a] = f;
  }
  __syncwarp();
  _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) {
    const float *c = cdof + 6 * i, *f = cfrc + 6 * PKI(dof_bodyid)[i];
    float bias = c[0] * f[0] + c[1] * f[1] + c[2] * f[2] + c[3] * f[3] + c[4] * f[4] + c[5] * f[5];
    float q = -PKF(dof_damping)[i] * qvel[i] - bias;
    int j = PKI(dof_jntid)[i];
    float k = PKF(jnt_stiffness)[j];
    if (k != 0 && PKI(jnt_type)[j] >= JNT_SLIDE) { int qa = PKI(jnt_qposadr)[j]; q -= k * (qpos[qa] - PKF(qpos_spring)[qa]); }
    for (int a = 0; a < m.nu; a++) q += PKF(act_moment)[a * nv + i] * actforce[a];
    qs[i] = q;
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------- S1m: constraints
__device__ __forceinline__ float impedance(const float* solimp, float pos, float margin) {
  float dmin = fminf(fmaxf(solimp[0], MINIMP), MAXIMP), dmax = fminf(fmaxf(solimp[1], MINIMP), MAXIMP);
  float width = fmaxf(solimp[2], 0.f), mid = fminf(fmaxf(solimp[3], MINIMP), MAXIMP), power = fmaxf(solimp[4], 1.f);
  if (dmin == dmax || width <= MINVAL) return 0.5f * (dmin + dmax);
  float x = fabsf((pos - margin) / width);
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  float y;
  if (power == 1) y = x;
  else if (x <= mid) y = powf(x, power) / powf(mid, power - 1);
  else y = 1 - powf(1 - x, power) / powf(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}

// R and reference acceleration of one row from solref/solimp [upstream mj_makeImpedance, mj_referenceConstraint]
__device__ __noinline__ void row_params(float timestep, const float* solref, const float* solimp, float pos, float margin,
                                        float diag, float vel, float* R, float* aref) {
  float imp = impedance(solimp, pos, margin);
  *R = fmaxf(MINVAL, (1 - imp) / imp * diag);
  float K, B, dmax = fminf(fmaxf(solimp[1], MINIMP), MAXIMP);
  if (solref[0] > 0) {
    float tc = fmaxf(solref[0], 2 * timestep), dr = solref[1];
    K = 1 / fmaxf(MINVAL, dmax * dmax * tc * tc * dr * dr);
    B = 2 / fmaxf(MINVAL, dmax * tc);
  } else {
    K = -solref[0] / fmaxf(MINVAL, dmax * dmax);
    B = -solref[1] / fmaxf(MINVAL, dmax);
  }
  *aref = -B * vel - K * imp * (pos - margin);
}

// Build constraint rows in the reference order: equality, friction loss, limits, contacts.
__device__ __forceinline__ void make_constraints(const DevModel& m, float* S, const float* __restrict__ G, int& ncon, int& ns_out,
                                                 int& nefc_out, int& flags, int lane) {
  const EnvLayout& o = m.L;
  const float *qpos = S + o.qpos, *qvel = S + o.qvel;
  float *sc1 = S + o.s_c1, *sc2 = S + o.s_c2, *eR = S + o.e_R, *eD = S + o.e_D, *earef = S + o.e_aref, *efl = S + o.e_floss;
  int *sd1 = (int*)(S + o.s_d1), *sd2 = (int*)(S + o.s_d2), *einfo = (int*)(S + o.e_info);
  int row0 = 0;
  {
    int nact = 0;
    for (int base = 0; base < m.neq; base += 32) {
      int e = base + lane;
      bool act = e < m.neq && PKI(eq_active0)[e];
      unsigned mask = __ballot_sync(FULL, act);
      if (act) {
        int row = row0 + nact + __popc(mask & ((1u << lane) - 1));
        int j1 = PKI(eq_obj1id)[e], j2 = PKI(eq_obj2id)[e];
        const float* c = PKF(eq_data) + 5 * e;
        int q1 = PKI(jnt_qposadr)[j1], d1 = PKI(jnt_dofadr)[j1];
        float pos, diag = PKF(dof_invweight0)[d1], vel, c2 = 0;
        int d2 = -1;
        if (j2 >= 0) {
          int q2 = PKI(jnt_qposadr)[j2];
          d2 = PKI(jnt_dofadr)[j2];
          float dif = qpos[q2] - PKF(qpos0)[q2];
          float poly = c[0] + dif * (c[1] + dif * (c[2] + dif * (c[3] + dif * c[4])));
          c2 = -(c[1] + dif * (2 * c[2] + dif * (3 * c[3] + dif * 4 * c[4])));
          pos = qpos[q1] - PKF(qpos0)[q1] - poly;
          diag += PKF(dof_invweight0)[d2];
          vel = qvel[d1] + c2 * qvel[d2];
        } else {
          pos = qpos[q1] - PKF(qpos0)[q1] - c[0];
          vel = qvel[d1];
        }
        float R, aref;
        row_params(m.timestep, PKF(eq_solref) + 2 * e, PKF(eq_solimp) + 5 * e, pos, 0.f, diag, vel, &R, &aref);
        sd1[row] = d1; sc1[row] = 1.f; sd2[row] = d2; sc2[row] = c2;
        eR[row] = R; eD[row] = 1.f / R; earef[row] = aref; efl[row] = 0; einfo[row] = CNSTR_EQUALITY | (e << 8);
      }
      nact += __popc(mask);
    }
    row0 += nact;
  }
  _Pragma("unroll 1") for (int k = lane; k < m.nfloss; k += 32) {
    int d = PKI(floss_list)[k], row = row0 + k;
    float R, aref;
    row_params(m.timestep, PKF(dof_solref) + 2 * d, PKF(dof_solimp) + 5 * d, 0.f, 0.f, PKF(dof_invweight0)[d], qvel[d], &R, &aref);
    sd1[row] = d; sc1[row] = 1.f; sd2[row] = -1; sc2[row] = 0;
    eR[row] = R; eD[row] = 1.f / R; earef[row] = aref; efl[row] = PKF(dof_frictionloss)[d]; einfo[row] = CNSTR_FRICTION | (d << 8);
  }
  row0 += m.nfloss;
  for (int base = 0; base < m.nlimited; base += 32) {
    int k = base + lane;
    bool lo = false, hi = false;
    int j = 0, d = 0;
    float dlo = 0, dhi = 0, margin = 0;
    if (k < m.nlimited) {
      j = PKI(limited_list)[k]; d = PKI(jnt_dofadr)[j];
      float q = qpos[PKI(jnt_qposadr)[j]];
      margin = PKF(jnt_margin)[j];
      dlo = q - PKF(jnt_range)[2 * j]; dhi = PKF(jnt_range)[2 * j + 1] - q;
      lo = dlo < margin; hi = dhi < margin;
    }
    unsigned mlo = __ballot_sync(FULL, lo), mhi = __ballot_sync(FULL, hi);
    unsigned below = (1u << lane) - 1;
    int r = row0 + __popc(mlo & below) + __popc(mhi & below);
    if (lo || hi) {
      for (int side = 0; side < 2; side++) {
        if (!(side ? hi : lo)) continue;
        float R, aref, sg = side ? -1.f : 1.f;
        row_params(m.timestep, PKF(jnt_solref) + 2 * j, PKF(jnt_solimp) + 5 * j, side ? dhi : dlo, margin, PKF(dof_invweight0)[d],
                   sg * qvel[d], &R, &aref);
        sd1[r] = d; sc1[r] = sg; sd2[r] = -1; sc2[r] = 0;
        eR[r] = R; eD[r] = 1.f / R; earef[r] = aref; efl[r] = 0; einfo[r] = CNSTR_LIMIT | (j << 8);
        r++;
      }
    }
    row0 += __popc(mlo) + __popc(mhi);
  }
  int ns = row0;
  ns_out = ns;
  // contacts: packed Jacobian rows (only the dofs on exactly one of the two bodies' chains), lane = dof
  float* J = S + o.J;
  const float *cdof = S + o.cdof, *xpos = G + o.xpos;   // G: the env's global block (part B)
  int crow = 0, jn = 0, nv = m.nv;
  for (int c = 0; c < ncon; c++) {
    float* con = S + o.con + c * CON_STRIDE;
    int dim = __float_as_int(con[C_DIM]);
    int b1 = __float_as_int(con[C_BODY1]), b2 = __float_as_int(con[C_BODY2]), pair = __float_as_int(con[C_PAIR]);
    const unsigned m2lo = PKI(body_dofmask)[2 * b2], m2hi = PKI(body_dofmask)[2 * b2 + 1];
    const unsigned lo = (unsigned)PKI(body_dofmask)[2 * b1] ^ m2lo, hi = (unsigned)PKI(body_dofmask)[2 * b1 + 1] ^ m2hi;
    const int w = __popc(lo) + __popc(hi);
    if (crow + dim > m.maxcrow || jn + dim * w > m.maxjnz) { flags |= 2; ncon = c; break; }
    float pos[3] = {con[C_POS], con[C_POS + 1], con[C_POS + 2]};
    _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) {
      if (!mask_has(lo, hi, i)) continue;
      float sg = mask_has(m2lo, m2hi, i) ? 1.f : -1.f;
      const float *cd = cdof + 6 * i, *ref = xpos + 3 * PKI(root_list)[PKI(body_rootidx)[PKI(dof_bodyid)[i]]];
      float off[3] = {pos[0] - ref[0], pos[1] - ref[1], pos[2] - ref[2]}, t[3];
      cross3(t, cd, off);
      float jr[3] = {sg * cd[0], sg * cd[1], sg * cd[2]};
      float jp[3] = {sg * (cd[3] + t[0]), sg * (cd[4] + t[1]), sg * (cd[5] + t[2])};
      float* Jc = J + jn + mask_pos(lo, hi, i);
      for (int r = 0; r < dim; r++) {
        const float* ax = con + C_FRAME + 3 * (r < 3 ? r : r - 3);
        Jc[r * w] = (r < 3) ? dot3(ax, jp) : dot3(ax, jr);
      }
    }
    __syncwarp();
    if (lane < dim) {
      int r = lane, row = ns + crow + r;
      const float* Jr = J + jn + r * w;
      float vel = 0;
      { unsigned a = lo, b = hi; int k = 0;
        while (a) { int d = __ffs(a) - 1; a &= a - 1; vel += Jr[k++] * qvel[d]; }
        while (b) { int d = __ffs(b) - 1; b &= b - 1; vel += Jr[k++] * qvel[32 + d]; } }
      float diag = (r < 3) ? (PKF(body_invweight0)[2 * b1] + PKF(body_invweight0)[2 * b2])
                           : (PKF(body_invweight0)[2 * b1 + 1] + PKF(body_invweight0)[2 * b2 + 1]);
      float R, aref;
      const float *solref = m.pair_solref + 2 * pair, *solimp = m.pair_solimp + 5 * pair;
      float inclmargin = m.pair_margin[pair] - m.pair_gap[pair];
      row_params(m.timestep, solref, solimp, r == 0 ? con[C_DIST] : 0.f, r == 0 ? inclmargin : 0.f, diag, vel, &R, &aref);
      eD[row] = R; earef[row] = aref;   // contact rows keep only D = 1 / R (inverted below); R itself is needed for friction-loss rows only
      einfo[row] = ((dim == 1) ? CNSTR_CONTACT_FRICTIONLESS : CNSTR_CONTACT_ELLIPTIC) | (c << 8);
    }
    __syncwarp();
    if (lane == 0) {
      int row = ns + crow;
      con[C_EFC] = __int_as_float(row);
      unsigned cm = 0;
      for (int k = 0; k < 16; k++) if (((k < 8 ? lo >> (4 * k) : hi >> (4 * k - 32)) & 0xFu) != 0) cm |= 1u << k;
      con[C_MASKLO] = __uint_as_float(lo); con[C_MASKHI] = __uint_as_float(hi);
      con[C_JOFS] = __int_as_float(jn); con[C_CMASK] = __uint_as_float(cm);
      if (dim > 1) {
        float R0 = eD[row], R1 = R0 / fmaxf(MINVAL, m.impratio), f0 = con[C_FRICTION];
        eD[row + 1] = R1;
        con[C_MU] = f0 * sqrtf(R1 / R0);
        for (int j = 2; j < dim; j++) { float fj = con[C_FRICTION + j - 1]; eD[row + j] = R1 * f0 * f0 / (fj * fj); }
      }
      for (int j = 0; j < dim; j++) eD[row + j] = 1.f / eD[row + j];
    }
    crow += dim; jn += dim * w;
  }
  __syncwarp();
  nefc_out = ns + crow;
}

// ----------------------------------------------------------------------------- S7: Newton solver
// Every warp of the CTA makes this call (`active` = it owns an env).  With `sync` the Newton loop is
// CTA-uniform: the warps meet before the gradient, the Hessian and the line search of every iteration
// (converged warps only keep the barriers company), which lets them share instruction-cache lines.
template <int TILE>
__device__ __forceinline__ int solve_constraints(const DevModel& m, float* S, int ns, int nefc, int ncon, int lane, bool active, int sync) {
  const EnvLayout& o = m.L;
  int nv = m.nv;
  float *qacc = S + o.qacc, *Ma = S + o.v_Ma, *grad = S + o.v_grad, *search = S + o.v_search, *mv = S + o.v_mv;
  float *jar = S + o.e_jar, *jv = S + o.e_jv, *force = S + o.e_force, *qfc = S + o.qfrc_con;
  const float *qs = S + o.qfrc_smooth, *qas = S + o.qacc_smooth, *warm = S + o.warm, *M = S + o.M, *aref = S + o.e_aref;
  bool done = !active || nefc == 0;
  if (active && nefc == 0) {
    _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) { qacc[i] = qas[i]; qfc[i] = 0; }
    __syncwarp();
  }
  Rows R;
  R.sd1 = (const int*)(S + o.s_d1); R.sd2 = (const int*)(S + o.s_d2); R.info = (int*)(S + o.e_info);
  R.sc1 = S + o.s_c1; R.sc2 = S + o.s_c2; R.J = S + o.J; R.eD = S + o.e_D; R.eR = S + o.e_R; R.efl = S + o.e_floss;
  R.con = S + o.con; R.jar = jar; R.jv = jv; R.force = force; R.ns = ns; R.nefc = nefc; R.ncon = ncon; R.ldj = o.ldj; R.nv = nv;
  float scale = 1.0f / (m.meaninertia * (nv > 1 ? nv : 1));
  float cw, cs;
  float cost = 0;
  if (!done) {
  // warm start vs. unconstrained acceleration: keep the cheaper one
  _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) qacc[i] = warm[i];
  __syncwarp();
  mul_J(R, jar, qacc, lane);
  symv(Ma, M, qacc, nv, o.ldm, lane);
  _Pragma("unroll 1") for (int r = lane; r < nefc; r += 32) jar[r] -= aref[r];
  __syncwarp();
  cw = eval_constraints<true>(R, 0.f, lane).x;
  {
    float gsum = 0;
    _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) gsum += 0.5f * (Ma[i] - qs[i]) * (qacc[i] - qas[i]);
    cw += warp_sum(gsum);
  }
  mul_J(R, jv, qas, lane);
  _Pragma("unroll 1") for (int r = lane; r < nefc; r += 32) { float t = jv[r] - aref[r]; jv[r] = t - jar[r]; }
  __syncwarp();
  cs = eval_constraints<false>(R, 1.0f, lane).x;
  cost = cw;
  if (!(cw <= cs)) {  // also catches NaN warm starts
    _Pragma("unroll 1") for (int r = lane; r < nefc; r += 32) jar[r] += jv[r];
    _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) qacc[i] = qas[i];
    __syncwarp();
    symv(Ma, M, qacc, nv, o.ldm, lane);
    cost = eval_constraints<true>(R, 0.f, lane).x;
  }
  }
  int iter = 0, nunres = 0;
  bool unresolved = false;
  while (true) {
    if (!sync && done) break;
    if (!done) {
      mul_JT(R, qfc, force, lane);
      float g2 = 0, dec = 0;
      _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) {
        float gi = Ma[i] - qs[i] - qfc[i];
        grad[i] = gi; search[i] = -gi; g2 += gi * gi; dec += gi * gi / M[i * o.ldm + i];
      }
      g2 = warp_sum(g2); dec = warp_sum(dec);
      __syncwarp();
      if (iter >= m.iterations) done = true;
      else if (iter > 0 && scale * sqrtf(g2) < m.tolerance) done = true;
      // fp32: the last step's improvement was below the resolution of the cost.  The cost is blind to light dofs
      // (a 1e-6 kg m^2 finger tip at 100 rad/s^2 is 5e-3 of a cost of 1e4), the inertia-scaled gradient is not:
      // stop only when 0.5 g^T diag(M)^-1 g (>= the Newton decrement's order) is below the tolerance as well
      else if (unresolved && (scale * 0.5f * dec < m.tolerance || ++nunres > 3)) done = true;
    }
    // the one CTA-wide barrier of the iteration sits in front of the factorisation; it also tells
    // every warp whether any warp of the CTA is still iterating
    if (!hessian_solve_any<TILE>(R, S + o.H, M, S + o.tmpJ, o.ldm, search, lane, !done, !done, sync)) break;
    if (done) continue;
    // expected decrease 0.5 * |grad . search| below tolerance: converged (well conditioned in fp32)
    float gs = 0, ss = 0;
    _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) { gs += grad[i] * search[i]; ss += search[i] * search[i]; }
    gs = warp_sum(gs); ss = warp_sum(ss);
    if (!(gs < 0) || scale * 0.5f * (-gs) < m.tolerance) { done = true; continue; }
    symv(mv, M, search, nv, o.ldm, lane);
    mul_J(R, jv, search, lane);
    float q1 = 0, q2 = 0;
    _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) { q1 += search[i] * (Ma[i] - qs[i]); q2 += search[i] * mv[i]; }
    q1 = warp_sum(q1); q2 = warp_sum(q2);
    // exact line search: safeguarded Newton on the monotone derivative p'(a)
    float gtol = m.tolerance * m.ls_tolerance * sqrtf(ss) / scale;
    // p'(0) = grad . search = gs and p''(0) = search^T H search = -gs, because search solves H search = -grad
    // with the Hessian of the CURRENT constraint states: the probe at 0 is free and the first trial step is 1
    float p1 = gs, p2 = -gs, a = 1.f, lo = 0, hi = -1, dlo = gs, dhi = 0;
    const float d0 = gs;
    for (int it = 0; it < m.ls_iterations; it++) {
      { float4 e = eval_constraints<false>(R, a, lane); p1 = e.y + q1 + a * q2; p2 = e.z + q2; }
      if (fabsf(p1) < gtol || fabsf(p1) < 1e-6f * fabsf(d0)) break;
      if (p1 < 0) { lo = a; dlo = p1; } else { hi = a; dhi = p1; }
      float an = (p2 > 0) ? a - p1 / p2 : -1.f;
      if (hi < 0) {
        if (!(an > lo)) an = 2 * a + 1e-12f;
      } else {
        if (!(an > lo && an < hi)) {
          an = lo + (hi - lo) * (-dlo) / (dhi - dlo);
          if (!(an > lo && an < hi)) an = 0.5f * (lo + hi);
        }
        if (hi - lo <= 1e-6f * hi) { a = an; break; }
      }
      a = an;
    }
    if (!(a > 0)) { done = true; continue; }
    _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) { qacc[i] += a * search[i]; Ma[i] += a * mv[i]; }
    _Pragma("unroll 1") for (int r = lane; r < nefc; r += 32) jar[r] += a * jv[r];
    __syncwarp();
    float oldcost = cost;
    cost = eval_constraints<true>(R, 0.f, lane).x;
    {
      float gsum = 0;
      _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) gsum += 0.5f * (Ma[i] - qs[i]) * (qacc[i] - qas[i]);
      cost += warp_sum(gsum);
    }
    iter++;
    // improvement below the solver tolerance: stop; below what fp32 can resolve in the cost: let the gradient decide
    unresolved = oldcost - cost < 2e-7f * fabsf(cost);
    if (!unresolved && oldcost - cost < m.tolerance / scale) {
      mul_JT(R, qfc, force, lane);
      done = true;
    }
  }
  return iter;
}

// ----------------------------------------------------------------------------- S9: implicitfast + advance
template <int TILE>
__device__ __forceinline__ void integrate(const DevModel& m, float* S, int lane, bool active, int sync) {
  const EnvLayout& o = m.L;
  int nv = m.nv, ld = o.ldm;
  float h = m.timestep;
  float *A = S + o.H, *rhs = S + o.v_tmp, *qpos = S + o.qpos, *qvel = S + o.qvel;
  const float *M = S + o.M, *actforce = S + o.actforce;
  if (active) copy_matrix(A, M, nv, ld, lane);
  if (active) _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) {
    float* Ai = A + i * ld;
    Ai[i] += h * PKF(dof_damping)[i];
    for (int a = 0; a < m.nu; a++) {
      float b2 = PKF(actuator_biasprm)[3 * a + 2];
      if (b2 == 0) continue;
      if (PKI(actuator_forcelimited)[a] &&
          (actforce[a] <= PKF(actuator_forcerange)[2 * a] || actforce[a] >= PKF(actuator_forcerange)[2 * a + 1]))
        continue;
      const float* mom = PKF(act_moment) + a * nv;
      float mi = mom[i];
      if (mi == 0) continue;
      float s = -h * b2 * mi;
      for (int k = 0; k <= i; k++) Ai[k] += s * mom[k];
    }
    rhs[i] = S[o.qfrc_smooth + i] + S[o.qfrc_con + i];
  }
  __syncwarp();
  chol_solve_any<TILE>(A, nv, ld, rhs, lane, active, sync);
  if (!active) return;
  _Pragma("unroll 1") for (int i = lane; i < nv; i += 32) qvel[i] += h * rhs[i];
  __syncwarp();
  _Pragma("unroll 1") for (int j = lane; j < m.njnt; j += 32) {
    int qa = PKI(jnt_qposadr)[j], da = PKI(jnt_dofadr)[j];
    if (PKI(jnt_type)[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) qpos[qa + k] += h * qvel[da + k];
      float w[3] = {qvel[da + 3], qvel[da + 4], qvel[da + 5]};
      float n = sqrtf(dot3(w, w));
      float q[4] = {qpos[qa + 3], qpos[qa + 4], qpos[qa + 5], qpos[qa + 6]};
      if (n >= MINVAL) {
        float sn, cs;
        sincosf(0.5f * h * n, &sn, &cs);
        float inv = sn / n;
        float dq[4] = {cs, w[0] * inv, w[1] * inv, w[2] * inv}, r[4];
        quat_mul(r, q, dq);
        q[0] = r[0]; q[1] = r[1]; q[2] = r[2]; q[3] = r[3];
      }
      quat_normalize(q);
      qpos[qa + 3] = q[0]; qpos[qa + 4] = q[1]; qpos[qa + 5] = q[2]; qpos[qa + 6] = q[3];
    } else {
      qpos[qa] += h * qvel[da];
    }
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------- S4/S8: gyro + accelerometer
// (runs after the solver: region X is free again, cacc is rebuilt there with qacc included)
__device__ __forceinline__ void imu_sensors(const DevModel& m, float* S, const float* __restrict__ G, float* sensordata_env, int lane) {
  if (sensordata_env == nullptr) return;
  const EnvLayout& o = m.L;
  const float *cdof = S + o.cdof, *cdofdot = G + o.cdofdot, *qvel = S + o.qvel, *qacc = S + o.qacc;   // G: the env's global block (part B)
  float* cacc = S + o.cacc;
  if (lane < 6) cacc[lane] = (lane < 3) ? 0.f : -m.gravity[lane - 3];
  __syncwarp();
  for (int lv = 0; lv < m.nlevel; lv++) {
    for (int idx = PKI(lvl_adr)[lv] + lane; idx < PKI(lvl_adr)[lv + 1]; idx += 32) {
      int b = PKI(lvl_body)[idx], p = PKI(body_parentid)[b];
      float ca[6];
#pragma unroll
      for (int k = 0; k < 6; k++) ca[k] = cacc[6 * p + k];
      int da = PKI(body_dofadr)[b];
      for (int k = 0; k < PKI(body_dofnum)[b]; k++)
        for (int c = 0; c < 6; c++) ca[c] += cdofdot[6 * (da + k) + c] * qvel[da + k] + cdof[6 * (da + k) + c] * qacc[da + k];
#pragma unroll
      for (int k = 0; k < 6; k++) cacc[6 * b + k] = ca[k];
    }
    __syncwarp();
  }
  _Pragma("unroll 1") for (int s = lane; s < m.nsensor; s += 32) {
    int type = PKI(sensor_type)[s];
    if (type == SENS_RANGE) continue;
    int site = PKI(sensor_objid)[s], adr = PKI(sensor_adr)[s], b = PKI(site_bodyid)[site];
    float q[4], R[9], Rb[9], p[3], t[3];
    quat_mul(q, G + o.xquat + 4 * b, PKF(site_quat) + 4 * site);
    quat_normalize(q);
    quat2mat(R, q);
    quat2mat(Rb, G + o.xquat + 4 * b);
    mat_vec(t, Rb, PKF(site_pos) + 3 * site);
    p[0] = G[o.xpos + 3 * b] + t[0]; p[1] = G[o.xpos + 3 * b + 1] + t[1]; p[2] = G[o.xpos + 3 * b + 2] + t[2];
    const float *cv = G + o.cvel + 6 * b, *ca = cacc + 6 * b, *ref = G + o.xpos + 3 * PKI(root_list)[PKI(body_rootidx)[b]];
    float out[3];
    if (type == SENS_GYRO) {
      matT_vec(out, R, cv);
    } else {
      float dif[3] = {p[0] - ref[0], p[1] - ref[1], p[2] - ref[2]}, lin[3], vlin[3], al[3], wl[3], vl[3];
      cross3(t, ca, dif); lin[0] = ca[3] + t[0]; lin[1] = ca[4] + t[1]; lin[2] = ca[5] + t[2];
      cross3(t, cv, dif); vlin[0] = cv[3] + t[0]; vlin[1] = cv[4] + t[1]; vlin[2] = cv[5] + t[2];
      matT_vec(al, R, lin); matT_vec(wl, R, cv); matT_vec(vl, R, vlin);
      cross3(t, wl, vl);
      out[0] = al[0] + t[0]; out[1] = al[1] + t[1]; out[2] = al[2] + t[2];
    }
    sensordata_env[adr] = out[0]; sensordata_env[adr + 1] = out[1]; sensordata_env[adr + 2] = out[2];
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------- kernels
struct FwdInfo { int ncon, ns, nefc, iter, nnarrow; };
__device__ __forceinline__ bool warp_bad(const float* x, int n, int lane) {
  bool bad = false;
  _Pragma("unroll 1") for (int i = lane; i < n; i += 32) bad |= !(fabsf(x[i]) < MAXVAL);
  return __any_sync(FULL, bad);
}

// TMA bulk copy of the model pack into shared memory (one elected thread issues, all wait)
__device__ __forceinline__ void load_pack(const uint32_t* src, int nwords) {
  __shared__ __align__(8) unsigned long long bar;
  unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&bar);
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
  unsigned bytes = (unsigned)nwords * 4u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
    const unsigned CH = 16384;
    for (unsigned off = 0; off < bytes; off += CH) {
      unsigned n = bytes - off < CH ? bytes - off : CH;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                   "l"((const char*)src + off), "r"(n), "r"(bar_addr)
                   : "memory");
    }
  }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_PACK:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra DONE_PACK;\n"
      "bra WAIT_PACK;\n"
      "DONE_PACK:\n"
      "}\n" ::"r"(bar_addr)
      : "memory");
}

// ---- kernel 1: smooth dynamics + broadphase, one warp per env -------------------------------------------
// Reads the state, writes the env's persistent block (a.pb), the broadphase result (a.npass, a.slot_pair, the
// work-item queue a.items) and the pose / actuator observations.  Cost per env is uniform, so envs are
// assigned statically; the CTA-wide stage barriers (a.sync_level & 1) keep the warps in the same code region.
extern "C" __global__ void __launch_bounds__(512, 1) ss_smooth_kernel(const DevModel m, const StepArgs a) {
  const EnvLayout& o = m.L;
  load_pack(m.pack, m.pk.nwords);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* S = smem + m.pk.nwords + (size_t)warp * o.total;
  const bool sync = a.sync_level & 1;
  for (int base = blockIdx.x * wpb; base < a.nenv; base += gridDim.x * wpb) {
    bool active = base + warp < a.nenv;
    int env = a.order[active ? base + warp : a.nenv - 1];   // idle warps shadow another env and store nothing
    _Pragma("unroll 1") for (int i = lane; i < m.nq; i += 32) S[o.qpos + i] = a.qpos[(size_t)env * m.nq + i];
    _Pragma("unroll 1") for (int i = lane; i < m.nv; i += 32) { S[o.qvel + i] = a.qvel[(size_t)env * m.nv + i]; S[o.warm + i] = a.warm[(size_t)env * m.nv + i]; }
    _Pragma("unroll 1") for (int i = lane; i < m.nu; i += 32) S[o.ctrl + i] = a.ctrl[(size_t)env * m.nu + i];
    __syncwarp();
    int flags = a.env_flags ? a.env_flags[env] : 0;
    const int flags0 = flags;
    // mj_checkPos / mj_checkVel: a bad state is replaced by qpos0 at rest
    if (warp_bad(S + o.qpos, m.nq, lane) || warp_bad(S + o.qvel, m.nv, lane)) {
      _Pragma("unroll 1") for (int i = lane; i < m.nq; i += 32) S[o.qpos + i] = PKF(qpos0)[i];
      _Pragma("unroll 1") for (int i = lane; i < m.nv; i += 32) { S[o.qvel + i] = 0; S[o.warm + i] = 0; }
      __syncwarp();
      flags |= 1;
    }
    if (sync) __syncthreads();
    kinematics(m, S, lane);
    if (sync) __syncthreads();
    crb_mass_matrix(m, S, lane);
    if (sync) __syncthreads();
    if (active) broadphase(m, a, S, env, flags, lane);
    if (sync) __syncthreads();
    velocity_stage(m, S, lane);
    if (sync) __syncthreads();
    smooth_forces(m, S, lane);
    if (active) {
      float4* dst = reinterpret_cast<float4*>(a.pb + (size_t)env * a.pb_stride);
      const float4* src = reinterpret_cast<const float4*>(S);
      _Pragma("unroll 4") for (int i = lane; i < o.pb >> 2; i += 32) dst[i] = src[i];
      if (a.xpos) for (int i = lane; i < m.nbody * 3; i += 32) a.xpos[(size_t)env * m.nbody * 3 + i] = S[o.xpos + i];
      if (a.xquat) for (int i = lane; i < m.nbody * 4; i += 32) a.xquat[(size_t)env * m.nbody * 4 + i] = S[o.xquat + i];
      if (a.act_length) for (int i = lane; i < m.nu; i += 32) a.act_length[(size_t)env * m.nu + i] = S[o.actlen + i];
      if (a.act_velocity) for (int i = lane; i < m.nu; i += 32) a.act_velocity[(size_t)env * m.nu + i] = S[o.actvel + i];
      if (a.env_flags && lane == 0 && flags != flags0) a.env_flags[env] = flags;
    }
    __syncwarp();
  }
}

// ---- kernel 2: narrowphase, one warp per (env, candidate pair) work item ---------------------------------
// Work items come from the broadphase queue through an atomic counter, so the expensive queries (MPR on mesh
// hulls, five of them per penetrating pair with multiccd) spread over all SMs instead of stalling the warps that
// share a CTA with their env.  Hull vertices are read through L1 (this kernel uses almost no shared memory).
extern "C" __global__ void __launch_bounds__(256, 2) ss_narrow_kernel(const DevModel m, const StepArgs a) {
  load_pack(m.pack, m.pk.nwords);
  const EnvLayout& o = m.L;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int so = m.pk.nwords + warp * NP_SMEM;
  const int nitem = *a.item_count;
  for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(a.item_next, 1);
    it = __shfl_sync(FULL, it, 0);
    if (it >= nitem) break;
    int item = a.items[it], env = item / a.maxslot, pair = a.slot_pair[item];
    const float* S = a.pb + (size_t)env * a.pb_stride;
    int pc = PKI(pair_cg)[pair], c1 = pc & 0xffff, c2 = pc >> 16;
    Cvx A, B;
    make_cvx(m, S, c1, A);
    make_cvx(m, S, c2, B);
    narrow_pair(m, A, B, m.pair_margin[pair], m.multiccd ? 1e-3f * fminf(PKF(cg_rbound)[c1], PKF(cg_rbound)[c2]) : 0.f, so, lane);
    float* r = a.rec + (size_t)item * NP_REC;
    int n = 4 + 8 * __float_as_int(smem[so]);
    for (int i = lane; i < n; i += 32) r[i] = smem[so + i];
    __syncwarp();
  }
  (void)o;
}

// contacts of one env in the reference order (pair order, then the narrowphase's own order) from the records of
// its broadphase slots: lane = slot for the counts (prefix sum), then each lane expands its slot's contacts
__device__ __forceinline__ void gather_contacts(const DevModel& m, const StepArgs& a, float* S, int env, int& ncon, int& npass_out, int& flags, int lane) {
  const EnvLayout& o = m.L;
  int npass = a.npass[env];
  npass_out = npass;
  ncon = 0;
  for (int base = 0; base < npass; base += 32) {
    int slot = base + lane, cnt = 0, pair = 0;
    const float* r = nullptr;
    if (slot < npass) {
      int item = env * a.maxslot + slot;
      r = a.rec + (size_t)item * NP_REC;
      cnt = __float_as_int(r[0]);
      pair = a.slot_pair[item];
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
    int start = ncon + incl - cnt;
    if (cnt) {
      int pc = GPI(pair_cg)[pair], c1 = pc & 0xffff, c2 = pc >> 16;
      for (int c = 0; c < cnt; c++) {
        int k = start + c;
        if (k >= m.maxcon) { flags |= 2; break; }
        float* cr = S + o.con + k * CON_STRIDE;
        const float* rec = r + 4 + 8 * c;
        float x[3] = {rec[0], rec[1], rec[2]}, y[3] = {0, 1, 0}, z[3];
        if (x[1] > 0.5f || x[1] < -0.5f) { y[1] = 0; y[2] = 1; }
        float d = dot3(x, y);
        y[0] -= x[0] * d; y[1] -= x[1] * d; y[2] -= x[2] * d;
        normalize3(y);
        cross3(z, x, y);
        for (int t = 0; t < 3; t++) { cr[C_POS + t] = rec[4 + t]; cr[C_FRAME + t] = x[t]; cr[C_FRAME + 3 + t] = y[t]; cr[C_FRAME + 6 + t] = z[t]; }
        cr[C_DIST] = rec[3]; cr[C_MU] = 0;
        cr[C_DIM] = __int_as_float(m.pair_condim[pair]); cr[C_PAIR] = __int_as_float(pair); cr[C_EFC] = __int_as_float(-1);
        cr[C_BODY1] = __int_as_float(PKI(cg_bodyid)[c1]); cr[C_BODY2] = __int_as_float(PKI(cg_bodyid)[c2]);
        for (int t = 0; t < 5; t++) cr[C_FRICTION + t] = m.pair_friction[5 * pair + t];
      }
    }
    ncon += __shfl_sync(FULL, incl, 31);
  }
  flags |= __reduce_or_sync(FULL, (unsigned)flags);
  if (ncon > m.maxcon) ncon = m.maxcon;
  __syncwarp();
}

// ---- kernel 3: constraints, Newton solver, sensors, integration; one warp per env ------------------------
// Work distribution: a.sync_level & 8 = lockstep mode: groups of `wpb` consecutive slots of the cost-sorted env
// order (heaviest first, api.cu:schedule_kernel) are handed out through an atomic counter, the warps of a CTA
// own envs of similar cost and meet at one barrier per Newton iteration (they then stream the long
// straight-line factorisation through the instruction cache together).  Otherwise every warp fetches its
// own env and runs free.
template <int TILE>
__global__ void __launch_bounds__(256, 1) ss_solve_kernel(const DevModel m, const StepArgs a) {
  const EnvLayout& o = m.L;
  load_pack(m.pack, m.pk.nwords3);   // part 1 of the pack only; pair_cg (part 2) is read from global memory (GPI)
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* S = smem + m.pk.nwords3 + (size_t)warp * o.total;
  __shared__ int s_group;
  const bool lock = (a.sync_level & 8) != 0;
  const int bar = lock ? (1 | ((wpb * 32) << 8)) : 0;
  for (;;) {
    int slot;
    if (lock) {
      if (threadIdx.x == 0) s_group = atomicAdd(a.work_counter, wpb);
      __syncthreads();
      slot = s_group + warp;
      __syncthreads();
      if (slot - warp >= a.nenv) break;
    } else {
      slot = 0;
      if (lane == 0) slot = atomicAdd(a.work_counter, 1);
      slot = __shfl_sync(FULL, slot, 0);
      if (slot >= a.nenv) break;
    }
    bool active = slot < a.nenv;
    int env = a.order[active ? slot : a.nenv - 1];  // idle warps shadow another env read-only and store nothing
    {
      const float4* src = reinterpret_cast<const float4*>(a.pb + (size_t)env * a.pb_stride);
      float4* dst = reinterpret_cast<float4*>(S);
      _Pragma("unroll 4") for (int i = lane; i < o.pbA >> 2; i += 32) dst[i] = src[i];
    }
    if (a.sync_level & 96) {   // debug: scrub the env slice behind the persistent block (32: zeros, 64: NaNs) to expose reads of stale shared memory
      float fill = (a.sync_level & 64) ? __int_as_float(0x7fc00000) : 0.f;
      _Pragma("unroll 1") for (int i = o.pbA + lane; i < o.total; i += 32) S[i] = fill;
    }
    __syncwarp();
    int flags = a.env_flags ? a.env_flags[env] : 0;
    float time = a.time ? a.time[env] : 0.f;
    FwdInfo fi = {0, 0, 0, 0, 0};
    gather_contacts(m, a, S, env, fi.ncon, fi.nnarrow, flags, lane);
    const float* G = a.pb + (size_t)env * a.pb_stride;
    make_constraints(m, S, G, fi.ncon, fi.ns, fi.nefc, flags, lane);
    // qacc_smooth = M^-1 qfrc_smooth (factor lives in the H buffer until the solver rebuilds it)
    _Pragma("unroll 1") for (int i = lane; i < m.nv; i += 32) S[o.qacc_smooth + i] = S[o.qfrc_smooth + i];
    copy_matrix(S + o.H, S + o.M, m.nv, o.ldm, lane);
    chol_solve_any<TILE>(S + o.H, m.nv, o.ldm, S + o.qacc_smooth, lane, true, bar);
    fi.iter = solve_constraints<TILE>(m, S, fi.ns, fi.nefc, fi.ncon, lane, true, bar);
    int cost = a.cost_w * fi.iter + fi.nnarrow;   // this step's cost: the predictor for the next launch's schedule
    _Pragma("unroll 1") for (int i = lane; i < m.nv; i += 32) S[o.warm + i] = S[o.qacc + i];
    __syncwarp();
    // mj_checkAcc: a bad acceleration resets the env to qpos0 at rest (no integration in this step)
    bool bad = warp_bad(S + o.qacc, m.nv, lane);
    if (bad) {
      _Pragma("unroll 1") for (int i = lane; i < m.nq; i += 32) S[o.qpos + i] = PKF(qpos0)[i];
      _Pragma("unroll 1") for (int i = lane; i < m.nv; i += 32) { S[o.qvel + i] = 0; S[o.warm + i] = 0; S[o.qacc + i] = 0; S[o.qfrc_con + i] = 0; S[o.qfrc_smooth + i] = 0; }
      __syncwarp();
      flags |= 1;
    }
    if (active) {
      // observations of the state the step started from (same convention as mjData after mj_step)
      if (!bad) imu_sensors(m, S, G, a.sensordata ? a.sensordata + (size_t)env * m.nsensordata : nullptr, lane);
      if (a.qacc) for (int i = lane; i < m.nv; i += 32) a.qacc[(size_t)env * m.nv + i] = S[o.qacc + i];
      if (a.ncon && lane == 0) a.ncon[env] = fi.ncon;
      if (a.solver_iter && lane == 0) a.solver_iter[env] = fi.iter;
      if (a.contact_geom || a.contact_dist || a.dbg_contact_pos || a.dbg_contact_normal)
        _Pragma("unroll 1") for (int c = lane; c < m.maxcon; c += 32) {
          const float* con = S + o.con + c * CON_STRIDE;
          bool live = c < fi.ncon;
          size_t k = (size_t)env * m.maxcon + c;
          if (a.contact_geom) {
            int pc = live ? GPI(pair_cg)[__float_as_int(con[C_PAIR])] : 0;
            a.contact_geom[2 * k] = live ? PKI(cg_geomid)[pc & 0xffff] : -1;
            a.contact_geom[2 * k + 1] = live ? PKI(cg_geomid)[pc >> 16] : -1;
          }
          if (a.contact_dist) a.contact_dist[k] = live ? con[C_DIST] : 0.f;
          if (a.dbg_contact_pos) for (int t = 0; t < 3; t++) a.dbg_contact_pos[3 * k + t] = live ? con[C_POS + t] : 0.f;
          if (a.dbg_contact_normal) for (int t = 0; t < 3; t++) a.dbg_contact_normal[3 * k + t] = live ? con[C_FRAME + t] : 0.f;
        }
      if (a.dbg_nefc && lane == 0) a.dbg_nefc[env] = fi.nefc;
      if (a.dbg_M) for (int i = lane; i < m.nv * m.nv; i += 32) a.dbg_M[(size_t)env * m.nv * m.nv + i] = S[o.M + (i / m.nv) * o.ldm + (i % m.nv)];
      if (a.dbg_qacc_smooth) for (int i = lane; i < m.nv; i += 32) a.dbg_qacc_smooth[(size_t)env * m.nv + i] = S[o.qacc_smooth + i];
      if (a.dbg_qfrc_smooth) for (int i = lane; i < m.nv; i += 32) a.dbg_qfrc_smooth[(size_t)env * m.nv + i] = S[o.qfrc_smooth + i];
      if (a.dbg_qfrc_constraint) for (int i = lane; i < m.nv; i += 32) a.dbg_qfrc_constraint[(size_t)env * m.nv + i] = S[o.qfrc_con + i];
      __syncwarp();
    }
    if (!a.forward_only) {
      if (!bad || lock) integrate<TILE>(m, S, lane, !bad, bar);
      time += m.timestep;
      if (active) {
        _Pragma("unroll 1") for (int i = lane; i < m.nq; i += 32) a.qpos[(size_t)env * m.nq + i] = S[o.qpos + i];
        _Pragma("unroll 1") for (int i = lane; i < m.nv; i += 32) { a.qvel[(size_t)env * m.nv + i] = S[o.qvel + i]; a.warm[(size_t)env * m.nv + i] = S[o.warm + i]; }
        if (a.time && lane == 0) a.time[env] = time;
        if (lane == 0) a.cost[env] = cost;
      }
    }
    if (a.env_flags && lane == 0 && active) a.env_flags[env] = flags;
    __syncwarp();
  }
}

// ---- host side: the solve kernel is instantiated per register tile; api.cu picks one by the model's nv -----------
int ss_solve_tile(int nv) { return nv <= 28 ? 28 : nv <= 32 ? 32 : nv <= 40 ? 40 : nv <= 44 ? 44 : nv <= 48 ? 48 : 64; }
#define SS_TILE_DISPATCH(tile, expr)                  \
  switch (tile) {                                     \
    case 28: { constexpr int T = 28; expr; } break;   \
    case 32: { constexpr int T = 32; expr; } break;   \
    case 40: { constexpr int T = 40; expr; } break;   \
    case 44: { constexpr int T = 44; expr; } break;   \
    case 48: { constexpr int T = 48; expr; } break;   \
    default: { constexpr int T = 64; expr; } break;   \
  }
cudaError_t ss_solve_set_smem(int tile, int bytes) {
  cudaError_t e = cudaSuccess;
  SS_TILE_DISPATCH(tile, e = cudaFuncSetAttribute(ss_solve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return e;
}
void ss_solve_launch(int tile, int grid, int block, size_t smem, cudaStream_t st, const DevModel& m, const StepArgs& a) {
  SS_TILE_DISPATCH(tile, (ss_solve_kernel<T><<<grid, block, smem, st>>>(m, a)));
}
