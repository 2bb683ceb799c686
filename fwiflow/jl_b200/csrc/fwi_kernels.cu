// Hand-written CUDA kernels (sm_100a) for the 2-D elastic FWI hot path.
//
// Three fused per-time-step kernels, each advancing a whole BATCH of shots per launch:
//   fwd_step_kernel     stress + source + velocity + record (+ boundary-frame save)
//                       replaces el_stress / add_source / el_velocity / recording / from_bnd x5
//                       (reference: libCUFD.cu:202-240)
//   rev_image_kernel    reverse velocity + frame restore + source removal + reverse stress
//                       + frame restore + lambda/mu/rho imaging (deterministic gather)
//                       replaces el_velocity(false) / to_bnd x5 / add_source(false) / el_stress(false)
//                       (reference: libCUFD.cu:380-403)
//   adj_step_kernel     source_grad + adjoint velocity + residual injection + adjoint stress
//                       replaces source_grad / el_velocity_adj / res_injection / el_stress_adj
//                       (reference: libCUFD.cu:376,405-427)
// All three use shared-memory tiles with halos (the second half-step is computed from the
// first half-step's tile without a round trip to HBM), staged with cp.async, one float4 "quad"
// of 4 consecutive z cells per thread (16-byte loads / stores everywhere), ping-pong state
// buffers, and the reference's arithmetic (float storage; with FWI_FP64_PROMOTE=1 the double
// promotion of the reference's C expressions is reproduced).  Derivatives multiply by 1/dz
// instead of dividing (<= 1 ulp per derivative).
#include <cstdio>
#include <cstdlib>

#include "fwi_kernels.cuh"

#ifndef FWI_FP64_PROMOTE
#define FWI_FP64_PROMOTE 0
#endif

namespace fwi {
namespace {

constexpr float C1 = 1.125f;
constexpr float C2 = (float)(1.0 / 24.0);
constexpr float SRC_SCALE = 2250000.0f;  // pow(1500,2)  utilities.cu:528

__device__ __forceinline__ float *plane_of(float *state, const Grid &g, int shot, int slot) {
  return state + ((long long)shot * S_COUNT + slot) * g.plane + g.origin;
}

// staggered first derivatives on a tile pointer p (center), stride s
__device__ __forceinline__ float d_minus(const float *p, int s, float rh) {
  return (C1 * (p[0] - p[-s]) - C2 * (p[s] - p[-2 * s])) * rh;
}
__device__ __forceinline__ float d_plus(const float *p, int s, float rh) {
  return (C1 * (p[s] - p[0]) - C2 * (p[2 * s] - p[-s])) * rh;
}
// adjoint-kernel spelling (el_stress_adj.cu:54-61): (-c1*(..) + c2*(..))/h
__device__ __forceinline__ float ad_minus(const float *p, int s, float rh) {
  return (-C1 * (p[0] - p[-s]) + C2 * (p[s] - p[-2 * s])) * rh;
}
__device__ __forceinline__ float ad_plus(const float *p, int s, float rh) {
  return (-C1 * (p[s] - p[0]) + C2 * (p[2 * s] - p[-s])) * rh;
}
__device__ __forceinline__ float ad_minus4(float m2, float m1, float c0, float p1, float rh) {
  return (-C1 * (c0 - m1) + C2 * (p1 - m2)) * rh;
}
__device__ __forceinline__ float ad_plus4(float m1, float c0, float p1, float p2, float rh) {
  return (-C1 * (p1 - c0) + C2 * (p2 - m1)) * rh;
}

// sigma += ((lam+2mu) e1 + lam e2) dt  with the reference's promotion (el_stress.cu:66-67)
__device__ __forceinline__ float stress_inc(float s, float lam, float mu, float e1, float e2, float dt, float sign) {
#if FWI_FP64_PROMOTE
  const double l2m = (double)lam + 2.0 * (double)mu;
  const double t = (l2m * (double)e1 + (double)(lam * e2)) * (double)dt;
  return (float)((double)s + (double)sign * t);
#else
  return s + sign * (((lam + 2.0f * mu) * e1 + lam * e2) * dt);
#endif
}

__device__ __forceinline__ bool z_in_pml(const Grid &g, int z) { return z < g.nPml || z > g.nz - g.nPml - g.nPad - 1; }
__device__ __forceinline__ bool x_in_pml_s(const Grid &g, int x) { return x < g.nPml || x > g.nx - g.nPml - 1; }
__device__ __forceinline__ bool x_in_pml_v(const Grid &g, int x) { return x < g.nPml || x > g.nx - g.nPml; }
__device__ __forceinline__ bool is_active(const Grid &g, int z, int x) {
  return z >= 2 && z <= g.az_hi && x >= 2 && x <= g.ax_hi;
}
__device__ __forceinline__ bool in_box(const Grid &g, int z, int x) {
  return z >= g.zlo && z <= g.zhi && x >= g.xlo && x <= g.xhi;
}

// =================================================================================================
// forward step: one float4 "quad" (4 consecutive z cells) per thread, 16 quads = one sigma-tile
// column per half-warp.  Velocity tile (halo rounded up to whole quads) staged with cp.async.
// =================================================================================================
struct F4 {
  float v[4];
};
__device__ __forceinline__ F4 ld4(const float *p) {
  const float4 t = *reinterpret_cast<const float4 *>(p);
  return F4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ void st4(float *p, const F4 &a) {
  *reinterpret_cast<float4 *>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// 7 consecutive samples w[0..6] = f[z-2 .. z+4]  ->  D-z at the 4 cells z..z+3
__device__ __forceinline__ void dz_minus4(const F4 &A, const F4 &B, const F4 &C, float rh, float *out) {
  const float w[7] = {A.v[2], A.v[3], B.v[0], B.v[1], B.v[2], B.v[3], C.v[0]};
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = (C1 * (w[k + 2] - w[k + 1]) - C2 * (w[k + 3] - w[k])) * rh;
}
// samples u[0..6] = f[z-1 .. z+5]  ->  D+z at the 4 cells
__device__ __forceinline__ void dz_plus4(const F4 &A, const F4 &B, const F4 &C, float rh, float *out) {
  const float u[7] = {A.v[3], B.v[0], B.v[1], B.v[2], B.v[3], C.v[0], C.v[1]};
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = (C1 * (u[k + 2] - u[k + 1]) - C2 * (u[k + 3] - u[k])) * rh;
}
// columns x-2, x-1, x, x+1 -> D-x ;  columns x-1, x, x+1, x+2 -> D+x   (same expression shape)
__device__ __forceinline__ void dx4(const F4 &m2, const F4 &m1, const F4 &c0, const F4 &p1, float rh, float *out) {
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = (C1 * (c0.v[k] - m1.v[k]) - C2 * (p1.v[k] - m2.v[k])) * rh;
}

constexpr int QS = (TILE_Z + 8) / 4;      // 16 quads: sigma-tile rows z0-4 .. z0+TILE_Z+3
constexpr int QV = (TILE_Z + 16) / 4;     // 18 quads: velocity-tile rows z0-8 .. z0+TILE_Z+7
constexpr int VP = QV * 4, SP = QS * 4;   // row pitches (floats)
constexpr int VC = TILE_X + 6, SC = TILE_X + 4;
static_assert(TILE_Z % 4 == 0 && QS == 16, "half-warp per sigma column");

// =================================================================================================
// model preparation
// =================================================================================================
__global__ void model_transpose_kernel(Grid g, const double *__restrict__ lam_in, const double *__restrict__ mu_in,
                                       const double *__restrict__ den_in, float *model) {
  __shared__ float t[3][32][33];
  const int xb = blockIdx.x * 32, zb = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int z = zb + r, x = xb + threadIdx.x;
    if (z < g.nz && x < g.nx) {
      const long long k = (long long)z * g.nx + x;  // row-major [z][x]  (libCUFD.cu:72-77)
      t[0][r][threadIdx.x] = (float)(lam_in[k] * 1e6);
      t[1][r][threadIdx.x] = (float)(mu_in[k] * 1e6);
      t[2][r][threadIdx.x] = (float)den_in[k];
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = xb + r, z = zb + threadIdx.x;
    if (z < g.nz && x < g.nx) {
      const long long o = g.origin + (long long)x * g.P + z;
      model[M_LAM * g.plane + o] = t[0][threadIdx.x][r];
      model[M_MU * g.plane + o] = t[1][threadIdx.x][r];
      model[M_DEN * g.plane + o] = t[2][threadIdx.x][r];
    }
  }
}

// mu_bar, averaged buoyancies (utilities.cu:125-152, Model.cu:67-73), max cp (utilities.cu:109-123), and the
// dt-scaled coefficient planes the step kernels read: lambda dt, (lambda + 2 mu) dt, mu_bar dt, byc_a dt, byc_b dt
__global__ void model_derive_kernel(Grid g, float *model, unsigned int *cpmax_bits) {
  const float *lam = model + M_LAM * g.plane, *mu = model + M_MU * g.plane, *den = model + M_DEN * g.plane;
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int x = blockIdx.y;
  float cp = 0.0f;
  if (z < g.nz) {
    const long long o = g.origin + (long long)x * g.P + z;
    float m = 0.0f, ba = (float)(1.0 / 1000.0), bb = (float)(1.0 / 1000.0);
    if (z >= 2 && z <= g.nz - 3 && x >= 2 && x <= g.nx - 3) {
      const float a = mu[o], b = mu[o + 1], c = mu[o + g.P], d = mu[o + g.P + 1];
      if (!(a == 0.0f || b == 0.0f || c == 0.0f || d == 0.0f))
        m = (float)(4.0 / (1.0 / (double)a + 1.0 / (double)b + 1.0 / (double)c + 1.0 / (double)d));
      ba = (float)(2.0 / (double)(den[o + 1] + den[o]));
      bb = (float)(2.0 / (double)(den[o + g.P] + den[o]));
    }
    model[M_AMU * g.plane + o] = m;
    model[M_BYA * g.plane + o] = ba;
    model[M_BYB * g.plane + o] = bb;
    // active region of both half-steps: 2 <= z <= nz-nPad-3, 2 <= x <= nx-3 (el_stress.cu:52, el_velocity.cu:47);
    // zero coefficients elsewhere make the update a no-op there without any predicate in the step kernels
    const bool act = z >= 2 && z <= g.az_hi && x >= 2 && x <= g.ax_hi;
    const double dt = act ? (double)g.dt : 0.0;
    model[M_LDT * g.plane + o] = (float)((double)lam[o] * dt);
    model[M_L2MDT * g.plane + o] = (float)(((double)lam[o] + 2.0 * (double)mu[o]) * dt);
    model[M_AMUDT * g.plane + o] = (float)((double)m * dt);
    model[M_BYADT * g.plane + o] = (float)((double)ba * dt);
    model[M_BYBDT * g.plane + o] = (float)((double)bb * dt);
    cp = (float)sqrt(((double)lam[o] + 2.0 * (double)mu[o]) / (double)den[o]);
    if (!(cp > 0.0f)) cp = 0.0f;  // NaN / negative never wins the max
  }
  for (int s = 16; s > 0; s >>= 1) cp = fmaxf(cp, __shfl_xor_sync(0xffffffffu, cp, s));
  if ((threadIdx.x & 31) == 0 && cp > 0.0f) atomicMax(cpmax_bits, __float_as_uint(cp));
}

// =================================================================================================
// residual / misfit
// =================================================================================================
// cuda_window with per-trace limits (utilities.cu:654-706), literally: float times, the ramps evaluated through
// double sin / cos, amp^2 * weight as the multiplier.  The reference leaves a trace UNTOUCHED (no weight either) when
// the window is empty ("Window error 1") -- reproduced.
__device__ __forceinline__ float trace_window(int idt, int nt, float dt, float t0, float t3, float weight, float ratio) {
  const double PI = 3.141592653589793238462643383279502884197169;
  const float t = idt * dt;
  const float t_max = nt * dt;
  if (t0 < 0.0f) t0 = 0.0f;
  if (t0 > t_max) t0 = t_max;
  if (t3 < 0.0f) t3 = 0.0f;
  if (t3 > t_max) t3 = t_max;
  const float offset = (t3 - t0) * ratio;
  if (offset <= 0.0f) return 1.0f;
  const float t1 = t0 + offset, t2 = t3 - offset;
  float amp;
  if (t >= t0 && t < t1)
    amp = (float)sin(PI / 2.0 * (double)(t - t0) / (double)(t1 - t0));
  else if (t >= t1 && t < t2)
    amp = 1.0f;
  else if (t >= t2 && t < t3)
    amp = (float)cos(PI / 2.0 * (double)(t - t2) / (double)(t3 - t2));
  else
    amp = 0.0f;
  return amp * amp * weight;
}

__global__ void residual_kernel(ResidualArgs a) {
  __shared__ float t_obs[32][33];
  __shared__ float t_res[32][33];
  __shared__ float t_syn[32][33];
  __shared__ double red[8];
  const int tb = blockIdx.x * 32, rb = blockIdx.y * 32;
  // observed data: [rec][time], time fastest
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int rec = rb + r, t = tb + threadIdx.x;
    t_obs[r][threadIdx.x] = (rec < a.nrec && t < a.nSteps) ? a.obs_rt[(long long)rec * a.nSteps + t] : 0.0f;
  }
  __syncthreads();
  double acc = 0.0;
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int t = tb + q, rec = rb + threadIdx.x;
    float res = 0.0f, sc = 0.0f, oc = 0.0f;
    if (t < a.nSteps && rec < a.nrec) {
      float w = a.w2[t], wr = w;
      if (a.win_start) {   // if_win: per-trace window + weight, ramp ratio 0.005 on the data and 0.1 on the residual
        w = trace_window(t, a.nSteps, a.dt, a.win_start[rec], a.win_end[rec], a.weights[rec], 0.005f);
        wr = trace_window(t, a.nSteps, a.dt, a.win_start[rec], a.win_end[rec], a.weights[rec], 0.1f);
      }
      oc = t_obs[threadIdx.x][q] * w;                   // cuda_window on obs  (libCUFD.cu:259-268)
      sc = a.syn_tr[(long long)t * a.nrp + rec] * w;    // cuda_window on syn  (libCUFD.cu:263-270)
      res = (t > 0) ? oc - sc : 0.0f;                   // gpuMinus            (utilities.cu:154-167)
      acc += (double)(res * res);                       // cuda_cal_objective  (utilities.cu:169-205)
      res *= wr;                                        // cuda_window on res  (libCUFD.cu:305-312)
      a.res_tr[(long long)t * a.nrp + rec] = res;
    }
    t_res[q][threadIdx.x] = res;
    t_syn[q][threadIdx.x] = sc;
    t_obs[threadIdx.x][q] = oc;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int rec = rb + r, t = tb + threadIdx.x;
    if (rec < a.nrec && t < a.nSteps) {
      const long long k = (long long)rec * a.nSteps + t;
      if (a.res_rt) a.res_rt[k] = t_res[threadIdx.x][r];
      if (a.syn_rt) a.syn_rt[k] = t_syn[threadIdx.x][r];
      if (a.obs_cond_rt) a.obs_cond_rt[k] = t_obs[r][threadIdx.x];
    }
  }
  // deterministic block reduction
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  const int w = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
  if (threadIdx.x == 0) red[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    double s = 0.0;
    for (int k = 0; k < (int)(blockDim.x * blockDim.y) / 32; k++) s += red[k];
    a.partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

__global__ void sum_partials_kernel(const double *partial, int n, float *out_j) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_j = (float)red[0];
}

// misfit = 0.5 * sum over shots of J_shot, accumulated in float in shot order (libCUFD.cu:110,294,529)
__global__ void misfit_kernel(const float *j_shot, int n, float *misfit_half) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < n; i++) s += j_shot[i];
    *misfit_half = (float)(0.5 * (double)s);
  }
}

__global__ void traces_to_rt_kernel(const float *tr, float *rt, int nrec, int nrp, int nSteps) {
  __shared__ float t[32][33];
  const int tb = blockIdx.x * 32, rb = blockIdx.y * 32;
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int ti = tb + q, rec = rb + threadIdx.x;
    t[q][threadIdx.x] = (ti < nSteps && rec < nrec) ? tr[(long long)ti * nrp + rec] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int rec = rb + r, ti = tb + threadIdx.x;
    if (rec < nrec && ti < nSteps) rt[(long long)rec * nSteps + ti] = t[threadIdx.x][r];
  }
}

// result planes are row-major [z][x] (libCUFD.cu:480-486).  Sums the per-slot accumulators in slot order and
// applies the reference's 4-point spray of the mu / rho imaging terms as a gather (el_stress.cu:113-124,
// el_velocity.cu:105-110, incl. the always-true x+1 guard that lands in column xhi+1).
__global__ void finalize_kernel(Grid g, const float *gacc, int nslots, const float *mu, const float *misfit_half,
                                float *result) {
  __shared__ float t[3][32][33];
  const int zb = blockIdx.x * 32, xb = blockIdx.y * 32;
  auto S = [&](int which, int z, int x) -> float {
    float s = 0.0f;
    const long long o = g.origin + (long long)x * g.P + z;
    for (int q = 0; q < nslots; q++) s += gacc[((long long)q * G_COUNT + which) * g.plane + o];
    return s;
  };
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = xb + r, z = zb + threadIdx.x;
    float gl = 0.0f, gm = 0.0f, gd = 0.0f;
    if (z >= g.zlo && z <= g.zhi && x >= g.xlo && x <= g.xhi + 1) {
      gl = S(G_LAM, z, x);
      gm = S(G_MU, z, x);
      float G = S(G_MUS, z, x) + S(G_MUS, z - 1, x) + S(G_MUS, z, x - 1);
      if (x <= g.xhi) G += S(G_MUS, z - 1, x - 1);
      if (G != 0.0f) {
        const float m = mu[(long long)x * g.P + z];
        gm += G / (m * m);
      }
      gd = S(G_RHO_A, z, x) + S(G_RHO_B, z, x) + S(G_RHO_A, z - 1, x) + S(G_RHO_B, z, x - 1);
    }
    t[0][r][threadIdx.x] = gl;
    t[1][r][threadIdx.x] = gm;
    t[2][r][threadIdx.x] = gd;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int z = zb + r, x = xb + threadIdx.x;
    if (z < g.nz && x < g.nx)
      for (int k = 0; k < 3; k++) result[((long long)k * g.nz + z) * g.nx + x] = t[k][threadIdx.x][r];
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0)
    result[3LL * g.nz * g.nx] = *misfit_half;
}

}  // namespace

// =================================================================================================
// launchers
// =================================================================================================

template <typename K>
static void configure_one(K kernel, size_t smem, const char *env, int carveout) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // shared-memory carveout in percent (-1 = driver default).  The rest of the 256 KB stays L1: the kernels that
  // read coefficients / CPML memory straight from global want it, the register-heavy adjoint step wants 2 CTAs.
  if (const char *e = getenv(env)) carveout = atoi(e);
  if (carveout >= 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
}

void configure_forward_kernels();   // fwi_forward.cu
void configure_backward_kernels();  // fwi_backward.cu

void configure_kernels() {
  configure_forward_kernels();
  configure_backward_kernels();
}

void launch_model_prep(const Grid &g, const double *d_lam, const double *d_mu, const double *d_den, float *model,
                       unsigned int *cpmax_bits, cudaStream_t s) {
  dim3 tb(32, 8);
  dim3 tg((g.nx + 31) / 32, (g.nz + 31) / 32);
  model_transpose_kernel<<<tg, tb, 0, s>>>(g, d_lam, d_mu, d_den, model);
  dim3 dg((g.nz + 127) / 128, g.nx);
  model_derive_kernel<<<dg, 128, 0, s>>>(g, model, cpmax_bits);
}

void launch_residual(const ResidualArgs &a, int *nblocks_out, cudaStream_t s) {
  dim3 tb(32, 8);
  dim3 tg((a.nSteps + 31) / 32, (a.nrec + 31) / 32);
  if (nblocks_out) *nblocks_out = tg.x * tg.y;
  residual_kernel<<<tg, tb, 0, s>>>(a);
}

void launch_sum_partials(const double *partial, int n, float *out_j, cudaStream_t s) {
  sum_partials_kernel<<<1, 256, 0, s>>>(partial, n, out_j);
}

void launch_misfit(const float *j_shot, int n, float *misfit_half, cudaStream_t s) {
  misfit_kernel<<<1, 32, 0, s>>>(j_shot, n, misfit_half);
}

void launch_traces_to_rt(const float *tr, float *rt, int nrec, int nrp, int nSteps, cudaStream_t s) {
  dim3 tb(32, 8);
  dim3 tg((nSteps + 31) / 32, (nrec + 31) / 32);
  traces_to_rt_kernel<<<tg, tb, 0, s>>>(tr, rt, nrec, nrp, nSteps);
}

void launch_finalize(const Grid &g, const float *gacc, int nslots, const float *mu, const float *misfit_half,
                     float *result, cudaStream_t s) {
  dim3 tb(32, 8);
  dim3 tg((g.nz + 31) / 32, (g.nx + 31) / 32);
  finalize_kernel<<<tg, tb, 0, s>>>(g, gacc, nslots, mu, misfit_half, result);
}

}  // namespace fwi
