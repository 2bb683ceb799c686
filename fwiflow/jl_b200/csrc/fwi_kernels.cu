// Support kernels of the FWI hot path (sm_100a): everything that runs once per gradient or once per shot rather than
// once per time step.  The three time-step kernels live in fwi_forward.cu / fwi_backward.cu.
//   model_transpose / model_derive   caller's double row-major MPa -> float planes, derived coefficients, max cp
//                                    (reference: Model.cu:36-87, utilities.cu:109-152)
//   residual / sum_partials / misfit taper or per-trace windows, res = obs - syn, sum res^2
//                                    (reference: libCUFD.cu:254-330, utilities.cu:154-205,654-747)
//   traces_to_rt                     [step][receiver] -> [receiver][time] (Shot<id>.bin layout)
//   finalize                         per-slot imaging accumulators -> [grad_lambda|grad_mu|grad_den|misfit], with the
//                                    reference's 4-point spray applied as a gather (el_stress.cu:113-124,
//                                    el_velocity.cu:101-110)
#include <cstdio>
#include <cstdlib>

#include "fwi_kernels.cuh"

namespace fwi {
namespace {

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

// the same for a caller whose (nz, nx) arrays are COLUMN-major (Julia): element (z, x) at x * nz + z is already the
// device order, so there is nothing to transpose (the optional layout flag)
__global__ void model_convert_cm_kernel(Grid g, const double *__restrict__ lam_in, const double *__restrict__ mu_in,
                                        const double *__restrict__ den_in, float *model) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
  if (z >= g.nz) return;
  const long long k = (long long)x * g.nz + z, o = g.origin + (long long)x * g.P + z;
  model[M_LAM * g.plane + o] = (float)(lam_in[k] * 1e6);
  model[M_MU * g.plane + o] = (float)(mu_in[k] * 1e6);
  model[M_DEN * g.plane + o] = (float)den_in[k];
}

// =================================================================================================
// velocity-space front end (src/FWI.jl:156-205, src/Utils.jl:221-227 on the device)
// =================================================================================================
// tf.pad(a, [nPml (nPml + nPad); nPml nPml], "SYMMETRIC"): padded index -> source index (edge repeated, reflecting on)
__device__ __forceinline__ int sym_index(int i, int pad, int n) {
  const int m = 2 * n;
  int r = (i - pad) % m;
  if (r < 0) r += m;
  return r < n ? r : m - 1 - r;
}
// the reference's gradient mask (src/FWI.jl:45-49): 1 inside the absorbing layers, minus the 10 rows under the top one
__device__ __forceinline__ bool fwi_mask(const Grid &g, int z, int x) {
  const int nz0 = g.nz - 2 * g.nPml - g.nPad, nx0 = g.nx - 2 * g.nPml;
  return z >= g.nPml + 10 && z < g.nPml + nz0 && x >= g.nPml && x < g.nPml + nx0;
}

// cp, cs, rho (unpadded (nz0, nx0) or padded (nz, nx), the caller's layout) -> padded, mask-blended with the reference
// models unless is_masked (src/FWI.jl:174-176) -> vel [3][nz nx] (kept for the chain rule) and lambda, mu [MPa], rho
// in model_in, all in the caller's layout and in double with NumPy's / Julia's operation order (no contraction), so that
// the float planes derived from them are the ones the host path produces.
__global__ void velocity_prep_kernel(Grid g, int column_major, int padded, int is_masked, const double *__restrict__ in,
                                     long long n_in, double *__restrict__ vel, double *__restrict__ model_in) {
  const long long n = (long long)g.nz * g.nx;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int z = column_major ? (int)(k % g.nz) : (int)(k / g.nx), x = column_major ? (int)(k / g.nz) : (int)(k % g.nx);
  const int nz0 = g.nz - 2 * g.nPml - g.nPad, nx0 = g.nx - 2 * g.nPml;
  long long j = k;
  if (!padded) {
    const int zs = sym_index(z, g.nPml, nz0), xs = sym_index(x, g.nPml, nx0);
    j = column_major ? (long long)xs * nz0 + zs : (long long)zs * nx0 + xs;
  }
  const bool use_ref = !is_masked && !fwi_mask(g, z, x);
  const double *src = in + (use_ref ? 3 * n_in : 0);
  const double cp = src[j], cs = src[n_in + j], den = src[2 * n_in + j];
  vel[k] = cp; vel[n + k] = cs; vel[2 * n + k] = den;
  const double cs2 = __dmul_rn(__dmul_rn(2.0, cs), cs);
  model_in[k] = __ddiv_rn(__dmul_rn(__dsub_rn(__dmul_rn(cp, cp), cs2), den), 1e6);
  model_in[n + k] = __ddiv_rn(__dmul_rn(__dmul_rn(cs, cs), den), 1e6);
  model_in[2 * n + k] = den;
}

// chain rule of velocity_to_moduli (what TF autodiff applies in the reference): packed float gradients w.r.t.
// (lambda, mu, rho) -> double gradients w.r.t. (cp, cs, rho) on the padded grid, times the mask unless is_masked
__global__ void velocity_grad_kernel(Grid g, int column_major, int is_masked, const float *__restrict__ result,
                                     const double *__restrict__ vel, double *__restrict__ out) {
  const long long n = (long long)g.nz * g.nx;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int z = column_major ? (int)(k % g.nz) : (int)(k / g.nx), x = column_major ? (int)(k / g.nz) : (int)(k % g.nx);
  const double gl = result[k], gm = result[n + k], gd = result[2 * n + k];
  const double cp = vel[k], cs = vel[n + k], den = vel[2 * n + k];
  double g_cp = __dmul_rn(__ddiv_rn(__dmul_rn(__dmul_rn(2.0, cp), den), 1e6), gl);
  double g_cs = __ddiv_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(-4.0, gl), __dmul_rn(2.0, gm)), cs), den), 1e6);
  const double a = __dmul_rn(__dsub_rn(__dmul_rn(cp, cp), __dmul_rn(__dmul_rn(2.0, cs), cs)), gl);
  const double b = __dmul_rn(__dmul_rn(cs, cs), gm);
  double g_rho = __dadd_rn(gd, __ddiv_rn(__dadd_rn(a, b), 1e6));
  if (!is_masked && !fwi_mask(g, z, x)) g_cp = g_cs = g_rho = 0.0;
  out[k] = g_cp; out[n + k] = g_cs; out[2 * n + k] = g_rho;
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
// applies the reference's 4-point spray of the mu imaging term as a gather (el_stress.cu:113-124, incl. the always-true
// x+1 guard that lands in column xhi+1); the density spray (el_velocity.cu:105-110) is gathered by the reverse kernel.
__global__ void finalize_kernel(Grid g, const float *gacc, int nslots, const float *mu, const float *misfit_half,
                                float *result, int column_major) {
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
      gd = S(G_RHO, z, x);   // the density spray is gathered per time step by the reverse kernel
    }
    t[0][r][threadIdx.x] = gl;
    t[1][r][threadIdx.x] = gm;
    t[2][r][threadIdx.x] = gd;
    if (column_major && z < g.nz && x < g.nx) {   // the caller's order is the device order: no transpose
      const long long n = (long long)g.nz * g.nx, k = (long long)x * g.nz + z;
      result[k] = gl; result[n + k] = gm; result[2 * n + k] = gd;
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32 && !column_major; r += blockDim.y) {
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
                       unsigned int *cpmax_bits, int column_major, cudaStream_t s) {
  if (column_major) {
    dim3 cg((g.nz + 127) / 128, g.nx);
    model_convert_cm_kernel<<<cg, 128, 0, s>>>(g, d_lam, d_mu, d_den, model);
  } else {
    dim3 tb(32, 8);
    dim3 tg((g.nx + 31) / 32, (g.nz + 31) / 32);
    model_transpose_kernel<<<tg, tb, 0, s>>>(g, d_lam, d_mu, d_den, model);
  }
  dim3 dg((g.nz + 127) / 128, g.nx);
  model_derive_kernel<<<dg, 128, 0, s>>>(g, model, cpmax_bits);
}

void launch_velocity_prep(const Grid &g, int column_major, int padded, int is_masked, const double *in, long long n_in,
                          double *vel, double *model_in, cudaStream_t s) {
  const long long n = (long long)g.nz * g.nx;
  velocity_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, column_major, padded, is_masked, in, n_in, vel, model_in);
}

void launch_velocity_grad(const Grid &g, int column_major, int is_masked, const float *result, const double *vel, double *out,
                          cudaStream_t s) {
  const long long n = (long long)g.nz * g.nx;
  velocity_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, column_major, is_masked, result, vel, out);
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
  if (nrec <= 0) return;   // a shot without receivers records nothing (empty Shot<id>.bin)
  dim3 tb(32, 8);
  dim3 tg((nSteps + 31) / 32, (nrec + 31) / 32);
  traces_to_rt_kernel<<<tg, tb, 0, s>>>(tr, rt, nrec, nrp, nSteps);
}

void launch_finalize(const Grid &g, const float *gacc, int nslots, const float *mu, const float *misfit_half,
                     float *result, int column_major, cudaStream_t s) {
  dim3 tb(32, 8);
  dim3 tg((g.nz + 31) / 32, (g.nx + 31) / 32);
  finalize_kernel<<<tg, tb, 0, s>>>(g, gacc, nslots, mu, misfit_half, result, column_major);
}

}  // namespace fwi
