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
// first half-step's tile without a round trip to HBM), ping-pong state buffers, and the
// reference's arithmetic (float storage, double promotion where the reference's C
// expressions promote -- FWI_FP64_PROMOTE).  Derivatives multiply by 1/dz instead of
// dividing (<= 1 ulp per derivative).
#include <cstdio>

#include "fwi_kernels.cuh"

#ifndef FWI_FP64_PROMOTE
#define FWI_FP64_PROMOTE 1
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

__device__ __forceinline__ int frame_index(const Grid &g, int z, int x) {
  if (z < g.zlo - 2 || z > g.zhi + 2 || x < g.xlo - 2 || x > g.xhi + 2) return -1;
  const int zr = z - (g.zlo - 2);
  if (x <= g.xlo + 2) return (x - (g.xlo - 2)) * g.f_nzB + zr;
  if (x >= g.xhi - 2) return (5 + x - (g.xhi - 2)) * g.f_nzB + zr;
  if (z <= g.zlo + 2) return 10 * g.f_nzB + (x - (g.xlo + 3)) * 10 + zr;
  if (z >= g.zhi - 2) return 10 * g.f_nzB + (x - (g.xlo + 3)) * 10 + 5 + (z - (g.zhi - 2));
  return -1;
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
// forward step
// =================================================================================================
constexpr int FV_Z = TILE_Z + 6, FV_X = TILE_X + 6;  // velocity tile (halo 3)
constexpr int FS_Z = TILE_Z + 4, FS_X = TILE_X + 4;  // stress tile   (halo 2)
constexpr size_t FWD_SMEM = (size_t)(2 * FV_Z * FV_X + 3 * FS_Z * FS_X) * sizeof(float);

template <bool SAVE>
__global__ void __launch_bounds__(NTHREADS) fwd_step_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ float smem[];
  float *s_vz = smem;
  float *s_vx = s_vz + FV_Z * FV_X;
  float *s_zz = s_vx + FV_Z * FV_X;
  float *s_xx = s_zz + FS_Z * FS_X;
  float *s_xz = s_xx + FS_Z * FS_X;
  const Grid &g = a.g;
  const int tid = threadIdx.x;
  const int shot = blockIdx.x % a.batch;
  const int tile = blockIdx.x / a.batch;
  const int tz = tile % g.tiles_z, tx = tile / g.tiles_z;
  const int z0 = tz * TILE_Z, x0 = tx * TILE_X;
  const int ntiles = g.tiles_z * g.tiles_x;
  const int r0 = a.st.rec_ptr[shot * (ntiles + 1) + tile];
  const int r1 = a.st.rec_ptr[shot * (ntiles + 1) + tile + 1];
  const int sz = a.st.src_z[shot], sx = a.st.src_x[shot];
  const bool src_here = sz >= z0 - 2 && sz < z0 + TILE_Z + 2 && sx >= x0 - 2 && sx < x0 + TILE_X + 2;
  if (z0 - 2 > g.az_hi && r1 == r0 && !src_here) return;  // nothing ever changes in this tile

  const int fin = a.cur ? S_FB : S_FA, fout = a.cur ? S_FA : S_FB;
  const int pin = a.cur ? S_PSI_B : S_PSI_A, pout = a.cur ? S_PSI_A : S_PSI_B;
  const float *vz_i = plane_of(a.state, g, shot, fin + F_VZ);
  const float *vx_i = plane_of(a.state, g, shot, fin + F_VX);
  const float *szz_i = plane_of(a.state, g, shot, fin + F_SZZ);
  const float *sxx_i = plane_of(a.state, g, shot, fin + F_SXX);
  const float *sxz_i = plane_of(a.state, g, shot, fin + F_SXZ);
  float *vz_o = plane_of(a.state, g, shot, fout + F_VZ);
  float *vx_o = plane_of(a.state, g, shot, fout + F_VX);
  float *szz_o = plane_of(a.state, g, shot, fout + F_SZZ);
  float *sxx_o = plane_of(a.state, g, shot, fout + F_SXX);
  float *sxz_o = plane_of(a.state, g, shot, fout + F_SXZ);
  const int P = g.P;
  const int xmax = g.nx + XM - 1;
  const float dt = g.dt, rdz = g.rdz, rdx = g.rdx;

  // ---- phase 1: velocity tile with halo 3 -> shared ----
  for (int i = tid; i < FV_Z * FV_X; i += NTHREADS) {
    const int lx = i / FV_Z, lz = i - lx * FV_Z;
    const int gx = min(x0 - 3 + lx, xmax);
    const long long off = (long long)gx * P + (z0 - 3 + lz);
    s_vz[i] = vz_i[off];
    s_vx[i] = vx_i[off];
  }
  __syncthreads();

  const bool pml_tile = (z0 - 2 < g.nPml) || (z0 + TILE_Z + 1 > g.nz - g.nPml - g.nPad - 1) || (x0 - 2 < g.nPml) ||
                        (x0 + TILE_X + 1 > g.nx - g.nPml - 1);
  bool frame_tile = false;
  float *frm = nullptr;
  if (SAVE) {
    frame_tile = !(z0 > g.zhi + 2 || z0 + TILE_Z - 1 < g.zlo - 2 || x0 > g.xhi + 2 || x0 + TILE_X - 1 < g.xlo - 2) &&
                 !(z0 > g.zlo + 2 && z0 + TILE_Z - 1 < g.zhi - 2 && x0 > g.xlo + 2 && x0 + TILE_X - 1 < g.xhi - 2);
    frm = a.frames + ((long long)shot * g.nSteps + a.it) * 5 * g.f_len;
  }

  // ---- phase 2: stress on the tile + halo 2 ----
  for (int i = tid; i < FS_Z * FS_X; i += NTHREADS) {
    const int lx = i / FS_Z, lz = i - lx * FS_Z;
    const int gz = z0 - 2 + lz, gx = x0 - 2 + lx;
    const long long off = (long long)min(gx, xmax) * P + gz;
    float szz = szz_i[off], sxx = sxx_i[off], sxz = sxz_i[off];
    const bool owner = lz >= 2 && lz < TILE_Z + 2 && lx >= 2 && lx < TILE_X + 2 && gz < g.nz && gx < g.nx;
    const float *pz = s_vz + (lx + 1) * FV_Z + (lz + 1);
    const float *px = s_vx + (lx + 1) * FV_Z + (lz + 1);
    if (SAVE && frame_tile && owner) {
      const int fi = frame_index(g, gz, gx);
      if (fi >= 0) {
        frm[F_SZZ * g.f_len + fi] = szz;
        frm[F_SXX * g.f_len + fi] = sxx;
        frm[F_SXZ * g.f_len + fi] = sxz;
        frm[F_VZ * g.f_len + fi] = pz[0];
        frm[F_VX * g.f_len + fi] = px[0];
      }
    }
    if (is_active(g, gz, gx)) {
      float dvz_dz = d_minus(pz, 1, rdz);
      float dvx_dx = d_minus(px, FV_Z, rdx);
      float dvx_dz = d_plus(px, 1, rdz);
      float dvz_dx = d_plus(pz, FV_Z, rdx);
      if (pml_tile) {
        if (z_in_pml(g, gz)) {
          const float *zp = a.pr.z + gz;
          float m = zp[PR_B * P] * plane_of(a.state, g, shot, pin + PSI_VZ_Z)[off] + zp[PR_A * P] * dvz_dz;
          dvz_dz = dvz_dz * zp[PR_RK * P] + m;
          float mh = zp[PR_BH * P] * plane_of(a.state, g, shot, pin + PSI_VX_Z)[off] + zp[PR_AH * P] * dvx_dz;
          dvx_dz = dvx_dz * zp[PR_RKH * P] + mh;
          if (owner) {
            plane_of(a.state, g, shot, pout + PSI_VZ_Z)[off] = m;
            plane_of(a.state, g, shot, pout + PSI_VX_Z)[off] = mh;
          }
        }
        if (x_in_pml_s(g, gx)) {
          const float *xp = a.pr.x + gx + XM;
          const int n = a.pr.nxp;
          float m = xp[PR_B * n] * plane_of(a.state, g, shot, pin + PSI_VX_X)[off] + xp[PR_A * n] * dvx_dx;
          dvx_dx = dvx_dx * xp[PR_RK * n] + m;
          float mh = xp[PR_BH * n] * plane_of(a.state, g, shot, pin + PSI_VZ_X)[off] + xp[PR_AH * n] * dvz_dx;
          dvz_dx = dvz_dx * xp[PR_RKH * n] + mh;
          if (owner) {
            plane_of(a.state, g, shot, pout + PSI_VX_X)[off] = m;
            plane_of(a.state, g, shot, pout + PSI_VZ_X)[off] = mh;
          }
        }
      }
      const float lam = a.m.lam[off], mu = a.m.mu[off], amu = a.m.amu[off];
      szz = stress_inc(szz, lam, mu, dvz_dz, dvx_dx, dt, 1.0f);
      sxx = stress_inc(sxx, lam, mu, dvx_dx, dvz_dz, dt, 1.0f);
      sxz = sxz + amu * (dvx_dz + dvz_dx) * dt;
    }
    if (gz == sz && gx == sx) {  // add_source (utilities.cu:521-537), point stamp
      const float amp = a.st.stf[shot * g.nSteps + a.it];
      szz += SRC_SCALE * amp * dt;
      sxx = (float)((double)sxx + 3.0 * (double)SRC_SCALE * (double)amp * (double)dt);
    }
    s_zz[i] = szz;
    s_xx[i] = sxx;
    s_xz[i] = sxz;
    if (owner) {
      szz_o[off] = szz;
      sxx_o[off] = sxx;
      sxz_o[off] = sxz;
    }
  }
  __syncthreads();

  // ---- recording at time index it+1 (utilities.cu:557-567) ----
  for (int k = r0 + tid; k < r1; k += NTHREADS) {
    const int loc = a.st.rec_loc[shot * a.st.nrp + k];
    const int lz = loc & 0xffff, lx = loc >> 16;
    const int j = (lx + 2) * FS_Z + lz + 2;
    a.traces[((long long)shot * g.nSteps + a.it + 1) * a.st.nrp + a.st.rec_id[shot * a.st.nrp + k]] =
        (float)((double)s_zz[j] + 3.0 * (double)s_xx[j]);
  }

  // ---- phase 3: velocity on the owner tile ----
  for (int i = tid; i < TILE_Z * TILE_X; i += NTHREADS) {
    const int lx = i / TILE_Z, lz = i - lx * TILE_Z;
    const int gz = z0 + lz, gx = x0 + lx;
    if (gz >= g.nz || gx >= g.nx) continue;
    const long long off = (long long)gx * P + gz;
    float vz = s_vz[(lx + 3) * FV_Z + lz + 3], vx = s_vx[(lx + 3) * FV_Z + lz + 3];
    if (is_active(g, gz, gx)) {
      const float *zz = s_zz + (lx + 2) * FS_Z + lz + 2;
      const float *xx = s_xx + (lx + 2) * FS_Z + lz + 2;
      const float *xz = s_xz + (lx + 2) * FS_Z + lz + 2;
      float dszz_dz = d_plus(zz, 1, rdz);
      float dsxz_dx = d_minus(xz, FS_Z, rdx);
      float dsxz_dz = d_minus(xz, 1, rdz);
      float dsxx_dx = d_plus(xx, FS_Z, rdx);
      if (pml_tile) {
        if (z_in_pml(g, gz)) {
          const float *zp = a.pr.z + gz;
          float *q1 = plane_of(a.state, g, shot, S_PHI_A + PHI_SZZ_Z) + off;
          float *q2 = plane_of(a.state, g, shot, S_PHI_A + PHI_SXZ_Z) + off;
          float m = zp[PR_BH * P] * (*q1) + zp[PR_AH * P] * dszz_dz;
          *q1 = m;
          dszz_dz = dszz_dz * zp[PR_RKH * P] + m;
          float m2 = zp[PR_B * P] * (*q2) + zp[PR_A * P] * dsxz_dz;
          *q2 = m2;
          dsxz_dz = dsxz_dz * zp[PR_RK * P] + m2;
        }
        if (x_in_pml_v(g, gx)) {
          const float *xp = a.pr.x + gx + XM;
          const int n = a.pr.nxp;
          float *q1 = plane_of(a.state, g, shot, S_PHI_A + PHI_SXZ_X) + off;
          float *q2 = plane_of(a.state, g, shot, S_PHI_A + PHI_SXX_X) + off;
          float m = xp[PR_B * n] * (*q1) + xp[PR_A * n] * dsxz_dx;
          *q1 = m;
          dsxz_dx = dsxz_dx * xp[PR_RK * n] + m;
          float m2 = xp[PR_BH * n] * (*q2) + xp[PR_AH * n] * dsxx_dx;
          *q2 = m2;
          dsxx_dx = dsxx_dx * xp[PR_RKH * n] + m2;
        }
      }
      vz += (dszz_dz + dsxz_dx) * a.m.bya[off] * dt;
      vx += (dsxz_dz + dsxx_dx) * a.m.byb[off] * dt;
    }
    vz_o[off] = vz;
    vx_o[off] = vx;
  }
}

// =================================================================================================
// reverse-time reconstruction + imaging condition
// =================================================================================================
constexpr int RS_Z = TILE_Z + 8, RS_X = TILE_X + 8;  // sigma^{it+1} tile (halo 4)
constexpr int RV_Z = TILE_Z + 4, RV_X = TILE_X + 4;  // v^{it}, ga, gb tile (halo 2)
constexpr int RG_Z = TILE_Z + 1, RG_X = TILE_X + 1;  // mu-spray tile (halo 1 on the low side)
constexpr size_t REV_SMEM = (size_t)(3 * RS_Z * RS_X + 4 * RV_Z * RV_X + RG_Z * RG_X) * sizeof(float);

__global__ void __launch_bounds__(NTHREADS) rev_image_kernel(const __grid_constant__ BwdArgs a, int tz_first,
                                                             int tx_first, int ntz) {
  extern __shared__ float smem[];
  float *s_zz = smem;
  float *s_xx = s_zz + RS_Z * RS_X;
  float *s_xz = s_xx + RS_Z * RS_X;
  float *s_vz = s_xz + RS_Z * RS_X;
  float *s_vx = s_vz + RV_Z * RV_X;
  float *s_ga = s_vx + RV_Z * RV_X;
  float *s_gb = s_ga + RV_Z * RV_X;
  float *s_sp = s_gb + RV_Z * RV_X;
  const Grid &g = a.g;
  const int tid = threadIdx.x;
  const int shot = blockIdx.x % a.batch;
  const int tile = blockIdx.x / a.batch;
  const int tz = tz_first + tile % ntz, tx = tx_first + tile / ntz;
  const int z0 = tz * TILE_Z, x0 = tx * TILE_X;
  const int P = g.P;
  const int xmax = g.nx + XM - 1;
  const float dt = g.dt, rdz = g.rdz, rdx = g.rdx;
  const int fin = a.cur_f ? S_FB : S_FA, fout = a.cur_f ? S_FA : S_FB;
  const int ain = a.cur_a ? S_AB : S_AA;
  const float *szz_i = plane_of(a.state, g, shot, fin + F_SZZ);
  const float *sxx_i = plane_of(a.state, g, shot, fin + F_SXX);
  const float *sxz_i = plane_of(a.state, g, shot, fin + F_SXZ);
  const float *vz_i = plane_of(a.state, g, shot, fin + F_VZ);
  const float *vx_i = plane_of(a.state, g, shot, fin + F_VX);
  const float *frm = a.frames + ((long long)shot * g.nSteps + a.it) * 5 * g.f_len;
  const int sz = a.st.src_z[shot], sx = a.st.src_x[shot];

  // ---- phase 1: sigma^{it+1} with halo 4 ----
  for (int i = tid; i < RS_Z * RS_X; i += NTHREADS) {
    const int lx = i / RS_Z, lz = i - lx * RS_Z;
    const int gx = min(x0 - 4 + lx, xmax);
    const long long off = (long long)gx * P + (z0 - 4 + lz);
    s_zz[i] = szz_i[off];
    s_xx[i] = sxx_i[off];
    s_xz[i] = sxz_i[off];
  }
  __syncthreads();

  // does this tile (+halo 2) touch the saved frames?
  const bool frame_tile =
      !(z0 - 2 > g.zlo + 2 && z0 + TILE_Z + 1 < g.zhi - 2 && x0 - 2 > g.xlo + 2 && x0 + TILE_X + 1 < g.xhi - 2);

  // ---- phase 2: v^{it} on tile + halo 2, density imaging terms ----
  {
    const float *vza = plane_of(a.state, g, shot, ain + F_VZ);
    const float *vxa = plane_of(a.state, g, shot, ain + F_VX);
    float *vz_o = plane_of(a.state, g, shot, fout + F_VZ);
    float *vx_o = plane_of(a.state, g, shot, fout + F_VX);
    for (int i = tid; i < RV_Z * RV_X; i += NTHREADS) {
      const int lx = i / RV_Z, lz = i - lx * RV_Z;
      const int gz = z0 - 2 + lz, gx = x0 - 2 + lx;
      const long long off = (long long)min(gx, xmax) * P + gz;
      float vz = vz_i[off], vx = vx_i[off];
      float ga = 0.0f, gb = 0.0f;
      const bool box = in_box(g, gz, gx);
      if (box) {
        const float *zz = s_zz + (lx + 2) * RS_Z + lz + 2;
        const float *xx = s_xx + (lx + 2) * RS_Z + lz + 2;
        const float *xz = s_xz + (lx + 2) * RS_Z + lz + 2;
        const float ea = d_plus(zz, 1, rdz) + d_minus(xz, RS_Z, rdx);
        const float eb = d_minus(xz, 1, rdz) + d_plus(xx, RS_Z, rdx);
        const float bya = a.m.bya[off], byb = a.m.byb[off];
        vz -= ea * bya * dt;
        vx -= eb * byb * dt;
        // el_velocity.cu:101-104
        ga = (float)((double)(-vza[off] * ea * dt) * (-((double)bya * (double)bya) / 2.0));
        gb = (float)((double)(-vxa[off] * eb * dt) * (-((double)byb * (double)byb) / 2.0));
      }
      int fi = -1;
      if (frame_tile) {
        fi = frame_index(g, gz, gx);
        if (fi >= 0) {
          vz = frm[F_VZ * g.f_len + fi];
          vx = frm[F_VX * g.f_len + fi];
        }
      }
      s_vz[i] = vz;
      s_vx[i] = vx;
      s_ga[i] = ga;
      s_gb[i] = gb;
      const bool owner = lz >= 2 && lz < TILE_Z + 2 && lx >= 2 && lx < TILE_X + 2;
      if (owner && (box || fi >= 0)) {
        vz_o[off] = vz;
        vx_o[off] = vx;
      }
    }
  }
  __syncthreads();

  // ---- phase 3a: mu "spray" amplitude of every source cell on [z0-1, z0+TZ) x [x0-1, x0+TX) ----
  {
    const float *sxza = plane_of(a.state, g, shot, ain + F_SXZ);
    for (int i = tid; i < RG_Z * RG_X; i += NTHREADS) {
      const int lx = i / RG_Z, lz = i - lx * RG_Z;
      const int gz = z0 - 1 + lz, gx = x0 - 1 + lx;
      float sp = 0.0f;
      if (in_box(g, gz, gx)) {
        const long long off = (long long)gx * P + gz;
        const float amu = a.m.amu[off];
        if (amu != 0.0f) {
          const float *pz = s_vz + (lx + 1) * RV_Z + lz + 1;
          const float *px = s_vx + (lx + 1) * RV_Z + lz + 1;
          const float e = d_plus(px, 1, rdz) + d_plus(pz, RV_Z, rdx);
          // el_stress.cu:114-116 with  amu / sum(1/mu) == amu^2 / 4  (amu = 4 / sum(1/mu))
          sp = -sxza[off] * e * dt * (250000.0f * amu) * amu;
        }
      }
      s_sp[i] = sp;
    }
  }
  __syncthreads();

  // ---- phase 3b: sigma^{it} on the owner tile, lambda / mu / rho accumulation (gather) ----
  {
    const float *szza = plane_of(a.state, g, shot, ain + F_SZZ);
    const float *sxxa = plane_of(a.state, g, shot, ain + F_SXX);
    float *szz_o = plane_of(a.state, g, shot, fout + F_SZZ);
    float *sxx_o = plane_of(a.state, g, shot, fout + F_SXX);
    float *sxz_o = plane_of(a.state, g, shot, fout + F_SXZ);
    float *gl = a.gacc + ((long long)shot * 3 + 0) * g.plane + g.origin;
    float *gm = a.gacc + ((long long)shot * 3 + 1) * g.plane + g.origin;
    float *gd = a.gacc + ((long long)shot * 3 + 2) * g.plane + g.origin;
    for (int i = tid; i < TILE_Z * TILE_X; i += NTHREADS) {
      const int lx = i / TILE_Z, lz = i - lx * TILE_Z;
      const int gz = z0 + lz, gx = x0 + lx;
      if (gz >= g.nz || gx >= g.nx) continue;
      const long long off = (long long)gx * P + gz;
      const bool box = in_box(g, gz, gx);
      int fi = frame_tile ? frame_index(g, gz, gx) : -1;
      if (box) {
        const float *pz = s_vz + (lx + 2) * RV_Z + lz + 2;
        const float *px = s_vx + (lx + 2) * RV_Z + lz + 2;
        const float dvz_dz = d_minus(pz, 1, rdz);
        const float dvx_dx = d_minus(px, RV_Z, rdx);
        const float dvx_dz = d_plus(px, 1, rdz);
        const float dvz_dx = d_plus(pz, RV_Z, rdx);
        const int j = (lx + 4) * RS_Z + lz + 4;
        float szz = s_zz[j], sxx = s_xx[j], sxz = s_xz[j];
        if (gz == sz && gx == sx) {  // add_source(isFor=false): utilities.cu:538-551
          const float amp = a.st.stf[shot * g.nSteps + a.it];
          szz -= SRC_SCALE * amp * dt;
          sxx = (float)((double)sxx - 3.0 * (double)SRC_SCALE * (double)amp * (double)dt);
        }
        const float lam = a.m.lam[off], mu = a.m.mu[off];
        szz = stress_inc(szz, lam, mu, dvz_dz, dvx_dx, dt, -1.0f);
        sxx = stress_inc(sxx, lam, mu, dvx_dx, dvz_dz, dt, -1.0f);
        sxz -= a.m.amu[off] * (dvx_dz + dvz_dx) * dt;
        if (fi >= 0) {
          szz = frm[F_SZZ * g.f_len + fi];
          sxx = frm[F_SXX * g.f_len + fi];
          sxz = frm[F_SXZ * g.f_len + fi];
        }
        szz_o[off] = szz;
        sxx_o[off] = sxx;
        sxz_o[off] = sxz;
        // el_stress.cu:109-111
        const float za = szza[off], xa = sxxa[off];
        gl[off] = (float)((double)gl[off] + (double)(-(za + xa) * (dvz_dz + dvx_dx) * dt) * 1e6);
        const double gm_dir =
            (-2.0 * (double)za * (double)dvz_dz * (double)dt - 2.0 * (double)xa * (double)dvx_dx * (double)dt) * 1e6;
        float gmv = (float)((double)gm[off] + gm_dir);
        const int q = (lx + 1) * RG_Z + lz + 1;
        const float G = s_sp[q] + s_sp[q - 1] + s_sp[q - RG_Z] + s_sp[q - RG_Z - 1];
        if (G != 0.0f) gmv += G / (mu * mu);
        gm[off] = gmv;
        const int v = (lx + 2) * RV_Z + lz + 2;
        gd[off] += s_ga[v] + s_gb[v] + s_ga[v - 1] + s_gb[v - RV_Z];
      } else {
        if (fi >= 0) {
          szz_o[off] = frm[F_SZZ * g.f_len + fi];
          sxx_o[off] = frm[F_SXX * g.f_len + fi];
          sxz_o[off] = frm[F_SXZ * g.f_len + fi];
        }
        // column xhi+1 receives the x+1 spray of the last box column (el_stress.cu:120, el_velocity.cu:109)
        if (gx == g.xhi + 1 && gz >= g.zlo && gz <= g.zhi) {
          const int q = (lx + 1) * RG_Z + lz + 1;
          const float G = s_sp[q - RG_Z];
          const float mu = a.m.mu[off];
          if (G != 0.0f) gm[off] += G / (mu * mu);
          const int v = (lx + 2) * RV_Z + lz + 2;
          gd[off] += s_gb[v - RV_Z];
        }
      }
    }
  }
}

// =================================================================================================
// adjoint step
// =================================================================================================
constexpr int AS_Z = TILE_Z + 6, AS_X = TILE_X + 6;  // adjoint stress tile (halo 3)
constexpr int AV_Z = TILE_Z + 4, AV_X = TILE_X + 4;  // adjoint velocity tile (halo 2)
constexpr size_t ADJ_SMEM = (size_t)(3 * AS_Z * AS_X + 2 * AV_Z * AV_X) * sizeof(float);

__global__ void __launch_bounds__(NTHREADS) adj_step_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ float smem[];
  float *s_zz = smem;
  float *s_xx = s_zz + AS_Z * AS_X;
  float *s_xz = s_xx + AS_Z * AS_X;
  float *s_vz = s_xz + AS_Z * AS_X;
  float *s_vx = s_vz + AV_Z * AV_X;
  const Grid &g = a.g;
  const int tid = threadIdx.x;
  const int shot = blockIdx.x % a.batch;
  const int tile = blockIdx.x / a.batch;
  const int tz = tile % g.tiles_z, tx = tile / g.tiles_z;
  const int z0 = tz * TILE_Z, x0 = tx * TILE_X;
  const int ntiles = g.tiles_z * g.tiles_x;
  const int r0 = a.st.rec_ptr[shot * (ntiles + 1) + tile];
  const int r1 = a.st.rec_ptr[shot * (ntiles + 1) + tile + 1];
  const int sz = a.st.src_z[shot], sx = a.st.src_x[shot];
  const bool src_owner = sz >= z0 && sz < z0 + TILE_Z && sx >= x0 && sx < x0 + TILE_X;
  if (z0 - 2 > g.az_hi && r1 == r0 && !src_owner) return;

  const int P = g.P;
  const int xmax = g.nx + XM - 1;
  const float dt = g.dt, rdz = g.rdz, rdx = g.rdx;
  const int ain = a.cur_a ? S_AB : S_AA, aout = a.cur_a ? S_AA : S_AB;
  const int pin = a.cur_a ? S_PSI_B : S_PSI_A, pout = a.cur_a ? S_PSI_A : S_PSI_B;
  const int qin = a.cur_a ? S_PHI_B : S_PHI_A, qout = a.cur_a ? S_PHI_A : S_PHI_B;
  const float *szz_i = plane_of(a.state, g, shot, ain + F_SZZ);
  const float *sxx_i = plane_of(a.state, g, shot, ain + F_SXX);
  const float *sxz_i = plane_of(a.state, g, shot, ain + F_SXZ);
  const float *vz_i = plane_of(a.state, g, shot, ain + F_VZ);
  const float *vx_i = plane_of(a.state, g, shot, ain + F_VX);
  const int nxp = a.pr.nxp;

  // source_grad (utilities.cu:582-593): adjoint stress at the source BEFORE this step's update
  if (src_owner && tid == 0) {
    const long long off = (long long)sx * P + sz;
    a.stf_grad[shot * g.nSteps + a.it] = (float)(-((double)szz_i[off] + 3.0 * (double)sxx_i[off]) * (double)dt);
  }

  // ---- phase 1: adjoint stress tile with halo 3 ----
  for (int i = tid; i < AS_Z * AS_X; i += NTHREADS) {
    const int lx = i / AS_Z, lz = i - lx * AS_Z;
    const int gx = min(x0 - 3 + lx, xmax);
    const long long off = (long long)gx * P + (z0 - 3 + lz);
    s_zz[i] = szz_i[off];
    s_xx[i] = sxx_i[off];
    s_xz[i] = sxz_i[off];
  }
  __syncthreads();

  // psi arrays only matter within 2 cells of the PML
  const int zq_lo = g.nPml + 2, zq_hi = g.nz - g.nPad - g.nPml - 3;  // z-type psi zone: z < zq_lo || z > zq_hi
  const int xq_lo = g.nPml + 2, xq_hi = g.nx - g.nPml - 3;
  const bool pml_tile = (z0 - 2 < zq_lo) || (z0 + TILE_Z + 1 > zq_hi) || (x0 - 2 < xq_lo) || (x0 + TILE_X + 1 > xq_hi);

  // ---- phase 2: adjoint velocity on tile + halo 2 (el_velocity_adj.cu:56-100) ----
  {
    float *vz_o = plane_of(a.state, g, shot, aout + F_VZ);
    float *vx_o = plane_of(a.state, g, shot, aout + F_VX);
    for (int i = tid; i < AV_Z * AV_X; i += NTHREADS) {
      const int lx = i / AV_Z, lz = i - lx * AV_Z;
      const int gz = z0 - 2 + lz, gx = x0 - 2 + lx;
      const long long off = (long long)min(gx, xmax) * P + gz;
      float vz = vz_i[off], vx = vx_i[off];
      const bool owner = lz >= 2 && lz < TILE_Z + 2 && lx >= 2 && lx < TILE_X + 2 && gz < g.nz && gx < g.nx;
      if (is_active(g, gz, gx)) {
        const float *zz = s_zz + (lx + 1) * AS_Z + lz + 1;
        const float *xx = s_xx + (lx + 1) * AS_Z + lz + 1;
        const float *xz = s_xz + (lx + 1) * AS_Z + lz + 1;
        const float lam = a.m.lam[off], mu = a.m.mu[off], amu = a.m.amu[off];
        const float dszz_dx = ad_plus(zz, AS_Z, rdx);
        const float dsxx_dx = ad_plus(xx, AS_Z, rdx);
        const float dsxz_dz = ad_minus(xz, 1, rdz);
        const float dszz_dz = ad_plus(zz, 1, rdz);
        const float dsxx_dz = ad_plus(xx, 1, rdz);
        const float dsxz_dx = ad_minus(xz, AS_Z, rdx);
        float rKx = 1.0f, rKxh = 1.0f, rKz = 1.0f, rKzh = 1.0f;
        float tpx1 = 0.0f, tpx2 = 0.0f, tpz1 = 0.0f, tpz2 = 0.0f;  // a * D(psi) terms
        bool zp = false, xp = false;
        if (pml_tile) {
          const float *zpf = a.pr.z + gz;
          const float *xpf = a.pr.x + gx + XM;
          zp = z_in_pml(g, gz);
          xp = x_in_pml_s(g, gx);
          rKx = xpf[PR_RK * nxp];
          rKxh = xpf[PR_RKH * nxp];
          rKz = zpf[PR_RK * P];
          rKzh = zpf[PR_RKH * P];
          const float ax = xpf[PR_A * nxp], axh = xpf[PR_AH * nxp], az = zpf[PR_A * P], azh = zpf[PR_AH * P];
          if (ax != 0.0f) tpx1 = ax * ad_plus(plane_of(a.state, g, shot, pin + PSI_VX_X) + off, P, rdx);
          if (azh != 0.0f) tpx2 = azh * ad_minus(plane_of(a.state, g, shot, pin + PSI_VX_Z) + off, 1, rdz);
          if (az != 0.0f) tpz1 = az * ad_plus(plane_of(a.state, g, shot, pin + PSI_VZ_Z) + off, 1, rdz);
          if (axh != 0.0f) tpz2 = axh * ad_minus(plane_of(a.state, g, shot, pin + PSI_VZ_X) + off, P, rdx);
        }
#if FWI_FP64_PROMOTE
        const double l2m = (double)lam + 2.0 * (double)mu;
        {
          const float t12 = tpx1 + lam * dszz_dx * rKx * dt;
          const double t3 = l2m * (double)dsxx_dx * (double)rKx * (double)dt;
          const double sum = (double)t12 + t3 + (double)tpx2 + (double)(amu * rKzh * dsxz_dz * dt);
          vx = (float)((double)vx + sum);
        }
        {
          const double t2 = l2m * (double)dszz_dz * (double)rKz * (double)dt;
          const double sum = (double)tpz1 + t2 + (double)(lam * dsxx_dz * rKz * dt) + (double)tpz2 +
                             (double)(amu * rKxh * dsxz_dx * dt);
          vz = (float)((double)vz + sum);
        }
#else
        const float l2m = lam + 2.0f * mu;
        vx += tpx1 + lam * dszz_dx * rKx * dt + l2m * dsxx_dx * rKx * dt + tpx2 + amu * rKzh * dsxz_dz * dt;
        vz += tpz1 + l2m * dszz_dz * rKz * dt + lam * dsxx_dz * rKz * dt + tpz2 + amu * rKxh * dsxz_dx * dt;
#endif
        if (owner && (xp || zp)) {  // phi memory, PML only (el_velocity_adj.cu:74-79,95-100)
          const float bya = a.m.bya[off], byb = a.m.byb[off];
          if (xp) {
            const float *xpf = a.pr.x + gx + XM;
            plane_of(a.state, g, shot, qout + PHI_SXX_X)[off] =
                xpf[PR_BH * nxp] * plane_of(a.state, g, shot, qin + PHI_SXX_X)[off] + byb * vx * dt;
            plane_of(a.state, g, shot, qout + PHI_SXZ_X)[off] =
                xpf[PR_B * nxp] * plane_of(a.state, g, shot, qin + PHI_SXZ_X)[off] + bya * vz * dt;
          }
          if (zp) {
            const float *zpf = a.pr.z + gz;
            plane_of(a.state, g, shot, qout + PHI_SXZ_Z)[off] =
                zpf[PR_B * P] * plane_of(a.state, g, shot, qin + PHI_SXZ_Z)[off] + byb * vx * dt;
            plane_of(a.state, g, shot, qout + PHI_SZZ_Z)[off] =
                zpf[PR_BH * P] * plane_of(a.state, g, shot, qin + PHI_SZZ_Z)[off] + bya * vz * dt;
          }
        }
      }
      s_vz[i] = vz;
      s_vx[i] = vx;
      if (owner) {
        vz_o[off] = vz;
        vx_o[off] = vx;
      }
    }
  }
  __syncthreads();

  // ---- residual injection at time index `it` (utilities.cu:569-580), owner cells only ----
  for (int k = r0 + tid; k < r1; k += NTHREADS) {
    const int loc = a.st.rec_loc[shot * a.st.nrp + k];
    const int lz = loc & 0xffff, lx = loc >> 16;
    const float r = a.res[((long long)shot * g.nSteps + a.it) * a.st.nrp + a.st.rec_id[shot * a.st.nrp + k]];
    const int j = (lx + 3) * AS_Z + lz + 3;
    atomicAdd(&s_zz[j], r);
    atomicAdd(&s_xx[j], 3.0f * r);
  }
  __syncthreads();

  // ---- phase 3: adjoint stress on the owner tile (el_stress_adj.cu:52-95) ----
  {
    float *szz_o = plane_of(a.state, g, shot, aout + F_SZZ);
    float *sxx_o = plane_of(a.state, g, shot, aout + F_SXX);
    float *sxz_o = plane_of(a.state, g, shot, aout + F_SXZ);
    for (int i = tid; i < TILE_Z * TILE_X; i += NTHREADS) {
      const int lx = i / TILE_Z, lz = i - lx * TILE_Z;
      const int gz = z0 + lz, gx = x0 + lx;
      if (gz >= g.nz || gx >= g.nx) continue;
      const long long off = (long long)gx * P + gz;
      const int j = (lx + 3) * AS_Z + lz + 3;
      float szz = s_zz[j], sxx = s_xx[j], sxz = s_xz[j];
      if (is_active(g, gz, gx)) {
        const float *pz = s_vz + (lx + 2) * AV_Z + lz + 2;
        const float *px = s_vx + (lx + 2) * AV_Z + lz + 2;
        const float dvz_dx = ad_plus(pz, AV_Z, rdx);
        const float dvx_dz = ad_plus(px, 1, rdz);
        const float dvx_dx = ad_minus(px, AV_Z, rdx);
        const float dvz_dz = ad_minus(pz, 1, rdz);
        const float bya = a.m.bya[off], byb = a.m.byb[off];
        float rKx = 1.0f, rKxh = 1.0f, rKz = 1.0f, rKzh = 1.0f;
        float t_xz_x = 0.0f, t_xz_z = 0.0f, t_xx = 0.0f, t_zz = 0.0f;  // a * D(phi_new) terms
        if (pml_tile) {
          const float *zpf = a.pr.z + gz;
          const float *xpf = a.pr.x + gx + XM;
          rKx = xpf[PR_RK * nxp];
          rKxh = xpf[PR_RKH * nxp];
          rKz = zpf[PR_RK * P];
          rKzh = zpf[PR_RKH * P];
          const float ax = xpf[PR_A * nxp], axh = xpf[PR_AH * nxp], az = zpf[PR_A * P], azh = zpf[PR_AH * P];
          // phi_new at a stencil point, recomputed from the velocity tile instead of being staged:
          //   phi_new = active & in-PML ? b * phi_old + byc * v_new * dt : phi_old
          auto phi_x = [&](int which, int dxs, const float *bprof, const float *byc, const float *sv) -> float {
            const int x2 = gx + dxs;
            const long long o2 = off + (long long)dxs * P;
            float ph = plane_of(a.state, g, shot, qin + which)[o2];
            if (is_active(g, gz, x2) && x_in_pml_s(g, x2)) ph = bprof[x2 + XM] * ph + byc[o2] * sv[dxs * AV_Z] * dt;
            return ph;
          };
          auto phi_z = [&](int which, int dzs, const float *bprof, const float *byc, const float *sv) -> float {
            const int z2 = gz + dzs;
            const long long o2 = off + dzs;
            float ph = plane_of(a.state, g, shot, qin + which)[o2];
            if (is_active(g, z2, gx) && z_in_pml(g, z2)) ph = bprof[z2] * ph + byc[o2] * sv[dzs] * dt;
            return ph;
          };
          if (ax != 0.0f) {  // D+x of phi_xz_x (b_x, byc_a * vz)
            const float *bp = a.pr.x + PR_B * nxp;
            t_xz_x = ax * ad_plus4(phi_x(PHI_SXZ_X, -1, bp, a.m.bya, pz), phi_x(PHI_SXZ_X, 0, bp, a.m.bya, pz),
                                   phi_x(PHI_SXZ_X, 1, bp, a.m.bya, pz), phi_x(PHI_SXZ_X, 2, bp, a.m.bya, pz), rdx);
          }
          if (az != 0.0f) {  // D+z of phi_xz_z (b_z, byc_b * vx)
            const float *bp = a.pr.z + PR_B * P;
            t_xz_z = az * ad_plus4(phi_z(PHI_SXZ_Z, -1, bp, a.m.byb, px), phi_z(PHI_SXZ_Z, 0, bp, a.m.byb, px),
                                   phi_z(PHI_SXZ_Z, 1, bp, a.m.byb, px), phi_z(PHI_SXZ_Z, 2, bp, a.m.byb, px), rdz);
          }
          if (axh != 0.0f) {  // D-x of phi_xx_x (b_x_half, byc_b * vx)
            const float *bp = a.pr.x + PR_BH * nxp;
            t_xx = axh * ad_minus4(phi_x(PHI_SXX_X, -2, bp, a.m.byb, px), phi_x(PHI_SXX_X, -1, bp, a.m.byb, px),
                                   phi_x(PHI_SXX_X, 0, bp, a.m.byb, px), phi_x(PHI_SXX_X, 1, bp, a.m.byb, px), rdx);
          }
          if (azh != 0.0f) {  // D-z of phi_zz_z (b_z_half, byc_a * vz)
            const float *bp = a.pr.z + PR_BH * P;
            t_zz = azh * ad_minus4(phi_z(PHI_SZZ_Z, -2, bp, a.m.bya, pz), phi_z(PHI_SZZ_Z, -1, bp, a.m.bya, pz),
                                   phi_z(PHI_SZZ_Z, 0, bp, a.m.bya, pz), phi_z(PHI_SZZ_Z, 1, bp, a.m.bya, pz), rdz);
          }
        }
        sxz += t_xz_x + dvz_dx * rKx * bya * dt + t_xz_z + dvx_dz * rKz * byb * dt;
        sxx += t_xx + byb * dvx_dx * rKxh * dt;
        szz += t_zz + bya * dvz_dz * rKzh * dt;
        if (pml_tile) {
          const bool xq = gx < xq_lo || gx > xq_hi;
          const bool zq = gz < zq_lo || gz > zq_hi;
          if (xq || zq) {
            const float lam = a.m.lam[off], mu = a.m.mu[off], amu = a.m.amu[off];
            const float *zpf = a.pr.z + gz;
            const float *xpf = a.pr.x + gx + XM;
            const double l2m = (double)lam + 2.0 * (double)mu;
            if (xq) {
              plane_of(a.state, g, shot, pout + PSI_VZ_X)[off] =
                  xpf[PR_BH * nxp] * plane_of(a.state, g, shot, pin + PSI_VZ_X)[off] + sxz * amu * dt;
              plane_of(a.state, g, shot, pout + PSI_VX_X)[off] =
                  (float)((double)(xpf[PR_B * nxp] * plane_of(a.state, g, shot, pin + PSI_VX_X)[off] +
                                   lam * szz * dt) +
                          l2m * (double)sxx * (double)dt);
            }
            if (zq) {
              plane_of(a.state, g, shot, pout + PSI_VX_Z)[off] =
                  zpf[PR_BH * P] * plane_of(a.state, g, shot, pin + PSI_VX_Z)[off] + sxz * amu * dt;
              plane_of(a.state, g, shot, pout + PSI_VZ_Z)[off] =
                  (float)((double)(zpf[PR_B * P] * plane_of(a.state, g, shot, pin + PSI_VZ_Z)[off]) +
                          l2m * (double)szz * (double)dt + (double)(lam * sxx * dt));
            }
          }
        }
      }
      szz_o[off] = szz;
      sxx_o[off] = sxx;
      sxz_o[off] = sxz;
    }
  }
}

// =================================================================================================
// model preparation
// =================================================================================================
__global__ void model_transpose_kernel(Grid g, const double *__restrict__ lam_in, const double *__restrict__ mu_in,
                                       const double *__restrict__ den_in, float *lam, float *mu, float *den) {
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
      lam[o] = t[0][threadIdx.x][r];
      mu[o] = t[1][threadIdx.x][r];
      den[o] = t[2][threadIdx.x][r];
    }
  }
}

// mu_bar, averaged buoyancies (utilities.cu:125-152, Model.cu:67-73), max cp (utilities.cu:109-123)
__global__ void model_derive_kernel(Grid g, const float *lam, const float *mu, const float *den, float *amu, float *bya,
                                    float *byb, unsigned int *cpmax_bits) {
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
    amu[o] = m;
    bya[o] = ba;
    byb[o] = bb;
    cp = (float)sqrt(((double)lam[o] + 2.0 * (double)mu[o]) / (double)den[o]);
    if (!(cp > 0.0f)) cp = 0.0f;  // NaN / negative never wins the max
  }
  for (int s = 16; s > 0; s >>= 1) cp = fmaxf(cp, __shfl_xor_sync(0xffffffffu, cp, s));
  if ((threadIdx.x & 31) == 0 && cp > 0.0f) atomicMax(cpmax_bits, __float_as_uint(cp));
}

// =================================================================================================
// residual / misfit
// =================================================================================================
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
      const float w = a.w2[t];
      oc = t_obs[threadIdx.x][q] * w;                   // cuda_window on obs  (libCUFD.cu:268)
      sc = a.syn_tr[(long long)t * a.nrp + rec] * w;    // cuda_window on syn  (libCUFD.cu:270)
      res = (t > 0) ? oc - sc : 0.0f;                   // gpuMinus            (utilities.cu:154-167)
      acc += (double)(res * res);                       // cuda_cal_objective  (utilities.cu:169-205)
      res *= w;                                         // cuda_window on res  (libCUFD.cu:312)
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

// result planes are row-major [z][x] (libCUFD.cu:480-486); sums the per-slot accumulators in slot order
__global__ void finalize_kernel(Grid g, const float *gacc, int nslots, const float *misfit_half, float *result) {
  __shared__ float t[32][33];
  const int zb = blockIdx.x * 32, xb = blockIdx.y * 32, k = blockIdx.z;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = xb + r, z = zb + threadIdx.x;
    float s = 0.0f;
    if (x < g.nx && z < g.nz)
      for (int q = 0; q < nslots; q++) s += gacc[((long long)q * 3 + k) * g.plane + g.origin + (long long)x * g.P + z];
    t[r][threadIdx.x] = s;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int z = zb + r, x = xb + threadIdx.x;
    if (z < g.nz && x < g.nx) result[((long long)k * g.nz + z) * g.nx + x] = t[threadIdx.x][r];
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && k == 0 && threadIdx.x == 0 && threadIdx.y == 0)
    result[3LL * g.nz * g.nx] = *misfit_half;
}

}  // namespace

// =================================================================================================
// launchers
// =================================================================================================
size_t forward_smem_bytes() { return FWD_SMEM; }
size_t reverse_smem_bytes() { return REV_SMEM; }
size_t adjoint_smem_bytes() { return ADJ_SMEM; }

void configure_kernels() {
  cudaFuncSetAttribute(fwd_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
  cudaFuncSetAttribute(fwd_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
  cudaFuncSetAttribute(rev_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REV_SMEM);
  cudaFuncSetAttribute(adj_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ADJ_SMEM);
}

void launch_forward_step(const FwdArgs &a, bool save_frames, cudaStream_t s) {
  const int blocks = a.batch * a.g.tiles_z * a.g.tiles_x;
  if (save_frames)
    fwd_step_kernel<true><<<blocks, NTHREADS, FWD_SMEM, s>>>(a);
  else
    fwd_step_kernel<false><<<blocks, NTHREADS, FWD_SMEM, s>>>(a);
}

void launch_reverse_imaging(const BwdArgs &a, cudaStream_t s) {
  const Grid &g = a.g;
  const int tz0 = max(g.zlo - 2, 0) / TILE_Z, tz1 = min(g.zhi + 2, g.nz - 1) / TILE_Z;
  const int tx0 = max(g.xlo - 2, 0) / TILE_X, tx1 = min(g.xhi + 2, g.nx - 1) / TILE_X;
  const int ntz = tz1 - tz0 + 1, ntx = tx1 - tx0 + 1;
  rev_image_kernel<<<a.batch * ntz * ntx, NTHREADS, REV_SMEM, s>>>(a, tz0, tx0, ntz);
}

void launch_adjoint_step(const BwdArgs &a, cudaStream_t s) {
  const int blocks = a.batch * a.g.tiles_z * a.g.tiles_x;
  adj_step_kernel<<<blocks, NTHREADS, ADJ_SMEM, s>>>(a);
}

void launch_model_prep(const Grid &g, const double *d_lam, const double *d_mu, const double *d_den, float *lam,
                       float *mu, float *den, float *amu, float *bya, float *byb, unsigned int *cpmax_bits,
                       cudaStream_t s) {
  dim3 tb(32, 8);
  dim3 tg((g.nx + 31) / 32, (g.nz + 31) / 32);
  model_transpose_kernel<<<tg, tb, 0, s>>>(g, d_lam, d_mu, d_den, lam, mu, den);
  dim3 dg((g.nz + 127) / 128, g.nx);
  model_derive_kernel<<<dg, 128, 0, s>>>(g, lam, mu, den, amu, bya, byb, cpmax_bits);
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

void launch_finalize(const Grid &g, const float *gacc, int nslots, const float *misfit_half, float *result,
                     cudaStream_t s) {
  dim3 tb(32, 8);
  dim3 tg((g.nz + 31) / 32, (g.nx + 31) / 32, 3);
  finalize_kernel<<<tg, tb, 0, s>>>(g, gacc, nslots, misfit_half, result);
}

}  // namespace fwi
