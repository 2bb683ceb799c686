// =============================================================================
// TEST INFRASTRUCTURE ONLY -- CPU oracle for the FWI hot path.
//
// A CPU restatement (C++17 + OpenMP, float32 storage with the reference's
// float64 promotions) of the algorithm behind `cufd()` in the reference,
// deps/CustomOps/FWI/Src/ (paths below are relative to that directory).
// It exists to CHECK the CUDA product path; nothing in fwiflow_jl_b200/ links,
// imports or executes it.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may use it.
//
// Parity pinning: the reference ships no golden vectors for this path
// (SURVEY.md section 8c).  The oracle is pinned against OUTPUTS OF THE
// REFERENCE ITSELF: oracle/_ref/libCUFD_ref.so (the unmodified reference CUDA
// library rebuilt for sm_100 by oracle/build_ref.sh) run on a B200 by
// tests/golden/make_golden.py; the resulting vectors are committed under
// tests/golden/ and checked by tests/test_oracle_golden.py.
//
// Every function cites the reference file:line it follows.  The arithmetic
// keeps the reference's mixed precision: variables are float, but expressions
// that contain the double literals 2.0 / 1.0 / 4.0 / PI / MEGA / RSXXZZ are
// evaluated in double exactly where the reference's C expressions are.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

constexpr double kPi = 3.141592653589793238462643383279502884197169;  // utilities.h:15
constexpr double kMega = 1e6;                                         // utilities.h:16
constexpr double kRsxxzz = 3.0;                                       // utilities.h:21

struct Geom {
  int nz, nx, nPml, nPad, nSteps;
  float dz, dx, dt, f0;
};

// 4th-order staggered operators.  c1 = 9/8, c2 = 1/24 as floats
// (el_stress.cu:45-46).  s = stride between the differenced samples.
const float C1 = 9.0 / 8.0;
const float C2 = 1.0 / 24.0;

// "backward" staggered difference: uses f[i-2s], f[i-s], f[i], f[i+s]
// (el_stress.cu:54-55 dvz_dz/dvx_dx; el_velocity.cu:51,64).
inline float dminus(const float *f, int64_t i, int64_t s, float h) {
  return (C1 * (f[i] - f[i - s]) - C2 * (f[i + s] - f[i - 2 * s])) / h;
}
// "forward" staggered difference: uses f[i-s], f[i], f[i+s], f[i+2s]
// (el_stress.cu:71-72 dvx_dz/dvz_dx; el_velocity.cu:50,65).
inline float dplus(const float *f, int64_t i, int64_t s, float h) {
  return (C1 * (f[i + s] - f[i]) - C2 * (f[i + 2 * s] - f[i - s])) / h;
}
// adjoint-kernel spellings: (-c1*(..) + c2*(..))/h (el_stress_adj.cu:54-61).
inline float adminus(const float *f, int64_t i, int64_t s, float h) {
  return (-C1 * (f[i] - f[i - s]) + C2 * (f[i + s] - f[i - 2 * s])) / h;
}
inline float adplus(const float *f, int64_t i, int64_t s, float h) {
  return (-C1 * (f[i + s] - f[i]) + C2 * (f[i + 2 * s] - f[i - s])) / h;
}

// ---- CPML profiles: utilities.cu:242-358 (cpmlInit) ------------------------
struct Profile {
  std::vector<float> K, a, b, Kh, ah, bh;
};

Profile cpml_profile(int N, int nPml, float dh, float f0, float dt) {
  Profile p;
  p.K.assign(N, 1.0f); p.Kh.assign(N, 1.0f);
  p.a.assign(N, 0.0f); p.ah.assign(N, 0.0f);
  p.b.assign(N, 0.0f); p.bh.assign(N, 0.0f);
  std::vector<float> damp(N, 0.0f), damph(N, 0.0f), alpha(N, 0.0f), alphah(N, 0.0f);
  const float Rcoef = 0.0008;
  const float K_MAX = 2.0;
  const float ALPHA_MAX = 2.0 * kPi * (f0 / 2.0);
  const float NPOWER = 8.0;
  const float c1 = 0.25, c2 = 0.75, c3 = 0.0;
  float thick = nPml * dh;
  float CpAve = 3000.0;  // utilities.cu:259 -- model independent
  float d0 = -(NPOWER + 1) * CpAve * std::log(Rcoef) / (2.0 * thick);
  auto dampf = [&](float dn) {
    return d0 * (c1 * dn + c2 * std::pow(dn, NPOWER) + c3 * std::pow(dn, 2 * NPOWER));
  };
  for (int i = 0; i < N; i++) {
    float depth, dn;
    depth = (nPml - i) * dh;                       // left edge, integer points
    if (depth >= 0.0) {
      dn = depth / thick;
      damp[i] = dampf(dn);
      p.K[i] = 1.0 + (K_MAX - 1.0) * std::pow(dn, NPOWER);
      alpha[i] = ALPHA_MAX * (1.0 - dn);
    }
    depth = (nPml - i - 0.5) * dh;                 // left edge, half points
    if (depth >= 0.0) {
      dn = depth / thick;
      damph[i] = dampf(dn);
      p.Kh[i] = 1.0 + (K_MAX - 1.0) * std::pow(dn, NPOWER);
      alphah[i] = ALPHA_MAX * (1.0 - dn);
    }
    depth = (nPml - N + i) * dh;                   // right edge, integer points
    if (depth >= 0.0) {
      dn = depth / thick;
      damp[i] = dampf(dn);
      p.K[i] = 1.0 + (K_MAX - 1.0) * std::pow(dn, NPOWER);
      alpha[i] = ALPHA_MAX * (1.0 - dn);
    }
    depth = (nPml - N + i + 0.5) * dh;             // right edge, half points
    if (depth >= 0.0) {
      dn = depth / thick;
      damph[i] = dampf(dn);
      p.Kh[i] = 1.0 + (K_MAX - 1.0) * powf(dn, NPOWER);
      alphah[i] = ALPHA_MAX * (1.0 - dn);
    }
    if (alpha[i] < 0.0) alpha[i] = 0.0;
    if (alphah[i] < 0.0) alphah[i] = 0.0;
    p.b[i] = expf(-(damp[i] / p.K[i] + alpha[i]) * dt);
    p.bh[i] = expf(-(damph[i] / p.Kh[i] + alphah[i]) * dt);
    if (std::fabs(damp[i]) > 1.0e-6)
      p.a[i] = damp[i] * (p.b[i] - 1.0) / (p.K[i] * (damp[i] + p.K[i] * alpha[i]));
    if (std::fabs(damph[i]) > 1.0e-6)
      p.ah[i] = damph[i] * (p.bh[i] - 1.0) / (p.Kh[i] * (damph[i] + p.Kh[i] * alphah[i]));
  }
  return p;
}

// ---- time taper without windows: utilities.cu:707-747 ----------------------
// returns the multiplier applied to sample idt (window_amp*window_amp, float)
inline bool taper_weight(int idt, int nt, float dt, float ratio, float *w2) {
  float window_amp = 1.0;
  float t = idt * dt;
  float t0 = 0;
  float t3 = nt * dt;
  float offset = nt * dt * ratio;
  if (2.0 * offset >= t3 - t0) return false;  // "Window error 2": sample untouched
  float t1 = t0 + offset;
  float t2 = t3 - offset;
  if (t >= t0 && t < t1) window_amp = std::sin(kPi / 2.0 * (t - t0) / (t1 - t0));
  else if (t >= t1 && t < t2) window_amp = 1.0;
  else if (t >= t2 && t < t3) window_amp = std::cos(kPi / 2.0 * (t - t2) / (t3 - t2));
  else window_amp = 0.0;
  *w2 = window_amp * window_amp;
  return true;
}
void taper(float *data, int nrec, int nt, float dt, float ratio) {
  for (int t = 0; t < nt; t++) {
    float w2;
    if (!taper_weight(t, nt, dt, ratio, &w2)) continue;
    for (int r = 0; r < nrec; r++) data[(int64_t)r * nt + t] *= w2;
  }
}

// ---- per-trace windows + weights (para "if_win") ----------------------------
// utilities.cu:654-706 (cuda_window with d_win_start / d_win_end / d_weights), applied at libCUFD.cu:257-266 to the
// observed and synthetic traces with ratio win_ratio and at libCUFD.cu:304-309 to the residual with ratio 0.1.
// An empty window ("Window error 1") leaves the trace untouched, weight included.
void trace_windows(float *data, int nrec, int nt, float dt, const float *win_start, const float *win_end,
                   const float *weights, float ratio) {
  const double PI = 3.141592653589793238462643383279502884197169;
  for (int r = 0; r < nrec; r++) {
    float t0 = win_start[r], t3 = win_end[r];
    const float t_max = nt * dt;
    if (t0 < 0.0f) t0 = 0.0f;
    if (t0 > t_max) t0 = t_max;
    if (t3 < 0.0f) t3 = 0.0f;
    if (t3 > t_max) t3 = t_max;
    const float offset = (t3 - t0) * ratio;
    if (offset <= 0.0f) continue;
    const float t1 = t0 + offset, t2 = t3 - offset;
    for (int it = 0; it < nt; it++) {
      const float t = it * dt;
      float amp;
      if (t >= t0 && t < t1)
        amp = (float)std::sin(PI / 2.0 * (double)(t - t0) / (double)(t1 - t0));
      else if (t >= t1 && t < t2)
        amp = 1.0f;
      else if (t >= t2 && t < t3)
        amp = (float)std::cos(PI / 2.0 * (double)(t - t2) / (double)(t3 - t2));
      else
        amp = 0.0f;
      data[(int64_t)r * nt + it] *= amp * amp * weights[r];
    }
  }
}

// ---- sum of squares with the reference's reduction order --------------------
// utilities.cu:169-205: 512 lanes each accumulate a strided subsequence with
// powf(a,2), then a halving tree.
float sum_sq_512(const float *err, int64_t ng) {
  const int B = 512;
  float s[B];
  for (int tid = 0; tid < B; tid++) {
    float acc = 0.0f;
    for (int64_t k = 0; k < (ng + B - 1) / B; k++) {
      int64_t id = k * B + tid;
      float a = (id < ng) ? err[id] : 0.0f;
      acc += a * a;
    }
    s[tid] = acc;
  }
  for (int h = B / 2; h >= 1; h /= 2)
    for (int tid = 0; tid < h; tid++) s[tid] += s[tid + h];
  return s[0];
}

// ---- state ------------------------------------------------------------------
struct State {
  Geom g;
  int64_t n;  // nz*nx
  // model (Model.cu:54-87)
  std::vector<float> lam, mu, den, amu, bya, byb;
  Profile pz, px;  // Cpml.cu:46-52 -- z profile has length nz-nPad
  // fields
  std::vector<float> vz, vx, szz, sxx, sxz;
  std::vector<float> vz_a, vx_a, szz_a, sxx_a, sxz_a;
  // CPML memory (libCUFD.cu:105-106)
  std::vector<float> m_vz_z, m_vz_x, m_vx_z, m_vx_x;      // from velocity derivatives
  std::vector<float> m_szz_z, m_sxx_x, m_sxz_z, m_sxz_x;  // from stress derivatives
  std::vector<float> gLam, gMu, gDen;
};

inline int64_t at(const Geom &g, int z, int x) { return (int64_t)x * g.nz + z; }

// Model.cu:15-93 + utilities.cu:125-152 (aveMuInit / aveBycInit)
void init_model(State &S, const double *Lambda, const double *Mu, const double *Den) {
  const Geom &g = S.g;
  S.n = (int64_t)g.nz * g.nx;
  S.lam.resize(S.n); S.mu.resize(S.n); S.den.resize(S.n);
  // libCUFD.cu:68-78: row-major [z][x] double (MPa) -> z-fastest float (Pa)
  for (int i = 0; i < g.nz; i++)
    for (int j = 0; j < g.nx; j++) {
      S.lam[at(g, i, j)] = Lambda[(int64_t)i * g.nx + j] * kMega;
      S.mu[at(g, i, j)] = Mu[(int64_t)i * g.nx + j] * kMega;
      S.den[at(g, i, j)] = Den[(int64_t)i * g.nx + j];
    }
  S.amu.assign(S.n, 0.0f);                 // Model.cu:67
  S.bya.assign(S.n, 1.0 / 1000.0);         // Model.cu:72
  S.byb.assign(S.n, 1.0 / 1000.0);         // Model.cu:73
  for (int x = 2; x <= g.nx - 3; x++)
    for (int z = 2; z <= g.nz - 3; z++) {
      float a = S.mu[at(g, z, x)], b = S.mu[at(g, z + 1, x)];
      float c = S.mu[at(g, z, x + 1)], d = S.mu[at(g, z + 1, x + 1)];
      if (a == 0.0 || b == 0.0 || c == 0.0 || d == 0.0) S.amu[at(g, z, x)] = 0.0;
      else S.amu[at(g, z, x)] = 4.0 / (1.0 / a + 1.0 / b + 1.0 / c + 1.0 / d);
      S.bya[at(g, z, x)] = 2.0 / (S.den[at(g, z + 1, x)] + S.den[at(g, z, x)]);
      S.byb[at(g, z, x)] = 2.0 / (S.den[at(g, z, x + 1)] + S.den[at(g, z, x)]);
    }
  S.pz = cpml_profile(g.nz - g.nPad, g.nPml, g.dz, g.f0, g.dt);
  S.px = cpml_profile(g.nx, g.nPml, g.dx, g.f0, g.dt);
  S.gLam.assign(S.n, 0.0f); S.gMu.assign(S.n, 0.0f); S.gDen.assign(S.n, 0.0f);
}

// Courant limit: utilities.cu:225-240 with Cp from velInit (utilities.cu:109-123)
bool courant_ok(const State &S, float *cn_out) {
  const Geom &g = S.g;
  float mx = 0.0f;
  bool first = true;
  for (int64_t i = 0; i < S.n; i++) {
    float cp = std::sqrt((S.lam[i] + 2.0 * S.mu[i]) / S.den[i]);
    if (first || cp > mx) { mx = cp; first = false; }
  }
  float dh_min = (g.dz < g.dx) ? g.dz : g.dx;
  float cn = mx * g.dt * sqrtf(2.0) * (1.0 / 24.0 + 9.0 / 8.0) / dh_min;
  if (cn_out) *cn_out = cn;
  return !(cn > 1.0);
}

inline bool in_zpml(const Geom &g, int z) { return z < g.nPml || z > g.nz - g.nPml - g.nPad - 1; }
inline bool in_xpml_s(const Geom &g, int x) { return x < g.nPml || x > g.nx - g.nPml - 1; }  // el_stress.cu:61
inline bool in_xpml_v(const Geom &g, int x) { return x < g.nPml || x > g.nx - g.nPml; }      // el_velocity.cu:56

// el_stress.cu:50-88 (isFor == true)
void stress_fwd(State &S) {
  const Geom &g = S.g;
  const int64_t nz = g.nz;
  const float dt = g.dt;
#pragma omp parallel for schedule(static)
  for (int x = 2; x <= g.nx - 3; x++) {
    const bool xp = in_xpml_s(g, x);
    for (int z = 2; z <= g.nz - g.nPad - 3; z++) {
      const int64_t i = at(g, z, x);
      const bool zp = in_zpml(g, z);
      float dvz_dz = dminus(S.vz.data(), i, 1, g.dz);
      float dvx_dx = dminus(S.vx.data(), i, nz, g.dx);
      if (zp) {
        S.m_vz_z[i] = S.pz.b[z] * S.m_vz_z[i] + S.pz.a[z] * dvz_dz;
        dvz_dz = dvz_dz / S.pz.K[z] + S.m_vz_z[i];
      }
      if (xp) {
        S.m_vx_x[i] = S.px.b[x] * S.m_vx_x[i] + S.px.a[x] * dvx_dx;
        dvx_dx = dvx_dx / S.px.K[x] + S.m_vx_x[i];
      }
      const float lam = S.lam[i], mu = S.mu[i];
      S.szz[i] += ((lam + 2.0 * mu) * dvz_dz + lam * dvx_dx) * dt;
      S.sxx[i] += (lam * dvz_dz + (lam + 2.0 * mu) * dvx_dx) * dt;
      float dvx_dz = dplus(S.vx.data(), i, 1, g.dz);
      float dvz_dx = dplus(S.vz.data(), i, nz, g.dx);
      if (zp) {
        S.m_vx_z[i] = S.pz.bh[z] * S.m_vx_z[i] + S.pz.ah[z] * dvx_dz;
        dvx_dz = dvx_dz / S.pz.Kh[z] + S.m_vx_z[i];
      }
      if (xp) {
        S.m_vz_x[i] = S.px.bh[x] * S.m_vz_x[i] + S.px.ah[x] * dvz_dx;
        dvz_dx = dvz_dx / S.px.Kh[x] + S.m_vz_x[i];
      }
      S.sxz[i] = S.sxz[i] + S.amu[i] * (dvx_dz + dvz_dx) * dt;
    }
  }
}

// el_velocity.cu:45-82 (isFor == true)
void velocity_fwd(State &S) {
  const Geom &g = S.g;
  const int64_t nz = g.nz;
  const float dt = g.dt;
#pragma omp parallel for schedule(static)
  for (int x = 2; x <= g.nx - 3; x++) {
    const bool xp = in_xpml_v(g, x);
    for (int z = 2; z <= g.nz - g.nPad - 3; z++) {
      const int64_t i = at(g, z, x);
      const bool zp = in_zpml(g, z);
      float dszz_dz = dplus(S.szz.data(), i, 1, g.dz);
      float dsxz_dx = dminus(S.sxz.data(), i, nz, g.dx);
      if (zp) {
        S.m_szz_z[i] = S.pz.bh[z] * S.m_szz_z[i] + S.pz.ah[z] * dszz_dz;
        dszz_dz = dszz_dz / S.pz.Kh[z] + S.m_szz_z[i];
      }
      if (xp) {
        S.m_sxz_x[i] = S.px.b[x] * S.m_sxz_x[i] + S.px.a[x] * dsxz_dx;
        dsxz_dx = dsxz_dx / S.px.K[x] + S.m_sxz_x[i];
      }
      S.vz[i] += (dszz_dz + dsxz_dx) * S.bya[i] * dt;
      float dsxz_dz = dminus(S.sxz.data(), i, 1, g.dz);
      float dsxx_dx = dplus(S.sxx.data(), i, nz, g.dx);
      if (zp) {
        S.m_sxz_z[i] = S.pz.b[z] * S.m_sxz_z[i] + S.pz.a[z] * dsxz_dz;
        dsxz_dz = dsxz_dz / S.pz.K[z] + S.m_sxz_z[i];
      }
      if (xp) {
        S.m_sxx_x[i] = S.px.bh[x] * S.m_sxx_x[i] + S.px.ah[x] * dsxx_dx;
        dsxx_dx = dsxx_dx / S.px.Kh[x] + S.m_sxx_x[i];
      }
      S.vx[i] += (dsxz_dz + dsxx_dx) * S.byb[i] * dt;
    }
  }
}

inline bool in_box(const Geom &g, int z, int x) {
  return z >= g.nPml && z <= g.nz - g.nPad - 1 - g.nPml && x >= g.nPml && x <= g.nx - 1 - g.nPml;
}

// el_velocity.cu:84-117 (isFor == false): reverse update on the inner box +
// density imaging condition.  The reference sprays with atomicAdd; here the
// spray is applied sequentially (column by column) -- same set of addends.
void velocity_bwd(State &S) {
  const Geom &g = S.g;
  const int64_t nz = g.nz;
  const float dt = g.dt;
  const int zlo = g.nPml, zhi = g.nz - g.nPad - 1 - g.nPml;
  const int xlo = g.nPml, xhi = g.nx - 1 - g.nPml;
  std::vector<float> ga(S.n, 0.0f), gb(S.n, 0.0f);
#pragma omp parallel for schedule(static)
  for (int x = xlo; x <= xhi; x++)
    for (int z = zlo; z <= zhi; z++) {
      const int64_t i = at(g, z, x);
      float dszz_dz = dplus(S.szz.data(), i, 1, g.dz);
      float dsxz_dx = dminus(S.sxz.data(), i, nz, g.dx);
      S.vz[i] -= (dszz_dz + dsxz_dx) * S.bya[i] * dt;
      float dsxz_dz = dminus(S.sxz.data(), i, 1, g.dz);
      float dsxx_dx = dplus(S.sxx.data(), i, nz, g.dx);
      S.vx[i] -= (dsxz_dz + dsxx_dx) * S.byb[i] * dt;
      ga[i] = -S.vz_a[i] * (dszz_dz + dsxz_dx) * dt * (-std::pow(S.bya[i], 2) / 2.0);
      gb[i] = -S.vx_a[i] * (dsxz_dz + dsxx_dx) * dt * (-std::pow(S.byb[i], 2) / 2.0);
    }
  for (int x = xlo; x <= xhi; x++)
    for (int z = zlo; z <= zhi; z++) {
      const int64_t i = at(g, z, x);
      S.gDen[i] += ga[i];
      S.gDen[i] += gb[i];
      if (z + 1 <= zhi) S.gDen[i + 1] += ga[i];
      // el_velocity.cu:109: `gidx+1<=gidx<=nx-1-nPml` is always true
      S.gDen[i + nz] += gb[i];
    }
}

// el_stress.cu:90-131 (isFor == false): reverse update + lambda / mu imaging.
void stress_bwd(State &S) {
  const Geom &g = S.g;
  const int64_t nz = g.nz;
  const float dt = g.dt;
  const int zlo = g.nPml, zhi = g.nz - g.nPad - 1 - g.nPml;
  const int xlo = g.nPml, xhi = g.nx - 1 - g.nPml;
  std::vector<float> sc(S.n, 0.0f);
  std::vector<uint8_t> has(S.n, 0);
#pragma omp parallel for schedule(static)
  for (int x = xlo; x <= xhi; x++)
    for (int z = zlo; z <= zhi; z++) {
      const int64_t i = at(g, z, x);
      float dvz_dz = dminus(S.vz.data(), i, 1, g.dz);
      float dvx_dx = dminus(S.vx.data(), i, nz, g.dx);
      const float lam = S.lam[i], mu = S.mu[i];
      S.szz[i] -= ((lam + 2.0 * mu) * dvz_dz + lam * dvx_dx) * dt;
      S.sxx[i] -= (lam * dvz_dz + (lam + 2.0 * mu) * dvx_dx) * dt;
      float dvx_dz = dplus(S.vx.data(), i, 1, g.dz);
      float dvz_dx = dplus(S.vz.data(), i, nz, g.dx);
      S.sxz[i] -= S.amu[i] * (dvx_dz + dvz_dx) * dt;
      S.gLam[i] += -(S.szz_a[i] + S.sxx_a[i]) * (dvz_dz + dvx_dx) * dt * kMega;
      S.gMu[i] += (-2.0 * S.szz_a[i] * dvz_dz * dt - 2.0 * S.sxx_a[i] * dvx_dx * dt) * kMega;
      if (S.amu[i] != 0.0) {
        float scale = -S.sxz_a[i] * (dvx_dz + dvz_dx) * dt * S.amu[i] /
                      (1.0 / S.mu[i] + 1.0 / S.mu[i + 1] + 1.0 / S.mu[i + nz] +
                       1.0 / S.mu[i + nz + 1]) * kMega;
        sc[i] = scale;
        has[i] = 1;
      }
    }
  for (int x = xlo; x <= xhi; x++)
    for (int z = zlo; z <= zhi; z++) {
      const int64_t i = at(g, z, x);
      if (!has[i]) continue;
      const float scale = sc[i];
      S.gMu[i] += 1.0 / std::pow((double)S.mu[i], 2) * scale;
      if (z + 1 <= zhi) S.gMu[i + 1] += 1.0 / std::pow((double)S.mu[i + 1], 2) * scale;
      // el_stress.cu:120: always-true guard
      S.gMu[i + nz] += 1.0 / std::pow((double)S.mu[i + nz], 2) * scale;
      if (z + 1 <= zhi && x + 1 <= xhi)
        S.gMu[i + nz + 1] += 1.0 / std::pow((double)S.mu[i + nz + 1], 2) * scale;
    }
}

// el_velocity_adj.cu:22-108
void velocity_adj(State &S) {
  const Geom &g = S.g;
  const int64_t nz = g.nz;
  const float dt = g.dt;
#pragma omp parallel for schedule(static)
  for (int x = 2; x <= g.nx - 3; x++) {
    const bool xp = in_xpml_s(g, x);
    for (int z = 2; z <= g.nz - g.nPad - 3; z++) {
      const int64_t i = at(g, z, x);
      const bool zp = in_zpml(g, z);
      const float lambda = S.lam[i], mu = S.mu[i];
      float dpsixx_dx = adplus(S.m_vx_x.data(), i, nz, g.dx);
      float dszz_dx = adplus(S.szz_a.data(), i, nz, g.dx);
      float dsxx_dx = adplus(S.sxx_a.data(), i, nz, g.dx);
      float dpsixz_dz = adminus(S.m_vx_z.data(), i, 1, g.dz);
      float dsxz_dz = adminus(S.sxz_a.data(), i, 1, g.dz);
      S.vx_a[i] += (S.px.a[x] * dpsixx_dx + lambda * dszz_dx / S.px.K[x] * dt +
                    (lambda + 2.0 * mu) * dsxx_dx / S.px.K[x] * dt +
                    S.pz.ah[z] * dpsixz_dz + S.amu[i] / S.pz.Kh[z] * dsxz_dz * dt);
      if (xp) S.m_sxx_x[i] = S.px.bh[x] * S.m_sxx_x[i] + S.byb[i] * S.vx_a[i] * dt;
      if (zp) S.m_sxz_z[i] = S.pz.b[z] * S.m_sxz_z[i] + S.byb[i] * S.vx_a[i] * dt;

      float dpsizz_dz = adplus(S.m_vz_z.data(), i, 1, g.dz);
      float dszz_dz = adplus(S.szz_a.data(), i, 1, g.dz);
      float dsxx_dz = adplus(S.sxx_a.data(), i, 1, g.dz);
      float dpsizx_dx = adminus(S.m_vz_x.data(), i, nz, g.dx);
      float dsxz_dx = adminus(S.sxz_a.data(), i, nz, g.dx);
      S.vz_a[i] += (S.pz.a[z] * dpsizz_dz + (lambda + 2.0 * mu) * dszz_dz / S.pz.K[z] * dt +
                    lambda * dsxx_dz / S.pz.K[z] * dt + S.px.ah[x] * dpsizx_dx +
                    S.amu[i] / S.px.Kh[x] * dsxz_dx * dt);
      if (xp) S.m_sxz_x[i] = S.px.b[x] * S.m_sxz_x[i] + S.bya[i] * S.vz_a[i] * dt;
      if (zp) S.m_szz_z[i] = S.pz.bh[z] * S.m_szz_z[i] + S.bya[i] * S.vz_a[i] * dt;
    }
  }
}

// el_stress_adj.cu:22-104 (psi arrays updated over the whole active region)
void stress_adj(State &S) {
  const Geom &g = S.g;
  const int64_t nz = g.nz;
  const float dt = g.dt;
#pragma omp parallel for schedule(static)
  for (int x = 2; x <= g.nx - 3; x++)
    for (int z = 2; z <= g.nz - g.nPad - 3; z++) {
      const int64_t i = at(g, z, x);
      const float lambda = S.lam[i], mu = S.mu[i];
      float dphi_xz_x_dx = adplus(S.m_sxz_x.data(), i, nz, g.dx);
      float dvz_dx = adplus(S.vz_a.data(), i, nz, g.dx);
      float dphi_xz_z_dz = adplus(S.m_sxz_z.data(), i, 1, g.dz);
      float dvx_dz = adplus(S.vx_a.data(), i, 1, g.dz);
      S.sxz_a[i] += S.px.a[x] * dphi_xz_x_dx + dvz_dx / S.px.K[x] * S.bya[i] * dt +
                    S.pz.a[z] * dphi_xz_z_dz + dvx_dz / S.pz.K[z] * S.byb[i] * dt;
      S.m_vz_x[i] = S.px.bh[x] * S.m_vz_x[i] + S.sxz_a[i] * S.amu[i] * dt;
      S.m_vx_z[i] = S.pz.bh[z] * S.m_vx_z[i] + S.sxz_a[i] * S.amu[i] * dt;

      float dphi_xx_x_dx = adminus(S.m_sxx_x.data(), i, nz, g.dx);
      float dvx_dx = adminus(S.vx_a.data(), i, nz, g.dx);
      float dphi_zz_z_dz = adminus(S.m_szz_z.data(), i, 1, g.dz);
      float dvz_dz = adminus(S.vz_a.data(), i, 1, g.dz);
      S.sxx_a[i] += S.px.ah[x] * dphi_xx_x_dx + S.byb[i] * dvx_dx / S.px.Kh[x] * dt;
      S.szz_a[i] += S.pz.ah[z] * dphi_zz_z_dz + S.bya[i] * dvz_dz / S.pz.Kh[z] * dt;
      S.m_vx_x[i] = S.px.b[x] * S.m_vx_x[i] + lambda * S.szz_a[i] * dt +
                    (lambda + 2.0 * mu) * S.sxx_a[i] * dt;
      S.m_vz_z[i] = S.pz.b[z] * S.m_vz_z[i] + (lambda + 2.0 * mu) * S.szz_a[i] * dt +
                    lambda * S.sxx_a[i] * dt;
    }
}

// ---- boundary frames: Boundary.cu:17-27, utilities.cu:361-424 ---------------
struct Frames {
  int nzB, nxB, L, len;
  std::vector<float> store[5];  // szz, sxz, sxx, vz, vx -- [it][len]
};

inline int64_t frame_cell(const Geom &g, const Frames &F, int k) {
  const int L = F.L, nzB = F.nzB, nxB = F.nxB;
  int iRow, jCol, z, x;
  if (k < L * nzB) {                       // left columns
    jCol = k / nzB; iRow = k - jCol * nzB;
    z = iRow + g.nPml - 2; x = jCol + g.nPml - 2;
  } else if (k < 2 * L * nzB) {            // right columns
    int q = k - L * nzB;
    jCol = q / nzB; iRow = q - jCol * nzB;
    z = iRow + g.nPml - 2; x = g.nx - g.nPml - jCol - 1 + 2;
  } else if (k < L * (2 * nzB + nxB)) {    // top rows
    int q = k - 2 * L * nzB;
    iRow = q / nxB; jCol = q - iRow * nxB;
    z = iRow + g.nPml - 2; x = jCol + g.nPml - 2;
  } else {                                 // bottom rows
    int q = k - L * (2 * nzB + nxB);
    iRow = q / nxB; jCol = q - iRow * nxB;
    z = g.nz - g.nPml - g.nPad - iRow - 1 + 2; x = jCol + g.nPml - 2;
  }
  return at(g, z, x);
}

}  // namespace

// =============================================================================
// C entry point (ctypes).  Mirrors cufd() (libCUFD.cu:34-580) with the file
// I/O lifted out: geometry comes as numbers, receivers as flat index arrays
// ALREADY offset by nPml (Src_Rec.cu:86-113 is done by the caller), observed
// data comes as [shot][rec][time] float32 arrays.
//   calc_id 0: misfit only   1: gradients   2: synthetic traces only
// syn_out / res_out / obs_cond_out (optional, may be NULL): per shot of the
// group, concatenated, each [nrec][nSteps] time-fastest (libCUFD.cu:514-521).
// snap_it >= 0: copy vx (forward, state at time snap_it before the update) into
// snap_fwd and vz (reconstructed, after backward step snap_it) into snap_back
// (libCUFD.cu:210-213, 429-431) for shot 0 of the group.
// returns 0, or 1 when the Courant limit is violated (utilities.cu:239).
// =============================================================================
static int cufd_impl(
    int nz, int nx, int nPml, int nPad, int nSteps, float dz, float dx, float dt, float f0,
    int calc_id, int group_size, const int *shot_ids,
    const double *Lambda, const double *Mu, const double *Den, const double *stf,
    const int *z_src, const int *x_src,          // [group_size], padded coords
    const int *rec_off,                          // [group_size+1] offsets into z_rec/x_rec
    const int *z_rec, const int *x_rec,          // padded coords
    const float *obs_in,                         // calc 0/1: concatenated [rec][time]
    double *misfit, double *grad_Lambda, double *grad_Mu, double *grad_Den, double *grad_stf,
    float *syn_out, float *res_out, float *obs_cond_out,
    int snap_it, float *snap_fwd, float *snap_back,
    const float *win_start, const float *win_end, const float *weights) {   // per receiver (rec_off), or all NULL
  State S;
  S.g = Geom{nz, nx, nPml, nPad, nSteps, dz, dx, dt, f0};
  const Geom &g = S.g;
  const bool if_res = (calc_id == 0 || calc_id == 1);   // Parameter.cpp:132-144
  const bool withAdj = (calc_id == 1);
  init_model(S, Lambda, Mu, Den);
  if (!courant_ok(S, nullptr)) return 1;
  const float win_ratio = 0.005;  // libCUFD.cu:63
  const int64_t n = S.n;

  Frames F;
  if (withAdj) {
    F.nzB = nz - 2 * nPml - nPad + 4;
    F.nxB = nx - 2 * nPml + 4;
    F.L = 5;
    F.len = 2 * (F.L * F.nzB + F.L * F.nxB);
    for (auto &s : F.store) s.assign((size_t)F.len * nSteps, 0.0f);
  }
  std::vector<int64_t> fcell;
  if (withAdj) {
    fcell.resize(F.len);
    for (int k = 0; k < F.len; k++) fcell[k] = frame_cell(g, F, k);
  }

  float h_l2Obj = 0.0;  // libCUFD.cu:110 -- float accumulator across shots
  int64_t trace_off = 0;
  for (int iShot = 0; iShot < group_size; iShot++) {
    const int nrec = rec_off[iShot + 1] - rec_off[iShot];
    const int *zr = z_rec + rec_off[iShot];
    const int *xr = x_rec + rec_off[iShot];
    const int64_t src = at(g, z_src[iShot], x_src[iShot]);
    // Src_Rec.cu:132-144: stf row = GLOBAL shot id, cast to float, tapered (ratio 0.001)
    std::vector<float> source(nSteps);
    for (int it = 0; it < nSteps; it++) source[it] = stf[(int64_t)shot_ids[iShot] * nSteps + it];
    taper(source.data(), 1, nSteps, dt, 0.001);

    // libCUFD.cu:163-182
    for (auto *f : {&S.vz, &S.vx, &S.szz, &S.sxx, &S.sxz, &S.vz_a, &S.vx_a, &S.szz_a, &S.sxx_a,
                    &S.sxz_a, &S.m_vz_z, &S.m_vz_x, &S.m_vx_z, &S.m_vx_x, &S.m_szz_z, &S.m_sxx_x,
                    &S.m_sxz_z, &S.m_sxz_x})
      f->assign(n, 0.0f);
    std::vector<float> data((size_t)nrec * nSteps, 0.0f), obs, res;
    if (if_res) {
      obs.assign(obs_in + trace_off, obs_in + trace_off + (int64_t)nrec * nSteps);
      res.assign((size_t)nrec * nSteps, 0.0f);
    }

    // ---- forward time loop: libCUFD.cu:202-240 ----
    for (int it = 0; it <= nSteps - 2; it++) {
      if (withAdj) {
        const float *flds[5] = {S.szz.data(), S.sxz.data(), S.sxx.data(), S.vz.data(), S.vx.data()};
        for (int f = 0; f < 5; f++) {
          float *dst = F.store[f].data() + (size_t)it * F.len;
          for (int k = 0; k < F.len; k++) dst[k] = flds[f][fcell[k]];
        }
      }
      if (it == snap_it && iShot == 0 && snap_fwd) std::memcpy(snap_fwd, S.vx.data(), n * sizeof(float));
      stress_fwd(S);
      {  // add_source: utilities.cu:521-555 -- the 9x9 stamp is exp(-1000 r^2): 1 at the centre, 0 elsewhere
        const float amp = source[it];
        const float scale = std::pow(1500.0, 2);
        S.szz[src] += scale * amp * dt * 1.0f;
        S.sxx[src] += kRsxxzz * scale * amp * dt * 1.0f;
      }
      velocity_fwd(S);
      for (int r = 0; r < nrec; r++) {  // recording: utilities.cu:557-567
        const int64_t c = at(g, zr[r], xr[r]);
        data[(int64_t)r * nSteps + it + 1] = S.szz[c] + kRsxxzz * S.sxx[c];
      }
    }

    // ---- residual: libCUFD.cu:254-330 ----
    if (if_res) {
      const bool if_win = win_start != nullptr;
      const float *w0 = if_win ? win_start + rec_off[iShot] : nullptr;
      const float *w1 = if_win ? win_end + rec_off[iShot] : nullptr;
      const float *ww = if_win ? weights + rec_off[iShot] : nullptr;
      if (if_win) {  // libCUFD.cu:257-266
        trace_windows(obs.data(), nrec, nSteps, dt, w0, w1, ww, win_ratio);
        trace_windows(data.data(), nrec, nSteps, dt, w0, w1, ww, win_ratio);
      } else {
        taper(obs.data(), nrec, nSteps, dt, win_ratio);
        taper(data.data(), nrec, nSteps, dt, win_ratio);
      }
      for (int r = 0; r < nrec; r++)  // gpuMinus: utilities.cu:154-167
        for (int t = 0; t < nSteps; t++) {
          const int64_t k = (int64_t)r * nSteps + t;
          res[k] = (t > 0) ? obs[k] - data[k] : 0.0f;
        }
      h_l2Obj += sum_sq_512(res.data(), (int64_t)nrec * nSteps);
      if (if_win) trace_windows(res.data(), nrec, nSteps, dt, w0, w1, ww, 0.1f);   // libCUFD.cu:304-309
      else taper(res.data(), nrec, nSteps, dt, win_ratio);
    }

    // ---- backward: libCUFD.cu:334-457 ----
    if (withAdj) {
      for (auto *f : {&S.vz_a, &S.vx_a, &S.szz_a, &S.sxx_a, &S.sxz_a, &S.m_vz_z, &S.m_vz_x, &S.m_vx_z,
                      &S.m_vx_x, &S.m_szz_z, &S.m_sxz_x, &S.m_sxz_z, &S.m_sxx_x})
        f->assign(n, 0.0f);
      std::vector<float> stfGrad(nSteps, 0.0f);
      auto inject = [&](int it) {  // res_injection: utilities.cu:569-580
        for (int r = 0; r < nrec; r++) {
          const int64_t c = at(g, zr[r], xr[r]);
          const float v = res[(int64_t)r * nSteps + it];
          S.szz_a[c] += v;
          S.sxx_a[c] += kRsxxzz * v;
        }
      };
      velocity_adj(S);
      inject(nSteps - 1);
      stress_adj(S);
      for (int it = nSteps - 2; it >= 0; it--) {
        stfGrad[it] = -(S.szz_a[src] + kRsxxzz * S.sxx_a[src]) * dt;  // source_grad: utilities.cu:582-593
        velocity_bwd(S);
        for (int k = 0; k < F.len; k++) {  // to_bnd(v): Boundary.cu:91-99
          S.vz[fcell[k]] = F.store[3][(size_t)it * F.len + k];
          S.vx[fcell[k]] = F.store[4][(size_t)it * F.len + k];
        }
        {
          const float amp = source[it];
          const float scale = std::pow(1500.0, 2);
          S.szz[src] -= scale * amp * dt * 1.0f;
          S.sxx[src] -= kRsxxzz * scale * amp * dt * 1.0f;
        }
        stress_bwd(S);
        for (int k = 0; k < F.len; k++) {  // to_bnd(sigma): Boundary.cu:80-90
          S.szz[fcell[k]] = F.store[0][(size_t)it * F.len + k];
          S.sxz[fcell[k]] = F.store[1][(size_t)it * F.len + k];
          S.sxx[fcell[k]] = F.store[2][(size_t)it * F.len + k];
        }
        velocity_adj(S);
        inject(it);
        stress_adj(S);
        if (it == snap_it && iShot == 0 && snap_back) std::memcpy(snap_back, S.vz.data(), n * sizeof(float));
      }
      if (grad_stf)  // libCUFD.cu:454-456: row = position in the group
        for (int it = 0; it < nSteps; it++) grad_stf[(int64_t)iShot * nSteps + it] = stfGrad[it];
    }

    if (syn_out) std::memcpy(syn_out + trace_off, data.data(), data.size() * sizeof(float));
    if (if_res && res_out) std::memcpy(res_out + trace_off, res.data(), res.size() * sizeof(float));
    if (if_res && obs_cond_out) std::memcpy(obs_cond_out + trace_off, obs.data(), obs.size() * sizeof(float));
    trace_off += (int64_t)nrec * nSteps;
  }

  if (withAdj) {  // libCUFD.cu:472-486
    for (int i = 0; i < nz; i++)
      for (int j = 0; j < nx; j++) {
        grad_Lambda[(int64_t)i * nx + j] = S.gLam[at(g, i, j)];
        grad_Mu[(int64_t)i * nx + j] = S.gMu[at(g, i, j)];
        grad_Den[(int64_t)i * nx + j] = S.gDen[at(g, i, j)];
      }
  }
  if (if_res && !withAdj) {  // libCUFD.cu:528-536
    h_l2Obj = 0.5 * h_l2Obj;
    if (misfit) *misfit = h_l2Obj;
  }
  return 0;
}

extern "C" int fwi_oracle_cufd(
    int nz, int nx, int nPml, int nPad, int nSteps, float dz, float dx, float dt, float f0,
    int calc_id, int group_size, const int *shot_ids,
    const double *Lambda, const double *Mu, const double *Den, const double *stf,
    const int *z_src, const int *x_src, const int *rec_off, const int *z_rec, const int *x_rec,
    const float *obs_in,
    double *misfit, double *grad_Lambda, double *grad_Mu, double *grad_Den, double *grad_stf,
    float *syn_out, float *res_out, float *obs_cond_out,
    int snap_it, float *snap_fwd, float *snap_back) {
  return cufd_impl(nz, nx, nPml, nPad, nSteps, dz, dx, dt, f0, calc_id, group_size, shot_ids, Lambda, Mu, Den, stf,
                   z_src, x_src, rec_off, z_rec, x_rec, obs_in, misfit, grad_Lambda, grad_Mu, grad_Den, grad_stf,
                   syn_out, res_out, obs_cond_out, snap_it, snap_fwd, snap_back, nullptr, nullptr, nullptr);
}

// same, with the per-trace windows / weights of para "if_win" (concatenated per receiver like z_rec)
extern "C" int fwi_oracle_cufd_win(
    int nz, int nx, int nPml, int nPad, int nSteps, float dz, float dx, float dt, float f0,
    int calc_id, int group_size, const int *shot_ids,
    const double *Lambda, const double *Mu, const double *Den, const double *stf,
    const int *z_src, const int *x_src, const int *rec_off, const int *z_rec, const int *x_rec,
    const float *obs_in,
    double *misfit, double *grad_Lambda, double *grad_Mu, double *grad_Den, double *grad_stf,
    float *syn_out, float *res_out, float *obs_cond_out,
    int snap_it, float *snap_fwd, float *snap_back,
    const float *win_start, const float *win_end, const float *weights) {
  return cufd_impl(nz, nx, nPml, nPad, nSteps, dz, dx, dt, f0, calc_id, group_size, shot_ids, Lambda, Mu, Den, stf,
                   z_src, x_src, rec_off, z_rec, x_rec, obs_in, misfit, grad_Lambda, grad_Mu, grad_Den, grad_stf,
                   syn_out, res_out, obs_cond_out, snap_it, snap_fwd, snap_back, win_start, win_end, weights);
}

// CPML profiles for inspection by the tests: out = [K,a,b,K_half,a_half,b_half] x N
extern "C" void fwi_oracle_cpml(int N, int nPml, float dh, float f0, float dt, float *out) {
  Profile p = cpml_profile(N, nPml, dh, f0, dt);
  const std::vector<float> *v[6] = {&p.K, &p.a, &p.b, &p.Kh, &p.ah, &p.bh};
  for (int k = 0; k < 6; k++) std::memcpy(out + (size_t)k * N, v[k]->data(), N * sizeof(float));
}

extern "C" float fwi_oracle_courant(int nz, int nx, float dz, float dx, float dt, const double *Lambda,
                                    const double *Mu, const double *Den) {
  State S;
  S.g = Geom{nz, nx, 0, 0, 0, dz, dx, dt, 1.0f};
  S.n = (int64_t)nz * nx;
  S.lam.resize(S.n); S.mu.resize(S.n); S.den.resize(S.n);
  for (int i = 0; i < nz; i++)
    for (int j = 0; j < nx; j++) {
      S.lam[at(S.g, i, j)] = Lambda[(int64_t)i * nx + j] * kMega;
      S.mu[at(S.g, i, j)] = Mu[(int64_t)i * nx + j] * kMega;
      S.den[at(S.g, i, j)] = Den[(int64_t)i * nx + j];
    }
  float cn = 0;
  courant_ok(S, &cn);
  return cn;
}
