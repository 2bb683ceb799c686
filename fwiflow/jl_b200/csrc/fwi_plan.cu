// Plan object + C ABI of libfwi_b200.so (include/fwi_b200.h).
//
// Host orchestration of the FWI hot path: what `cufd()` does in the reference
// (deps/CustomOps/FWI/Src/libCUFD.cu:34-580), re-organised around a reusable
// device-resident plan: shots advance in batches (one launch per time step for the
// whole batch), all buffers live for the life of the plan, results stay on the device
// until asked for.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <list>
#include <memory>
#include <thread>
#include <mutex>
#include <string>
#include <vector>

#include "fwi_host.hpp"
#include "fwi_kernels.cuh"

using namespace fwi;

#define CUDA_OK(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      throw Error(FWI_B200_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                                         std::to_string(__LINE__) + " (" #call ")");                    \
  } while (0)

namespace {

// cudaMalloc failed: the host-buffer entry points answer it by evicting idle cached plans of the device and retrying
struct OomError : Error {
  explicit OomError(const std::string &m) : Error(FWI_B200_ERR_CUDA, m) {}
};

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    const cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), bytes);
    if (e == cudaErrorMemoryAllocation) {
      cudaGetLastError();   // not sticky: clear it
      p = nullptr;
      throw OomError("out of device memory: cudaMalloc of " + std::to_string(bytes >> 20) + " MiB failed");
    }
    CUDA_OK(e);
    n = count;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

}  // namespace

struct fwi_b200_plan {
  int gpu = 0;
  Para para;
  Survey survey;
  Grid g{};
  int group = 0, batch = 0, nrp = 0, max_nrec = 0, ntiles = 0;
  std::vector<int> shot_ids;
  cudaStream_t stream = nullptr;
  std::mutex mu;  // one evaluation at a time per plan
  std::mutex call_mu;  // held by a host-buffer entry point for its whole set_model .. get_result sequence on a cached plan
  int max_batch_req = 0;  // the caller's max_batch (0 = from free memory); kept for re-planning after an eviction
  int layout = 0;         // 0: the caller's (nz, nx) grids are row-major [z][x] (TensorFlow, the reference); 1: column-major (Julia)
  long long launches = 0;
  bool model_set = false, stf_set = false;
  std::vector<char> obs_set;
  int last_calc = -1;
  int cur_f_last = 0;  // forward-field buffer holding the newest state of the last batch

  // device memory
  DevBuf<float> model;        // lam mu den amu bya byb planes
  DevBuf<double> model_in;    // 3 * nz*nx staging of the caller's doubles
  DevBuf<double> vel_in, vel, vel_grad;   // velocity-space front end: staged inputs, masked padded (cp, cs, rho), chain-rule output
  int vel_masked = -1;        // is_masked of the last set_velocities (-1: the model was not set through velocities)
  DevBuf<unsigned int> cpmax;
  DevBuf<float> zprof, xprof, w2;
  DevBuf<float> state, gacc, frames, syn_tr, res_tr;
  DevBuf<float> obs_rt, syn_rt, res_rt, obs_cond_rt;  // [group][max_nrec*nSteps]
  DevBuf<int> src_z, src_x, rec_ptr, rec_loc, rec_id;
  DevBuf<float> win;          // [group][3][max_nrec] win_start | win_end | weights (para if_win)
  DevBuf<float> stf, stf_grad, j_shot, misfit_half, result;
  DevBuf<int> unit_counter;   // reverse step: dynamic unit claim {next unit, CTAs done}
  DevBuf<double> partial;
  int partial_per_shot = 0;
  size_t trace_stride = 0;  // max_nrec * nSteps
  TmaMaps tm{};             // TMA descriptors of the state / model buffers (re-encoded when `state` moves)
  float *tm_state = nullptr, *tm_gacc = nullptr;

  // Overlapped loading of Data/Shot<id>.bin for the host-buffer entry points: a host thread reads the files into two
  // pinned staging buffers and copies them on `copy_stream` while the forward time loop is already being enqueued
  // and executed; the observations are first needed by the residual kernels, which wait on `obs_ready`.
  std::thread loader;
  int loader_code = 0;
  std::string loader_err;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t obs_ready = nullptr, pin_ev[2] = {nullptr, nullptr};
  float *pin[2] = {nullptr, nullptr};
  size_t pin_n = 0;
  size_t tm_state_n = 0;

  float *mplane(int k) { return model.p + (long long)k * g.plane; }

  ~fwi_b200_plan() {
    cudaSetDevice(gpu);
    if (loader.joinable()) loader.join();
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    if (obs_ready) cudaEventDestroy(obs_ready);
    for (int k = 0; k < 2; k++) {
      if (pin_ev[k]) cudaEventDestroy(pin_ev[k]);
      if (pin[k]) cudaFreeHost(pin[k]);
    }
    if (stream) cudaStreamSynchronize(stream);
    model.release(); model_in.release(); vel_in.release(); vel.release(); vel_grad.release(); cpmax.release(); zprof.release(); xprof.release(); w2.release();
    state.release(); gacc.release(); frames.release(); syn_tr.release(); res_tr.release();
    obs_rt.release(); syn_rt.release(); res_rt.release(); obs_cond_rt.release();
    src_z.release(); src_x.release(); rec_ptr.release(); rec_loc.release(); rec_id.release(); win.release();
    stf.release(); stf_grad.release(); j_shot.release(); misfit_half.release(); result.release(); unit_counter.release();
    partial.release();
    if (stream) cudaStreamDestroy(stream);
  }
};

#ifndef FWI_TILE_ZOFF
#define FWI_TILE_ZOFF 1
#endif
#ifndef FWI_SKIP_DEAD_ROWS
#define FWI_SKIP_DEAD_ROWS 1
#endif
#ifndef FWI_ADJ_INDEP
#define FWI_ADJ_INDEP 1
#endif
#ifndef FWI_FRAME_RING
#define FWI_FRAME_RING 2   // saved boundary ring: 2 = the two cells outside the box (default), 5 = the reference's ring (set_option)
#endif
#ifndef FWI_MERGED_BWD
#define FWI_MERGED_BWD 0   // measured: the merged backward launch moves 12 % fewer DRAM bytes but is latency-bound (196 KB of shared
                           // memory leave 60 KB of L1, 59 spilled registers): C2 322 vs 280 ms, C3 15.0 vs 13.3 s per gradient
#endif

namespace {

// depth of the saved boundary ring: 5 = the reference's (2 cells outside + 3 inside the box), 2 = the two outside cells
std::atomic<int> g_frame_ring{FWI_FRAME_RING};
// 1: backward loop = one merged launch per time index (bwd_step_kernel); 0: reverse + adjoint launches (A/B, set_option)
#ifndef FWI_DYN_UNITS
#define FWI_DYN_UNITS 1
#endif
std::atomic<int> g_merged_bwd{FWI_MERGED_BWD};

int round_up(int v, int m) { return (v + m - 1) / m * m; }

void use_device(int gpu) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    throw Error(FWI_B200_ERR_CUDA, std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                                       "); libfwi_b200 has no CPU fallback");
  if (gpu < 0 || gpu >= n) throw Error(FWI_B200_ERR_ARG, "gpu_id " + std::to_string(gpu) + " out of range");
  CUDA_OK(cudaSetDevice(gpu));
}

void build_grid(const Para &p, Grid &g) {
  g.nz = p.nz; g.nx = p.nx; g.nPml = p.nPml; g.nPad = p.nPad; g.nSteps = p.nSteps;
  g.P = round_up(p.nz, 32);
  g.plane = (long long)(p.nx + 2 * XM) * g.P + 2 * SLACK;
  g.plane = (g.plane + 31) / 32 * 32;
  g.origin = SLACK + (long long)XM * g.P;
  g.az_hi = p.nz - p.nPad - 3;
  g.zlive = FWI_SKIP_DEAD_ROWS ? std::min(p.nz, (g.az_hi + 1 + 3) / 4 * 4) : p.nz;
  g.ax_hi = p.nx - 3;
  g.zlo = p.nPml; g.zhi = p.nz - p.nPad - 1 - p.nPml;
  g.xlo = p.nPml; g.xhi = p.nx - 1 - p.nPml;
  {
    // Row offset of the tile grid (a multiple of 4, so that quads stay 16-byte aligned): among the offsets that do not
    // cost an extra row of tiles, take the one with the fewest tile rows whose region (rows z0-4 .. z0+TILE_Z+3)
    // touches an absorbing layer or its 2-cell fringe -- only those rows of tiles execute CPML code.
    const int rows = std::min(p.nz, g.zlive);
    const int t0 = (rows + TILE_Z - 1) / TILE_Z;
    const int zq_lo = p.nPml + 2, zq_hi = p.nz - p.nPad - p.nPml - 3;
    int best_off = 0, best_cnt = 1 << 30;
    for (int off = 0; off > -TILE_Z; off -= 4) {
      const int t = (rows - off + TILE_Z - 1) / TILE_Z;
      if (t > t0) continue;
      int cnt = 0;
      for (int k = 0; k < t; k++) {
        const int z0 = k * TILE_Z + off;
        if (z0 - 4 < zq_lo || z0 + TILE_Z + 3 > zq_hi) cnt++;
      }
      if (cnt < best_cnt) { best_cnt = cnt; best_off = off; }
      if (!FWI_TILE_ZOFF) break;
    }
    g.z_off = best_off;
    g.tiles_z = (rows - g.z_off + TILE_Z - 1) / TILE_Z;
  }
  g.tiles_x = (p.nx + TILE_X - 1) / TILE_X;
  g.dt = p.dt;
  g.rdz = (float)(1.0 / (double)p.dz);
  g.rdx = (float)(1.0 / (double)p.dx);
  g.dz = p.dz;
  g.dx = p.dx;
  // frames: left/right bands = 10 columns x the quads covering rows zlo-2 .. zhi+2; top/bottom bands = 2 quads each
  // for the columns in between (5 consecutive rows always span exactly two aligned quads)
  if (g.zhi - g.zlo + 1 < 8 || g.xhi - g.xlo + 1 < 8)
    throw Error(FWI_B200_ERR_GEOM, "grid too small: fewer than 8 cells between the PML layers");
  // Ring depth: the reverse-time update of a box cell reads two cells beyond the box (4th-order staggered stencils),
  // so those two are all that MUST be replayed from the forward pass; the reference also overwrites the three box cells
  // next to them with their forward values (5-deep ring, Boundary.cu:17-27).  frame_ring = 5 keeps that, 2 stores
  // 2.1x fewer bytes per step (C5: 12.8 instead of 27 GB per shot) and lets the box cells be reconstructed like every
  // other box cell; the gradients agree to rounding (tests/test_parity_gpu.py::test_thin_boundary_frames).
  g.f_in = g_frame_ring.load(std::memory_order_relaxed) >= 5 ? 3 : 0;
  g.f_zq0 = (g.zlo - 2) & ~3;
  g.f_nqB = ((g.zhi + 2) >> 2) - (g.f_zq0 >> 2) + 1;
  g.f_tq0 = (g.zlo - 2) >> 2;
  g.f_ntq = ((g.zlo - 1 + g.f_in) >> 2) - g.f_tq0 + 1;
  g.f_bq0 = (g.zhi + 1 - g.f_in) >> 2;
  g.f_nbq = ((g.zhi + 2) >> 2) - g.f_bq0 + 1;
  g.f_len = 4 * (2 * (2 + g.f_in) * g.f_nqB + (g.f_ntq + g.f_nbq) * std::max(0, g.xhi - g.xlo + 1 - 2 * g.f_in));
}

void upload_profiles(fwi_b200_plan &pl) {
  const Grid &g = pl.g;
  const Para &p = pl.para;
  CpmlProfiles hz = cpml_profiles(p.nz - p.nPad, p.nPml, p.dz, p.f0, p.dt);  // Cpml.cu:46-48
  CpmlProfiles hx = cpml_profiles(p.nx, p.nPml, p.dx, p.f0, p.dt);           // Cpml.cu:50-52
  const int nxp = p.nx + 2 * XM;
  std::vector<float> z((size_t)PR_COUNT * g.P, 0.0f), x((size_t)PR_COUNT * nxp, 0.0f);
  for (int i = 0; i < g.P; i++) { z[PR_RK * g.P + i] = 1.0f; z[PR_RKH * g.P + i] = 1.0f; z[PR_B * g.P + i] = 1.0f; z[PR_BH * g.P + i] = 1.0f; }
  for (int i = 0; i < nxp; i++) { x[PR_RK * nxp + i] = 1.0f; x[PR_RKH * nxp + i] = 1.0f; x[PR_B * nxp + i] = 1.0f; x[PR_BH * nxp + i] = 1.0f; }
  for (int i = 0; i < p.nz - p.nPad; i++) {
    z[PR_RK * g.P + i] = 1.0f / hz.K[i];  z[PR_A * g.P + i] = hz.a[i];  z[PR_B * g.P + i] = hz.b[i];
    z[PR_RKH * g.P + i] = 1.0f / hz.Kh[i]; z[PR_AH * g.P + i] = hz.ah[i]; z[PR_BH * g.P + i] = hz.bh[i];
  }
  for (int i = 0; i < p.nx; i++) {
    x[PR_RK * nxp + i + XM] = 1.0f / hx.K[i];  x[PR_A * nxp + i + XM] = hx.a[i];  x[PR_B * nxp + i + XM] = hx.b[i];
    x[PR_RKH * nxp + i + XM] = 1.0f / hx.Kh[i]; x[PR_AH * nxp + i + XM] = hx.ah[i]; x[PR_BH * nxp + i + XM] = hx.bh[i];
  }
  // one spare row before and after: quads that straddle z < 0 or z >= P read (and discard) them
  pl.zprof.alloc(z.size() + 2 * (size_t)g.P);
  pl.xprof.alloc(x.size());
  CUDA_OK(cudaMemset(pl.zprof.p, 0, pl.zprof.bytes()));
  CUDA_OK(cudaMemcpy(pl.zprof.p + g.P, z.data(), z.size() * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(pl.xprof.p, x.data(), x.size() * sizeof(float), cudaMemcpyHostToDevice));
  std::vector<float> w2;
  if (!taper_weights(p.nSteps, p.dt, 0.005f, w2)) w2.assign(p.nSteps, 1.0f);  // libCUFD.cu:63,268-270
  pl.w2.alloc(w2.size());
  CUDA_OK(cudaMemcpy(pl.w2.p, w2.data(), w2.size() * sizeof(float), cudaMemcpyHostToDevice));
}

// receiver tables: per shot, receivers sorted by owner tile (CSR) so that the tile that owns a
// receiver cell records / injects it from shared memory.
void upload_tables(fwi_b200_plan &pl) {
  const Grid &g = pl.g;
  const int G = pl.group;
  pl.ntiles = g.tiles_z * g.tiles_x;
  pl.max_nrec = 0;
  for (const Shot &s : pl.survey.shots) pl.max_nrec = std::max<int>(pl.max_nrec, (int)s.z_rec.size());
  pl.nrp = std::max(32, round_up(pl.max_nrec, 32));
  std::vector<int> sz(G), sx(G), ptr((size_t)G * (pl.ntiles + 1), 0), loc((size_t)G * pl.nrp, 0), rid((size_t)G * pl.nrp, 0);
  for (int i = 0; i < G; i++) {
    const Shot &s = pl.survey.shots[i];
    // the reference stamps a 9x9 patch around the source (utilities.cu:529-536): keep it inside the grid
    if (s.z_src < 4 || s.z_src > g.nz - 5 || s.x_src < 4 || s.x_src > g.nx - 5)
      throw Error(FWI_B200_ERR_GEOM, "shot" + std::to_string(s.id) + ": source outside the padded grid");
    if (s.z_src > g.az_hi)   // a source in the nPad rows never radiates (its cell is not updated); refuse it instead
      throw Error(FWI_B200_ERR_GEOM, "shot" + std::to_string(s.id) + ": source inside the inactive nPad rows");
    sz[i] = s.z_src;
    sx[i] = s.x_src;
    const int nrec = (int)s.z_rec.size();
    std::vector<int> cnt(pl.ntiles + 1, 0);
    std::vector<int> tile_of(nrec);
    for (int r = 0; r < nrec; r++) {
      const int z = s.z_rec[r], x = s.x_rec[r];
      if (z < 0 || z >= g.nz || x < 0 || x >= g.nx)
        throw Error(FWI_B200_ERR_GEOM, "shot" + std::to_string(s.id) + ": receiver " + std::to_string(r) +
                                           " outside the padded grid");
      if (z > g.az_hi)   // the nPad rows are never updated and are not stored by this implementation (Grid::zlive)
        throw Error(FWI_B200_ERR_GEOM, "shot" + std::to_string(s.id) + ": receiver " + std::to_string(r) +
                                           " inside the inactive nPad rows");
      tile_of[r] = (x / TILE_X) * g.tiles_z + ((z - g.z_off) / TILE_Z);
      cnt[tile_of[r] + 1]++;
    }
    for (int t = 0; t < pl.ntiles; t++) cnt[t + 1] += cnt[t];
    int *p = ptr.data() + (size_t)i * (pl.ntiles + 1);
    std::copy(cnt.begin(), cnt.end(), p);
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (int r = 0; r < nrec; r++) {  // stable: receivers of a tile keep their file order
      const int k = fill[tile_of[r]]++;
      loc[(size_t)i * pl.nrp + k] = ((s.z_rec[r] - g.z_off) % TILE_Z) | ((s.x_rec[r] % TILE_X) << 16);
      rid[(size_t)i * pl.nrp + k] = r;
    }
  }
  if (pl.para.if_win) {
    std::vector<float> w((size_t)G * 3 * std::max(pl.max_nrec, 1), 0.0f);
    for (int i = 0; i < G; i++) {
      const Shot &s = pl.survey.shots[i];
      float *b = w.data() + (size_t)i * 3 * pl.max_nrec;
      std::copy(s.win_start.begin(), s.win_start.end(), b);
      std::copy(s.win_end.begin(), s.win_end.end(), b + pl.max_nrec);
      std::copy(s.weights.begin(), s.weights.end(), b + 2 * pl.max_nrec);
    }
    pl.win.alloc(w.size());
    CUDA_OK(cudaMemcpy(pl.win.p, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  pl.src_z.alloc(G); pl.src_x.alloc(G); pl.rec_ptr.alloc(ptr.size()); pl.rec_loc.alloc(loc.size()); pl.rec_id.alloc(rid.size());
  CUDA_OK(cudaMemcpy(pl.src_z.p, sz.data(), G * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(pl.src_x.p, sx.data(), G * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(pl.rec_ptr.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(pl.rec_loc.p, loc.data(), loc.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(pl.rec_id.p, rid.data(), rid.size() * sizeof(int), cudaMemcpyHostToDevice));
}

size_t per_shot_bytes(const fwi_b200_plan &pl, bool with_frames) {
  const Grid &g = pl.g;
  size_t b = (size_t)(S_COUNT + G_COUNT) * g.plane * sizeof(float);
  b += (size_t)2 * g.nSteps * pl.nrp * sizeof(float);
  if (with_frames) b += (size_t)5 * g.f_len * g.nSteps * sizeof(float);
  return b;
}

void choose_batch(fwi_b200_plan &pl, int max_batch) {
  size_t free_b = 0, total_b = 0;
  CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
  // traces of the whole group (obs + syn) stay resident as well
  const size_t resident = (size_t)2 * pl.group * pl.trace_stride * sizeof(float) + (size_t)8 * pl.g.plane * sizeof(float);
  const size_t budget = free_b > resident ? (size_t)((free_b - resident) * 0.85) : 0;
  size_t fit = budget / per_shot_bytes(pl, true);
  if (fit < 1)
    throw Error(FWI_B200_ERR_CUDA,
                "not enough device memory: the observed + synthetic traces of the whole group (" +
                    std::to_string(resident >> 20) + " MiB for " + std::to_string(pl.group) +
                    " shots) stay resident and one shot needs " + std::to_string(per_shot_bytes(pl, true) >> 20) +
                    " MiB more (wavefields + boundary frames) out of " + std::to_string(free_b >> 20) +
                    " MiB free; call with fewer shots per group");
  int b = (int)std::min<size_t>(fit, (size_t)pl.group);
  // Enough tiles to fill the machine, no more -- and few enough that the x-neighbour of a tile is still in L2 when its
  // halo is read: items run tile-major with the shots of a tile innermost, so the tile one column over comes
  // tiles_z * batch items later, ~0.1 MB of traffic each against 126 MB of L2.  Measured on the C3 grid (tiles_z = 20):
  // 131 us per shot and time index at batch 25, 176 us at batch 64.
  const int l2_cap = std::max(4, 640 / std::max(1, pl.g.tiles_z));
  b = std::min(b, std::min(64, l2_cap));
  if (max_batch > 0) b = std::min(b, max_batch);
  pl.batch = std::max(1, b);
}

void alloc_run_buffers(fwi_b200_plan &pl, int calc_id) {
  const Grid &g = pl.g;
  pl.state.alloc((size_t)pl.batch * S_COUNT * g.plane);
  if (calc_id == 1) pl.gacc.alloc((size_t)pl.batch * G_COUNT * g.plane);
  if (pl.tm_state != pl.state.p || pl.tm_state_n != pl.state.n || (calc_id == 1 && pl.tm_gacc != pl.gacc.p)) {
    encode_tma_maps(g, pl.state.p, (long long)pl.batch * S_COUNT, calc_id == 1 ? pl.gacc.p : nullptr,
                    (long long)pl.batch * G_COUNT, pl.model.p, &pl.tm);
    pl.tm_state = pl.state.p;
    pl.tm_state_n = pl.state.n;
    if (calc_id == 1) pl.tm_gacc = pl.gacc.p;
  }
  pl.syn_tr.alloc((size_t)pl.batch * g.nSteps * pl.nrp);
  if (calc_id != 2) {
    pl.res_tr.alloc((size_t)pl.batch * g.nSteps * pl.nrp);
    const int nb = ((g.nSteps + 31) / 32) * ((pl.max_nrec + 31) / 32);
    pl.partial_per_shot = nb;
    pl.partial.alloc((size_t)pl.batch * nb);
  }
  if (calc_id == 1) {
    pl.frames.alloc((size_t)pl.batch * g.nSteps * 5 * g.f_len);
  }
  if (calc_id == 2 || pl.para.save_scratch) pl.syn_rt.alloc((size_t)pl.group * pl.trace_stride);
  if (pl.para.save_scratch && calc_id == 1) {
    pl.res_rt.alloc((size_t)pl.group * pl.trace_stride);
    pl.obs_cond_rt.alloc((size_t)pl.group * pl.trace_stride);
  }
}

// give the per-run buffers back (after an out-of-memory failure, before the batch is planned again)
void release_run_buffers(fwi_b200_plan &pl) {
  pl.state.release(); pl.gacc.release(); pl.frames.release(); pl.syn_tr.release(); pl.res_tr.release();
  pl.partial.release(); pl.syn_rt.release(); pl.res_rt.release(); pl.obs_cond_rt.release();
  pl.tm_state = nullptr; pl.tm_gacc = nullptr; pl.tm_state_n = 0;
}

ShotTables tables_for(fwi_b200_plan &pl, int first) {
  ShotTables st;
  st.src_z = pl.src_z.p + first;
  st.src_x = pl.src_x.p + first;
  st.stf = pl.stf.p + (size_t)first * pl.g.nSteps;
  st.rec_ptr = pl.rec_ptr.p + (size_t)first * (pl.ntiles + 1);
  st.rec_loc = pl.rec_loc.p + (size_t)first * pl.nrp;
  st.rec_id = pl.rec_id.p + (size_t)first * pl.nrp;
  st.nrp = pl.nrp;
  return st;
}

Model model_of(fwi_b200_plan &pl) {
  Model m;
  m.lam = pl.mplane(M_LAM) + pl.g.origin;
  m.mu = pl.mplane(M_MU) + pl.g.origin;
  m.amu = pl.mplane(M_AMU) + pl.g.origin;
  m.bya = pl.mplane(M_BYA) + pl.g.origin;
  m.byb = pl.mplane(M_BYB) + pl.g.origin;
  m.ldt = pl.mplane(M_LDT) + pl.g.origin;
  return m;
}

// ---- overlapped observation loading (host-buffer entry points) ----
void start_obs_load(fwi_b200_plan &pl) {
  if (pl.loader.joinable()) pl.loader.join();
  pl.obs_rt.alloc((size_t)pl.group * pl.trace_stride);
  if (!pl.copy_stream) {
    CUDA_OK(cudaStreamCreateWithFlags(&pl.copy_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&pl.obs_ready, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) CUDA_OK(cudaEventCreateWithFlags(&pl.pin_ev[k], cudaEventDisableTiming));
  }
  if (pl.pin_n < pl.trace_stride) {
    for (int k = 0; k < 2; k++) {
      if (pl.pin[k]) cudaFreeHost(pl.pin[k]);
      pl.pin[k] = nullptr;
      CUDA_OK(cudaMallocHost(reinterpret_cast<void **>(&pl.pin[k]), std::max<size_t>(pl.trace_stride, 1) * sizeof(float)));
    }
    pl.pin_n = pl.trace_stride;
  }
  // the previous evaluation may still be reading obs_rt on the compute stream
  CUDA_OK(cudaStreamSynchronize(pl.stream));
  pl.loader_code = 0;
  pl.loader_err.clear();
  fwi_b200_plan *p = &pl;
  pl.loader = std::thread([p] {
    try {
      if (cudaSetDevice(p->gpu) != cudaSuccess) throw Error(FWI_B200_ERR_CUDA, "loader: cudaSetDevice failed");
      for (int i = 0; i < p->group; i++) {
        const int b = i & 1;
        const size_t n = p->survey.shots[i].z_rec.size() * (size_t)p->g.nSteps;
        if (i >= 2) CUDA_OK(cudaEventSynchronize(p->pin_ev[b]));   // the copy out of this staging buffer is done
        read_f32(p->para.data_dir_name + "/Shot" + std::to_string(p->shot_ids[i]) + ".bin", p->pin[b], n);
        CUDA_OK(cudaMemcpyAsync(p->obs_rt.p + (size_t)i * p->trace_stride, p->pin[b], n * sizeof(float),
                                cudaMemcpyHostToDevice, p->copy_stream));
        CUDA_OK(cudaEventRecord(p->pin_ev[b], p->copy_stream));
      }
      CUDA_OK(cudaEventRecord(p->obs_ready, p->copy_stream));
    } catch (const Error &e) {
      p->loader_code = e.code;
      p->loader_err = e.what();
    } catch (const std::exception &e) {
      p->loader_code = FWI_B200_ERR_IO;
      p->loader_err = e.what();
    }
  });
  for (int i = 0; i < pl.group; i++) pl.obs_set[i] = 2;   // 2 = on its way
}

// called before anything reads obs_rt: the loader has recorded obs_ready (or failed), then stream `s` waits for it
void finish_obs_load(fwi_b200_plan &pl, cudaStream_t s) {
  if (!pl.loader.joinable()) return;
  pl.loader.join();
  if (pl.loader_code != 0) {
    for (int i = 0; i < pl.group; i++) pl.obs_set[i] = 0;
    throw Error(pl.loader_code, pl.loader_err);
  }
  for (int i = 0; i < pl.group; i++) pl.obs_set[i] = 1;
  CUDA_OK(cudaStreamWaitEvent(s, pl.obs_ready, 0));
}

// an asynchronous load that nobody consumed (failed or forward-only call): wait for it before obs_rt is touched again
void settle_obs_load(fwi_b200_plan &pl) {
  if (!pl.loader.joinable()) return;
  pl.loader.join();
  if (pl.copy_stream) cudaStreamSynchronize(pl.copy_stream);
  for (int i = 0; i < pl.group; i++) pl.obs_set[i] = pl.loader_code == 0 ? 1 : 0;
}

// counter pair of the reverse step's dynamic unit claim (zero whenever no reverse launch is in flight: the last CTA of a
// launch rewinds it); nullptr when the option is off
static std::atomic<int> g_dyn_units{FWI_DYN_UNITS};
static int *unit_counter_of(fwi_b200_plan &pl, cudaStream_t s) {
  if (!g_dyn_units.load(std::memory_order_relaxed)) return nullptr;
  if (!pl.unit_counter.p) pl.unit_counter.alloc(2);
  CUDA_OK(cudaMemsetAsync(pl.unit_counter.p, 0, 2 * sizeof(int), s));   // (also heals the pair after an aborted run)
  return pl.unit_counter.p;
}

void run_locked(fwi_b200_plan &pl, int calc_id, cudaStream_t s) {
  const Grid &g = pl.g;
  if (calc_id < 0 || calc_id > 2) throw Error(FWI_B200_ERR_ARG, "invalid calc_id " + std::to_string(calc_id));
  if (!pl.model_set) throw Error(FWI_B200_ERR_ARG, "plan: model not set");
  if (!pl.stf_set) throw Error(FWI_B200_ERR_ARG, "plan: source time functions not set");
  if (calc_id != 2)
    for (int i = 0; i < pl.group; i++)
      if (!pl.obs_set[i]) throw Error(FWI_B200_ERR_ARG, "plan: observed data of shot " + std::to_string(pl.shot_ids[i]) + " not set");
  alloc_run_buffers(pl, calc_id);
  const bool if_res = calc_id != 2, with_adj = calc_id == 1;
  Model m = model_of(pl);
  FwdArgs fa{};
  fa.g = g; fa.m = m; fa.tm = pl.tm;
  fa.pr.z = pl.zprof.p + g.P; fa.pr.x = pl.xprof.p; fa.pr.nxp = g.nx + 2 * XM;
  BwdArgs ba{};
  ba.g = g; ba.m = m; ba.pr = fa.pr; ba.tm = pl.tm;
  // shots per imaging-accumulator slot (the merged backward kernel keeps a slot per shot)
  const int nb_max = std::min(pl.batch, pl.group);
  ba.acc_group = g_merged_bwd.load(std::memory_order_relaxed) ? 1 : reverse_acc_group(g, nb_max, g_dyn_units.load(std::memory_order_relaxed) != 0);
  ba.unit_counter = unit_counter_of(pl, s);

  if (with_adj) {
    CUDA_OK(cudaMemsetAsync(pl.gacc.p, 0, pl.gacc.bytes(), s));
    CUDA_OK(cudaMemsetAsync(pl.stf_grad.p, 0, pl.stf_grad.bytes(), s));
  }
  if (if_res) CUDA_OK(cudaMemsetAsync(pl.j_shot.p, 0, pl.j_shot.bytes(), s));
  const int N = g.nSteps;
  for (int first = 0; first < pl.group; first += pl.batch) {
    const int nb = std::min(pl.batch, pl.group - first);
    ShotTables st = tables_for(pl, first);
    // libCUFD.cu:163-187: zero wavefields, CPML memory and the synthetic traces
    CUDA_OK(cudaMemsetAsync(pl.state.p, 0, (size_t)nb * S_COUNT * g.plane * sizeof(float), s));
    CUDA_OK(cudaMemsetAsync(pl.syn_tr.p, 0, (size_t)nb * N * pl.nrp * sizeof(float), s));
    fa.st = st; fa.state = pl.state.p; fa.traces = pl.syn_tr.p; fa.frames = with_adj ? pl.frames.p : nullptr; fa.batch = nb;
    for (int it = 0; it <= N - 2; it++) {  // libCUFD.cu:202-240
      fa.it = it; fa.cur = it & 1;
      launch_forward_step(fa, with_adj, s);
      pl.launches++;
    }
    int cur_f = (N - 1) & 1;
    pl.cur_f_last = cur_f;
    if (!if_res) {
      for (int k = 0; k < nb; k++) {
        const int nrec = (int)pl.survey.shots[first + k].z_rec.size();
        launch_traces_to_rt(pl.syn_tr.p + (size_t)k * N * pl.nrp, pl.syn_rt.p + (size_t)(first + k) * pl.trace_stride, nrec, pl.nrp, N, s);
        pl.launches++;
      }
      continue;
    }
    // residual + misfit: libCUFD.cu:254-330
    finish_obs_load(pl, s);
    for (int k = 0; k < nb; k++) {
      const int nrec = (int)pl.survey.shots[first + k].z_rec.size();
      ResidualArgs ra{};
      ra.syn_tr = pl.syn_tr.p + (size_t)k * N * pl.nrp;
      ra.obs_rt = pl.obs_rt.p + (size_t)(first + k) * pl.trace_stride;
      ra.w2 = pl.w2.p;
      ra.dt = pl.para.dt;
      if (pl.para.if_win) {
        ra.win_start = pl.win.p + (size_t)(first + k) * 3 * pl.max_nrec;
        ra.win_end = ra.win_start + pl.max_nrec;
        ra.weights = ra.win_end + pl.max_nrec;
      }
      ra.res_tr = pl.res_tr.p + (size_t)k * N * pl.nrp;
      const bool keep = pl.para.save_scratch && with_adj;
      ra.syn_rt = keep ? pl.syn_rt.p + (size_t)(first + k) * pl.trace_stride : nullptr;
      ra.res_rt = keep ? pl.res_rt.p + (size_t)(first + k) * pl.trace_stride : nullptr;
      ra.obs_cond_rt = keep ? pl.obs_cond_rt.p + (size_t)(first + k) * pl.trace_stride : nullptr;
      ra.partial = pl.partial.p + (size_t)k * pl.partial_per_shot;
      ra.nrec = nrec; ra.nrp = pl.nrp; ra.nSteps = N;
      if (nrec > 0) {
        CUDA_OK(cudaMemsetAsync(ra.res_tr, 0, (size_t)N * pl.nrp * sizeof(float), s));
        int nblk = 0;
        launch_residual(ra, &nblk, s);
        launch_sum_partials(ra.partial, nblk, pl.j_shot.p + first + k, s);
        pl.launches += 2;
      }
    }
    if (!with_adj) continue;
    // backward: libCUFD.cu:334-457
    CUDA_OK(cudaMemset2DAsync(pl.state.p + (size_t)S_AA * g.plane, (size_t)S_COUNT * g.plane * sizeof(float), 0,
                              (size_t)(S_COUNT - S_AA) * g.plane * sizeof(float), nb, s));
    ba.st = st; ba.state = pl.state.p; ba.res = pl.res_tr.p; ba.frames = pl.frames.p; ba.gacc = pl.gacc.p;
    ba.stf_grad = pl.stf_grad.p + (size_t)first * N; ba.batch = nb;
    int cur_a = 0;
    if (g_merged_bwd.load(std::memory_order_relaxed)) {
      // ONE launch per time index: the adjoint step of index it+1 (libCUFD.cu:353-373 for the first one, :405-427
      // after) and the reverse step it+1 -> it with imaging (:380-403), which needs exactly the adjoint state that
      // step has just produced -- handed over in registers (bwd_step_kernel)
      for (int it = N - 2; it >= 0; it--) {
        ba.it = it; ba.cur_f = cur_f; ba.cur_a = cur_a;
        launch_backward_merged(ba, s);
        pl.launches++;
        cur_f ^= 1;
        cur_a ^= 1;
      }
      // (the adjoint step of index 0: only its source_grad part is observable; kept for grad_stf[0])
      ba.it = 0; ba.cur_f = cur_f; ba.cur_a = cur_a;
      launch_adjoint_step(ba, s);
      pl.launches++;
    } else {
      ba.it = N - 1; ba.cur_f = cur_f; ba.cur_a = cur_a;   // priming: libCUFD.cu:353-373
      launch_adjoint_step(ba, s);
      pl.launches++;
      cur_a ^= 1;
      for (int it = N - 2; it >= 0; it--) {  // libCUFD.cu:374-445
        ba.it = it; ba.cur_f = cur_f; ba.cur_a = cur_a;
        launch_reverse_imaging(ba, s);
        pl.launches++;
        // (at it == 0 only the source_grad part of this launch is observable; kept for grad_stf[0])
        ba.indep = FWI_ADJ_INDEP;   // follows the reverse step of the same time index: independent of it
        launch_adjoint_step(ba, s);
        ba.indep = 0;
        pl.launches++;
        cur_f ^= 1;
        cur_a ^= 1;
      }
    }
    pl.cur_f_last = cur_f;
  }
  if (if_res) {
    launch_misfit(pl.j_shot.p, pl.group, pl.misfit_half.p, s);
    pl.launches++;
  }
  if (with_adj) {
    launch_finalize(g, pl.gacc.p, (nb_max + ba.acc_group - 1) / ba.acc_group, m.mu, pl.misfit_half.p, pl.result.p, pl.layout, s);
    pl.launches++;
  } else if (if_res) {
    CUDA_OK(cudaMemcpyAsync(pl.result.p + 3LL * g.nz * g.nx, pl.misfit_half.p, sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  CUDA_OK(cudaGetLastError());
  pl.last_calc = calc_id;
}

void check(int rc) {
  if (rc != FWI_B200_OK) throw Error(rc, last_error_cstr());
}

template <typename F>
int guarded(F &&f) {
  try {
    set_last_error("");
    f();
    return FWI_B200_OK;
  } catch (const Error &e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return FWI_B200_ERR_ARG;
  }
}

}  // namespace

// =================================================================================================
// plan API
// =================================================================================================
extern "C" int fwi_b200_plan_create(fwi_b200_plan **out, const char *para_fname, int gpu_id, int group_size,
                                    const int *shot_ids, int max_batch) {
  return guarded([&] {
    if (!out || !para_fname || group_size <= 0 || !shot_ids) throw Error(FWI_B200_ERR_ARG, "plan_create: bad arguments");
    *out = nullptr;
    std::unique_ptr<fwi_b200_plan> pl(new fwi_b200_plan());
    pl->para = read_para(para_fname);
    pl->survey = read_survey(pl->para.survey_fname, pl->para.nPml, group_size, shot_ids, pl->para.if_win);
    pl->group = group_size;
    pl->shot_ids.assign(shot_ids, shot_ids + group_size);
    pl->obs_set.assign(group_size, 0);
    build_grid(pl->para, pl->g);
    use_device(gpu_id);
    pl->gpu = gpu_id;
    configure_kernels();
    CUDA_OK(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    const Grid &g = pl->g;
    pl->model.alloc((size_t)M_COUNT * g.plane);
    CUDA_OK(cudaMemset(pl->model.p, 0, pl->model.bytes()));
    pl->model_in.alloc((size_t)3 * g.nz * g.nx);
    pl->cpmax.alloc(1);
    upload_profiles(*pl);
    upload_tables(*pl);
    pl->trace_stride = (size_t)pl->max_nrec * g.nSteps;
    pl->stf.alloc((size_t)group_size * g.nSteps);
    pl->stf_grad.alloc((size_t)group_size * g.nSteps);
    pl->j_shot.alloc(group_size);
    pl->misfit_half.alloc(1);
    pl->result.alloc((size_t)3 * g.nz * g.nx + 1);
    CUDA_OK(cudaMemset(pl->result.p, 0, pl->result.bytes()));
    CUDA_OK(cudaMemset(pl->misfit_half.p, 0, sizeof(float)));
    CUDA_OK(cudaMemset(pl->stf_grad.p, 0, pl->stf_grad.bytes()));
    pl->max_batch_req = max_batch;
    choose_batch(*pl, max_batch);
    *out = pl.release();
  });
}

extern "C" void fwi_b200_plan_destroy(fwi_b200_plan *plan) { delete plan; }

static void finish_model_locked(fwi_b200_plan *pl);


static void plan_set_model_impl(fwi_b200_plan *pl, const double *Lambda, const double *Mu, const double *Den) {
  if (!pl || !Lambda || !Mu || !Den) throw Error(FWI_B200_ERR_ARG, "set_model: null pointer");
  std::lock_guard<std::mutex> lk(pl->mu);
  use_device(pl->gpu);
  const Grid &g = pl->g;
  const size_t n = (size_t)g.nz * g.nx;
  cudaStream_t s = pl->stream;
  CUDA_OK(cudaMemcpyAsync(pl->model_in.p, Lambda, n * sizeof(double), cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(pl->model_in.p + n, Mu, n * sizeof(double), cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(pl->model_in.p + 2 * n, Den, n * sizeof(double), cudaMemcpyHostToDevice, s));
  pl->vel_masked = -1;
  finish_model_locked(pl);
}

// model_in (lambda, mu [MPa], rho; caller's layout) -> float planes, derived coefficients, Courant check
static void finish_model_locked(fwi_b200_plan *pl) {
  const Grid &g = pl->g;
  const size_t n = (size_t)g.nz * g.nx;
  cudaStream_t s = pl->stream;
  CUDA_OK(cudaMemsetAsync(pl->cpmax.p, 0, sizeof(unsigned int), s));
  launch_model_prep(g, pl->model_in.p, pl->model_in.p + n, pl->model_in.p + 2 * n, pl->model.p, pl->cpmax.p, pl->layout, s);
  pl->launches += 2;
  unsigned int bits = 0;
  CUDA_OK(cudaMemcpyAsync(&bits, pl->cpmax.p, sizeof(bits), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  float cpmax;
  std::memcpy(&cpmax, &bits, sizeof(float));
  const float cn = courant_number(cpmax, pl->para.dt, pl->para.dz, pl->para.dx);
  pl->model_set = false;
  if (cn > 1.0f)  // utilities.cu:239 exits silently; we report
    throw Error(FWI_B200_ERR_CFL, "Courant number " + std::to_string(cn) + " > 1 (max cp " + std::to_string(cpmax) + ")");
  pl->model_set = true;
}

extern "C" int fwi_b200_plan_set_model(fwi_b200_plan *pl, const double *Lambda, const double *Mu, const double *Den) {
  return guarded([&] { plan_set_model_impl(pl, Lambda, Mu, Den); });
}

static void plan_set_stf_impl(fwi_b200_plan *pl, const double *stf) {
  if (!pl || !stf) throw Error(FWI_B200_ERR_ARG, "set_stf: null pointer");
  std::lock_guard<std::mutex> lk(pl->mu);
  use_device(pl->gpu);
  const int N = pl->g.nSteps;
  std::vector<float> w2;
  const bool ok = taper_weights(N, pl->para.dt, 0.001f, w2);  // Src_Rec.cu:140
  std::vector<float> h((size_t)pl->group * N);
  for (int i = 0; i < pl->group; i++) {
    if (pl->shot_ids[i] < 0) throw Error(FWI_B200_ERR_ARG, "negative shot id");
    const double *row = stf + (size_t)pl->shot_ids[i] * N;  // row = GLOBAL shot id (Src_Rec.cu:135)
    for (int t = 0; t < N; t++) {
      float v = (float)row[t];
      if (ok) v *= w2[t];
      h[(size_t)i * N + t] = v;
    }
  }
  CUDA_OK(cudaMemcpyAsync(pl->stf.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, pl->stream));
  CUDA_OK(cudaStreamSynchronize(pl->stream));
  pl->stf_set = true;
}

extern "C" int fwi_b200_plan_set_velocities(fwi_b200_plan *pl, const double *cp, const double *cs, const double *rho,
                                            const double *cp_ref, const double *cs_ref, const double *rho_ref, int is_masked,
                                            int padded) {
  return guarded([&] {
    if (!pl || !cp || !cs || !rho) throw Error(FWI_B200_ERR_ARG, "set_velocities: null pointer");
    if (!is_masked && !(cp_ref && cs_ref && rho_ref))
      throw Error(FWI_B200_ERR_ARG, "set_velocities: cp_ref, cs_ref, rho_ref are required when is_masked is 0 (src/FWI.jl:174-176)");
    std::lock_guard<std::mutex> lk(pl->mu);
    use_device(pl->gpu);
    const Grid &g = pl->g;
    const long long nz0 = g.nz - 2 * g.nPml - g.nPad, nx0 = g.nx - 2 * g.nPml;
    if (nz0 <= 0 || nx0 <= 0) throw Error(FWI_B200_ERR_GEOM, "set_velocities: no cells between the absorbing layers");
    const long long n = (long long)g.nz * g.nx, n_in = padded ? n : nz0 * nx0;
    pl->vel_in.alloc((size_t)6 * n_in);
    pl->vel.alloc((size_t)3 * n);
    cudaStream_t s = pl->stream;
    const double *src[6] = {cp, cs, rho, cp_ref, cs_ref, rho_ref};
    for (int k = 0; k < (is_masked ? 3 : 6); k++)
      CUDA_OK(cudaMemcpyAsync(pl->vel_in.p + k * n_in, src[k], n_in * sizeof(double), cudaMemcpyHostToDevice, s));
    launch_velocity_prep(g, pl->layout, padded ? 1 : 0, is_masked ? 1 : 0, pl->vel_in.p, n_in, pl->vel.p, pl->model_in.p, s);
    pl->launches++;
    pl->vel_masked = is_masked ? 1 : 0;
    finish_model_locked(pl);
  });
}

extern "C" int fwi_b200_plan_get_velocity_gradients(fwi_b200_plan *pl, double *misfit, double *g_cp, double *g_cs,
                                                    double *g_rho) {
  return guarded([&] {
    if (!pl || !g_cp || !g_cs || !g_rho) throw Error(FWI_B200_ERR_ARG, "get_velocity_gradients: null pointer");
    std::lock_guard<std::mutex> lk(pl->mu);
    use_device(pl->gpu);
    if (pl->vel_masked < 0 || pl->last_calc != 1)
      throw Error(FWI_B200_ERR_ARG, "get_velocity_gradients: needs set_velocities followed by a calc_id 1 run");
    CUDA_OK(cudaDeviceSynchronize());
    const Grid &g = pl->g;
    const size_t n = (size_t)g.nz * g.nx;
    pl->vel_grad.alloc(3 * n);
    launch_velocity_grad(g, pl->layout, pl->vel_masked, pl->result.p, pl->vel.p, pl->vel_grad.p, pl->stream);
    pl->launches++;
    CUDA_OK(cudaMemcpyAsync(g_cp, pl->vel_grad.p, n * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
    CUDA_OK(cudaMemcpyAsync(g_cs, pl->vel_grad.p + n, n * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
    CUDA_OK(cudaMemcpyAsync(g_rho, pl->vel_grad.p + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
    float m = 0;
    CUDA_OK(cudaMemcpyAsync(&m, pl->result.p + 3 * n, sizeof(float), cudaMemcpyDeviceToHost, pl->stream));
    CUDA_OK(cudaStreamSynchronize(pl->stream));
    if (misfit) *misfit = m;
  });
}

extern "C" int fwi_b200_plan_set_layout(fwi_b200_plan *pl, int layout) {
  return guarded([&] {
    if (!pl || (layout != 0 && layout != 1)) throw Error(FWI_B200_ERR_ARG, "set_layout: layout must be 0 (row-major) or 1 (column-major)");
    std::lock_guard<std::mutex> lk(pl->mu);
    if (pl->layout != layout) pl->model_set = false;   // the resident model was converted under the other convention
    pl->layout = layout;
  });
}

extern "C" int fwi_b200_plan_set_stf(fwi_b200_plan *pl, const double *stf) {
  return guarded([&] { plan_set_stf_impl(pl, stf); });
}

extern "C" int fwi_b200_plan_set_obs(fwi_b200_plan *pl, int ishot, const float *obs) {
  return guarded([&] {
    if (!pl || !obs || ishot < 0 || ishot >= pl->group) throw Error(FWI_B200_ERR_ARG, "set_obs: bad arguments");
    std::lock_guard<std::mutex> lk(pl->mu);
    use_device(pl->gpu);
    settle_obs_load(*pl);
    pl->obs_rt.alloc((size_t)pl->group * pl->trace_stride);
    const size_t n = pl->survey.shots[ishot].z_rec.size() * (size_t)pl->g.nSteps;
    CUDA_OK(cudaMemcpyAsync(pl->obs_rt.p + (size_t)ishot * pl->trace_stride, obs, n * sizeof(float), cudaMemcpyHostToDevice, pl->stream));
    CUDA_OK(cudaStreamSynchronize(pl->stream));
    pl->obs_set[ishot] = 1;
  });
}

extern "C" int fwi_b200_plan_load_obs_files(fwi_b200_plan *pl) {
  return guarded([&] {
    if (!pl) throw Error(FWI_B200_ERR_ARG, "load_obs_files: null plan");
    std::lock_guard<std::mutex> lk(pl->mu);
    use_device(pl->gpu);
    settle_obs_load(*pl);
    pl->obs_rt.alloc((size_t)pl->group * pl->trace_stride);
    std::vector<float> h(pl->trace_stride);
    for (int i = 0; i < pl->group; i++) {
      const size_t n = pl->survey.shots[i].z_rec.size() * (size_t)pl->g.nSteps;
      read_f32(pl->para.data_dir_name + "/Shot" + std::to_string(pl->shot_ids[i]) + ".bin", h.data(), n);
      CUDA_OK(cudaMemcpyAsync(pl->obs_rt.p + (size_t)i * pl->trace_stride, h.data(), n * sizeof(float), cudaMemcpyHostToDevice, pl->stream));
      CUDA_OK(cudaStreamSynchronize(pl->stream));
      pl->obs_set[i] = 1;
    }
  });
}

static void plan_run_impl(fwi_b200_plan *pl, int calc_id, void *stream, int sync) {
  if (!pl) throw Error(FWI_B200_ERR_ARG, "run: null plan");
  std::lock_guard<std::mutex> lk(pl->mu);
  use_device(pl->gpu);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : pl->stream;
  run_locked(*pl, calc_id, s);
  if (sync) CUDA_OK(cudaStreamSynchronize(s));
}

extern "C" int fwi_b200_plan_run(fwi_b200_plan *pl, int calc_id, void *stream, int sync) {
  return guarded([&] { plan_run_impl(pl, calc_id, stream, sync); });
}

extern "C" float *fwi_b200_plan_result_device(fwi_b200_plan *pl) { return pl ? pl->result.p : nullptr; }
extern "C" size_t fwi_b200_plan_result_count(fwi_b200_plan *pl) { return pl ? (size_t)3 * pl->g.nz * pl->g.nx + 1 : 0; }

static void plan_get_result_impl(fwi_b200_plan *pl, double *misfit, double *gl, double *gm, double *gd, double *gs) {
  if (!pl) throw Error(FWI_B200_ERR_ARG, "get_result: null plan");
  std::lock_guard<std::mutex> lk(pl->mu);
  use_device(pl->gpu);
  CUDA_OK(cudaDeviceSynchronize());
  const size_t n = (size_t)pl->g.nz * pl->g.nx;
  if (gl || gm || gd) {
    std::vector<float> h(3 * n + 1);
    CUDA_OK(cudaMemcpy(h.data(), pl->result.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    if (gl) for (size_t i = 0; i < n; i++) gl[i] = h[i];
    if (gm) for (size_t i = 0; i < n; i++) gm[i] = h[n + i];
    if (gd) for (size_t i = 0; i < n; i++) gd[i] = h[2 * n + i];
    if (misfit) *misfit = h[3 * n];
  } else if (misfit) {
    float m = 0;
    CUDA_OK(cudaMemcpy(&m, pl->result.p + 3 * n, sizeof(float), cudaMemcpyDeviceToHost));
    *misfit = m;
  }
  if (gs) {
    std::vector<float> h((size_t)pl->group * pl->g.nSteps);
    CUDA_OK(cudaMemcpy(h.data(), pl->stf_grad.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); i++) gs[i] = h[i];
  }
}

extern "C" int fwi_b200_plan_get_result(fwi_b200_plan *pl, double *misfit, double *gl, double *gm, double *gd, double *gs) {
  return guarded([&] { plan_get_result_impl(pl, misfit, gl, gm, gd, gs); });
}

extern "C" int fwi_b200_plan_get_traces(fwi_b200_plan *pl, int ishot, int which, float *out) {
  return guarded([&] {
    if (!pl || !out || ishot < 0 || ishot >= pl->group) throw Error(FWI_B200_ERR_ARG, "get_traces: bad arguments");
    std::lock_guard<std::mutex> lk(pl->mu);
    use_device(pl->gpu);
    CUDA_OK(cudaDeviceSynchronize());
    DevBuf<float> *src = which == 0 ? &pl->syn_rt : which == 1 ? &pl->res_rt : which == 2 ? &pl->obs_cond_rt : nullptr;
    if (!src || !src->p) throw Error(FWI_B200_ERR_ARG, "get_traces: these traces were not kept by the last run");
    const size_t n = pl->survey.shots[ishot].z_rec.size() * (size_t)pl->g.nSteps;
    CUDA_OK(cudaMemcpy(out, src->p + (size_t)ishot * pl->trace_stride, n * sizeof(float), cudaMemcpyDeviceToHost));
  });
}

static void write_group_files(fwi_b200_plan *pl, DevBuf<float> &buf, const std::string &dir, const std::string &stem) {
  std::vector<float> h(pl->trace_stride);
  for (int i = 0; i < pl->group; i++) {
    const size_t n = pl->survey.shots[i].z_rec.size() * (size_t)pl->g.nSteps;
    CUDA_OK(cudaMemcpy(h.data(), buf.p + (size_t)i * pl->trace_stride, n * sizeof(float), cudaMemcpyDeviceToHost));
    write_f32(dir + "/" + stem + std::to_string(pl->shot_ids[i]) + ".bin", h.data(), n);
  }
}

static void plan_write_obs_files_impl(fwi_b200_plan *pl) {
  if (!pl) throw Error(FWI_B200_ERR_ARG, "write_obs_files: null plan");
  std::lock_guard<std::mutex> lk(pl->mu);
  use_device(pl->gpu);
  CUDA_OK(cudaDeviceSynchronize());
  if (pl->last_calc != 2 || !pl->syn_rt.p) throw Error(FWI_B200_ERR_ARG, "write_obs_files: last run was not calc_id 2");
  write_group_files(pl, pl->syn_rt, pl->para.data_dir_name, "Shot");  // libCUFD.cu:514-521
}

extern "C" int fwi_b200_plan_write_obs_files(fwi_b200_plan *pl) {
  return guarded([&] { plan_write_obs_files_impl(pl); });
}

extern "C" int fwi_b200_plan_info(fwi_b200_plan *pl, int *nz, int *nx, int *nSteps, int *nPml, int *nPad, int *group_size,
                                  int *batch, int *max_nrec) {
  if (!pl) return FWI_B200_ERR_ARG;
  if (nz) *nz = pl->g.nz;
  if (nx) *nx = pl->g.nx;
  if (nSteps) *nSteps = pl->g.nSteps;
  if (nPml) *nPml = pl->g.nPml;
  if (nPad) *nPad = pl->g.nPad;
  if (group_size) *group_size = pl->group;
  if (batch) *batch = pl->batch;
  if (max_nrec) *max_nrec = pl->max_nrec;
  return FWI_B200_OK;
}

extern "C" int fwi_b200_plan_shot_geometry(fwi_b200_plan *pl, int ishot, int *z_src, int *x_src, int *nrec, int *z_rec, int *x_rec) {
  if (!pl || ishot < 0 || ishot >= pl->group) return FWI_B200_ERR_ARG;
  const Shot &s = pl->survey.shots[ishot];
  if (z_src) *z_src = s.z_src;
  if (x_src) *x_src = s.x_src;
  if (nrec) *nrec = (int)s.z_rec.size();
  if (z_rec) std::copy(s.z_rec.begin(), s.z_rec.end(), z_rec);
  if (x_rec) std::copy(s.x_rec.begin(), s.x_rec.end(), x_rec);
  return FWI_B200_OK;
}

extern "C" long long fwi_b200_plan_launch_count(fwi_b200_plan *pl) { return pl ? pl->launches : 0; }

extern "C" int fwi_b200_plan_get_field(fwi_b200_plan *pl, int ishot, int field, float *out) {
  return guarded([&] {
    if (!pl || !out || field < 0 || field > 4 || ishot < 0 || ishot >= pl->batch) throw Error(FWI_B200_ERR_ARG, "get_field: bad arguments");
    std::lock_guard<std::mutex> lk(pl->mu);
    use_device(pl->gpu);
    CUDA_OK(cudaDeviceSynchronize());
    const Grid &g = pl->g;
    std::vector<float> h((size_t)g.plane);
    const int slot = (pl->cur_f_last ? S_FB : S_FA) + field;
    CUDA_OK(cudaMemcpy(h.data(), pl->state.p + ((size_t)ishot * S_COUNT + slot) * g.plane, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int z = 0; z < g.nz; z++)
      for (int x = 0; x < g.nx; x++) out[(size_t)z * g.nx + x] = h[g.origin + (size_t)x * g.P + z];
  });
}

extern "C" int fwi_b200_plan_time_kernel(fwi_b200_plan *pl, int which, int iters, void *stream, float *ms_per_launch,
                                         double *alg_bytes) {
  return guarded([&] {
    if (!pl || iters <= 0 || which < 0 || which > 4) throw Error(FWI_B200_ERR_ARG, "time_kernel: bad arguments");
    std::lock_guard<std::mutex> lk(pl->mu);
    use_device(pl->gpu);
    const Grid &g = pl->g;
    alloc_run_buffers(*pl, 1);
    settle_obs_load(*pl);
    if (!pl->obs_rt.p) pl->obs_rt.alloc((size_t)pl->group * pl->trace_stride);
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : pl->stream;
    const int nb = std::min(pl->batch, pl->group);
    FwdArgs fa{};
    fa.g = g; fa.m = model_of(*pl); fa.tm = pl->tm;
    fa.pr.z = pl->zprof.p + g.P; fa.pr.x = pl->xprof.p; fa.pr.nxp = g.nx + 2 * XM;
    fa.st = tables_for(*pl, 0); fa.state = pl->state.p; fa.traces = pl->syn_tr.p; fa.frames = pl->frames.p; fa.batch = nb;
    BwdArgs ba{};
    ba.g = g; ba.m = fa.m; ba.pr = fa.pr; ba.tm = pl->tm; ba.st = fa.st; ba.state = pl->state.p; ba.res = pl->res_tr.p;
    ba.frames = pl->frames.p; ba.gacc = pl->gacc.p; ba.stf_grad = pl->stf_grad.p; ba.batch = nb;
    ba.acc_group = which == 2 ? reverse_acc_group(g, nb, g_dyn_units.load(std::memory_order_relaxed) != 0) : 1;
    ba.unit_counter = unit_counter_of(*pl, s);
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    auto one = [&](int k) {
      const int it = 1 + (k % std::max(1, g.nSteps - 3));
      if (which <= 1) { fa.it = it; fa.cur = k & 1; launch_forward_step(fa, which == 1, s); }
      else if (which == 2) { ba.it = it; ba.cur_f = k & 1; ba.cur_a = 0; launch_reverse_imaging(ba, s); }
      else if (which == 3) { ba.it = it; ba.cur_f = 0; ba.cur_a = k & 1; launch_adjoint_step(ba, s); }
      else { ba.it = it; ba.cur_f = k & 1; ba.cur_a = k & 1; launch_backward_merged(ba, s); }
    };
    for (int k = 0; k < 3; k++) one(k);
    CUDA_OK(cudaEventRecord(e0, s));
    for (int k = 0; k < iters; k++) one(k);
    CUDA_OK(cudaEventRecord(e1, s));
    CUDA_OK(cudaEventSynchronize(e1));
    pl->launches += iters + 3;
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_launch) *ms_per_launch = ms / iters;
    if (alg_bytes) {
      // algorithmic bytes per cell-update
      const double cells = (double)g.nz * g.nx * nb;
      const double box = (double)(g.zhi - g.zlo + 1) * (g.xhi - g.xlo + 1) * nb;
      const double fz = 2.0 * g.nPml / g.nz, fx = 2.0 * g.nPml / g.nx;
      double b = 0;
      // 60 B forward, 124 B backward = 64 (reverse + imaging) + 60 (adjoint)
      if (which <= 1) b = cells * (60.0 + 32.0 * (fz + fx));
      else if (which == 2) b = box * 64.0;         // forward state R+W (40) + three gradient accumulators R+W (24)
      else if (which == 3) b = cells * (60.0 + 64.0 * (fz + fx));  // adjoint state R+W, 5 coefficients, psi/phi in the strips
      else b = cells * (60.0 + 64.0 * (fz + fx)) + box * 64.0;     // merged backward launch: both of the above = 124 B
      if (which == 1) b += (double)nb * 5 * g.f_len * 4.0;
      *alg_bytes = b;
    }
  });
}

extern "C" int fwi_b200_grid_info(const char *para_fname, int *out) {
  return guarded([&] {
    if (!para_fname || !out) throw Error(FWI_B200_ERR_ARG, "grid_info: null pointer");
    const Para p = read_para(para_fname);
    Grid g{};
    build_grid(p, g);
    const int v[12] = {g.nz, g.nx, g.P, g.zlive, g.z_off, g.tiles_z, g.tiles_x, g.f_len, g.zlo, g.zhi, g.xlo, g.xhi};
    std::copy(v, v + 12, out);
  });
}

extern "C" const char *fwi_b200_version(void) {
  return "{\"name\":\"fwi_b200\",\"abi\":2,\"arch\":\"sm_100a\",\"tile\":[56,28],\"fp64_promote\":false,"
         "\"if_win\":true,\"multi_gpu\":true}";
}

extern "C" const char *fwi_b200_last_error(void) { return last_error_cstr(); }

// =================================================================================================
// reference-compatible host-buffer entry points, on top of a small plan cache
// =================================================================================================
namespace {

struct CacheEntry {
  int gpu;
  std::string key;
  std::shared_ptr<fwi_b200_plan> plan;   // shared: a call in flight keeps its plan alive if another thread evicts the entry
};
std::mutex g_cache_mu;
std::list<CacheEntry> g_cache;           // most recently used first
constexpr size_t kCachePerGpu = 8;       // per device: baseline + 5 monitor surveys of a time-lapse inversion stay resident (C4)

std::string cache_key(const char *para_fname, const Para &p, const std::string &survey_text, int group, const int *ids) {
  std::string k = std::string(para_fname) + "\n" + p.text + "\n" + survey_text + "\n" +
                  std::to_string(g_frame_ring.load(std::memory_order_relaxed)) + "\n";
  for (int i = 0; i < group; i++) k += std::to_string(ids[i]) + ",";
  return k;
}

// g_cache_mu held.  Drops cached plans of `gpu` that no call holds (use_count() == 1), least recently used first, until
// at most `max_left` entries of that device remain.  Their device memory is returned by the plan destructor.
size_t evict_idle_locked(int gpu, size_t max_left) {
  size_t n = 0, freed = 0;
  for (const CacheEntry &e : g_cache) n += e.gpu == gpu;
  auto it = g_cache.end();
  while (it != g_cache.begin() && n > max_left) {
    --it;
    if (it->gpu == gpu && it->plan.use_count() == 1) {
      it = g_cache.erase(it);
      --n;
      ++freed;
    }
  }
  return freed;
}

std::shared_ptr<fwi_b200_plan> cached_plan(const char *para_fname, int gpu, int group, const int *ids) {
  Para p = read_para(para_fname);
  std::string survey_text;
  {
    FILE *fp = std::fopen(p.survey_fname.c_str(), "rb");
    if (!fp) throw Error(FWI_B200_ERR_IO, "cannot open survey file '" + p.survey_fname + "'");
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, fp)) > 0) survey_text.append(buf, n);
    std::fclose(fp);
  }
  const std::string key = cache_key(para_fname, p, survey_text, group, ids);
  std::lock_guard<std::mutex> lk(g_cache_mu);
  for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
    if (it->gpu == gpu && it->key == key) {
      g_cache.splice(g_cache.begin(), g_cache, it);
      return g_cache.front().plan;
    }
  evict_idle_locked(gpu, kCachePerGpu - 1);
  fwi_b200_plan *raw = nullptr;
  int rc = fwi_b200_plan_create(&raw, para_fname, gpu, group, ids, 0);
  if (rc == FWI_B200_ERR_CUDA && evict_idle_locked(gpu, 0) > 0)   // idle plans were holding the memory: drop them, plan again
    rc = fwi_b200_plan_create(&raw, para_fname, gpu, group, ids, 0);
  if (rc != FWI_B200_OK) throw Error(rc, last_error_cstr());
  std::shared_ptr<fwi_b200_plan> plan(raw);
  // a batch squeezed by what idle cached plans still hold: give their memory back and size the batch again
  if (plan->batch < std::min(plan->group, 64) && evict_idle_locked(gpu, 0) > 0) choose_batch(*plan, 0);
  g_cache.push_front(CacheEntry{gpu, key, plan});
  return plan;
}

// one evaluation through a cached plan with host buffers.  fetch == false leaves the result on the device (the
// multi-GPU driver reduces it there); the caller holds pl->call_mu.
void host_eval(fwi_b200_plan *pl, const double *Lambda, const double *Mu, const double *Den, const double *stf,
               int calc_id, bool sync) {
  plan_set_model_impl(pl, Lambda, Mu, Den);
  plan_set_stf_impl(pl, stf);
  for (int attempt = 0;; attempt++) {
    try {
      if (calc_id != 2) {   // Data/Shot<id>.bin -> device, overlapped with the forward time loop
        std::lock_guard<std::mutex> lk(pl->mu);
        use_device(pl->gpu);
        start_obs_load(*pl);
      }
      plan_run_impl(pl, calc_id, nullptr, sync ? 1 : 0);
      return;
    } catch (const OomError &) {
      // Buffers are allocated lazily, so plans created back to back budget the same free memory; and idle cached plans
      // of this device may hold most of it.  Drop the idle ones, return this plan's own run buffers, size the batch
      // for what is free now, and try once more.
      if (attempt) throw;
      {
        std::lock_guard<std::mutex> lk(pl->mu);
        use_device(pl->gpu);
        settle_obs_load(*pl);
        CUDA_OK(cudaDeviceSynchronize());
        release_run_buffers(*pl);
      }
      {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        evict_idle_locked(pl->gpu, 0);
      }
      std::lock_guard<std::mutex> lk(pl->mu);
      use_device(pl->gpu);
      choose_batch(*pl, pl->max_batch_req);
    }
  }
}

void write_scratch(fwi_b200_plan *pl) {   // libCUFD.cu:493-511
  std::lock_guard<std::mutex> lk(pl->mu);
  write_group_files(pl, pl->res_rt, pl->para.scratch_dir_name, "Residual_Shot");
  write_group_files(pl, pl->syn_rt, pl->para.scratch_dir_name, "Syn_Shot");
  write_group_files(pl, pl->obs_cond_rt, pl->para.scratch_dir_name, "CondObs_Shot");
  std::vector<float> h((size_t)pl->group * pl->g.nSteps);
  CUDA_OK(cudaMemcpy(h.data(), pl->stf.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (int i = 0; i < pl->group; i++)
    write_f32(pl->para.scratch_dir_name + "/src_updated" + std::to_string(pl->shot_ids[i]) + ".bin",
              h.data() + (size_t)i * pl->g.nSteps, pl->g.nSteps);
}

int host_call(double *misfit, double *gl, double *gm, double *gd, double *gs, const double *Lambda, const double *Mu,
              const double *Den, const double *stf, int calc_id, bool also_misfit, int gpu_id, int group_size,
              const int *shot_ids, const char *para_fname, int layout = 0) {
  return guarded([&] {
    if (layout != 0 && layout != 1) throw Error(FWI_B200_ERR_ARG, "layout must be 0 (row-major) or 1 (column-major)");
    if (calc_id < 0 || calc_id > 2) throw Error(FWI_B200_ERR_ARG, "Invalid calc_id " + std::to_string(calc_id));
    if (!Lambda || !Mu || !Den || !stf || !shot_ids || !para_fname || group_size <= 0)
      throw Error(FWI_B200_ERR_ARG, "cufd: null input");
    if (calc_id == 1 && !(gl && gm && gd && gs) && !also_misfit)
      throw Error(FWI_B200_ERR_ARG, "cufd: calc_id 1 needs the four gradient outputs");
    const std::shared_ptr<fwi_b200_plan> hold = cached_plan(para_fname, gpu_id, group_size, shot_ids);
    fwi_b200_plan *pl = hold.get();
    // one caller at a time per cached plan, for the WHOLE sequence: two threads with the same (para, gpu, shot ids) but
    // different models must not interleave set_model / run / get_result
    std::lock_guard<std::mutex> call(pl->call_mu);
    check(fwi_b200_plan_set_layout(pl, layout));
    host_eval(pl, Lambda, Mu, Den, stf, calc_id, true);
    if (calc_id == 2) {
      plan_write_obs_files_impl(pl);
      if (misfit) *misfit = 0.0;  // FwiOp.cpp:316
    } else if (calc_id == 0) {
      plan_get_result_impl(pl, misfit, nullptr, nullptr, nullptr, nullptr);
    } else {
      // the reference never writes *misfit for calc_id 1 (libCUFD.cu:528); the fused entry point does
      plan_get_result_impl(pl, also_misfit ? misfit : nullptr, gl, gm, gd, gs);
      if (pl->para.save_scratch) write_scratch(pl);
    }
  });
}

}  // namespace

extern "C" int fwi_b200_cufd(double *misfit, double *gl, double *gm, double *gd, double *gs, const double *Lambda,
                             const double *Mu, const double *Den, const double *stf, int calc_id, int gpu_id,
                             int group_size, const int *shot_ids, const char *para_fname) {
  return host_call(misfit, gl, gm, gd, gs, Lambda, Mu, Den, stf, calc_id, false, gpu_id, group_size, shot_ids, para_fname);
}

extern "C" int fwi_b200_cufd_ex(double *misfit, double *gl, double *gm, double *gd, double *gs, const double *Lambda,
                                const double *Mu, const double *Den, const double *stf, int calc_id, int gpu_id,
                                int group_size, const int *shot_ids, const char *para_fname, int layout, int with_misfit) {
  return host_call(misfit, gl, gm, gd, gs, Lambda, Mu, Den, stf, calc_id, with_misfit != 0, gpu_id, group_size, shot_ids,
                   para_fname, layout);
}

extern "C" int fwi_b200_forward(double *misfit, const double *Lambda, const double *Mu, const double *Den, const double *stf,
                                int gpu_id, int group_size, const int *shot_ids, const char *para_fname) {
  return host_call(misfit, nullptr, nullptr, nullptr, nullptr, Lambda, Mu, Den, stf, 0, false, gpu_id, group_size, shot_ids, para_fname);
}

extern "C" int fwi_b200_backward(double *gl, double *gm, double *gd, double *gs, const double *Lambda, const double *Mu,
                                 const double *Den, const double *stf, int gpu_id, int group_size, const int *shot_ids,
                                 const char *para_fname) {
  return host_call(nullptr, gl, gm, gd, gs, Lambda, Mu, Den, stf, 1, false, gpu_id, group_size, shot_ids, para_fname);
}

extern "C" int fwi_b200_obscalc(double *misfit, const double *Lambda, const double *Mu, const double *Den, const double *stf,
                                int gpu_id, int group_size, const int *shot_ids, const char *para_fname) {
  return host_call(misfit, nullptr, nullptr, nullptr, nullptr, Lambda, Mu, Den, stf, 2, false, gpu_id, group_size, shot_ids, para_fname);
}

extern "C" int fwi_b200_misfit_and_gradient(double *misfit, double *gl, double *gm, double *gd, double *gs,
                                            const double *Lambda, const double *Mu, const double *Den, const double *stf,
                                            int gpu_id, int group_size, const int *shot_ids, const char *para_fname) {
  return host_call(misfit, gl, gm, gd, gs, Lambda, Mu, Den, stf, 1, true, gpu_id, group_size, shot_ids, para_fname);
}

// =================================================================================================
// several GPUs inside ONE process: shots sharded over the devices, gradients summed with NCCL
// =================================================================================================
namespace {

// NCCL is bound at run time (dlopen of libnccl.so.2): inside a process that already carries a copy (PyTorch bundles
// one) that copy is the one found, and the library has no link-time dependency a single-GPU user would have to satisfy.
struct NcclApi {
  void *h = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    auto sym = [&](const char *n) {
      void *p = dlsym(api.h, n);
      if (!p && err.empty()) err = std::string("libnccl: missing symbol ") + n;
      return p;
    };
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  });
  if (!err.empty()) throw Error(FWI_B200_ERR_CUDA, "gradient_multi needs NCCL for more than one device: " + err);
  return api;
}

#define NCCL_OK(call)                                                                                       \
  do {                                                                                                      \
    ncclResult_t r_ = (call);                                                                               \
    if (r_ != ncclSuccess)                                                                                  \
      throw Error(FWI_B200_ERR_CUDA, std::string("NCCL: ") + nccl_api().GetErrorString(r_) + " (" #call ")"); \
  } while (0)

// one communicator set per ordered device list, created once (ncclCommInitAll) and kept until fwi_b200_release()
struct CommSet {
  std::vector<int> devs;
  std::vector<ncclComm_t> comms;
};
std::mutex g_comm_mu;            // also serialises the collectives of concurrent gradient_multi calls
std::list<CommSet> g_comms;

CommSet &comm_set_locked(const std::vector<int> &devs) {
  for (CommSet &c : g_comms)
    if (c.devs == devs) return c;
  CommSet c;
  c.devs = devs;
  c.comms.resize(devs.size());
  NCCL_OK(nccl_api().CommInitAll(c.comms.data(), (int)devs.size(), devs.data()));
  g_comms.push_back(std::move(c));
  return g_comms.back();
}

void destroy_comms() {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  for (CommSet &c : g_comms)
    for (ncclComm_t m : c.comms) nccl_api().CommDestroy(m);
  g_comms.clear();
}

struct Shard {
  int gpu = 0;
  std::vector<int> ids, pos;            // shot ids and their positions in the caller's group
  std::shared_ptr<fwi_b200_plan> plan;
  std::unique_lock<std::mutex> call;    // the plan's call lock, held from the evaluation to the last copy
  int rc = FWI_B200_OK;
  std::string err;
};

}  // namespace

extern "C" int fwi_b200_gradient_multi(double *misfit, double *gl, double *gm, double *gd, double *gs,
                                       const double *Lambda, const double *Mu, const double *Den, const double *stf,
                                       int ngpu, const int *gpu_ids, int group_size, const int *shot_ids,
                                       const char *para_fname) {
  return guarded([&] {
    if (ngpu <= 0 || !gpu_ids || group_size <= 0 || !shot_ids || !para_fname || !Lambda || !Mu || !Den || !stf)
      throw Error(FWI_B200_ERR_ARG, "gradient_multi: bad arguments");
    for (int a = 0; a < ngpu; a++)
      for (int b = a + 1; b < ngpu; b++)
        if (gpu_ids[a] == gpu_ids[b]) throw Error(FWI_B200_ERR_ARG, "gradient_multi: duplicate gpu id");
    const Para para = read_para(para_fname);
    const size_t n = (size_t)para.nz * para.nx;
    const int N = para.nSteps;
    std::vector<Shard> shards(std::min(ngpu, group_size));
    for (int k = 0; k < group_size; k++) {   // round-robin, a true partition of the group
      Shard &sh = shards[k % shards.size()];
      sh.ids.push_back(shot_ids[k]);
      sh.pos.push_back(k);
    }
    const bool reduce = shards.size() > 1;
    if (reduce) nccl_api();   // fail before any work if NCCL cannot be loaded
    // the cached plan of every shard, and its call lock (taken and released by THIS thread, in device order)
    for (size_t r = 0; r < shards.size(); r++) {
      Shard &sh = shards[r];
      sh.gpu = gpu_ids[r];
      sh.plan = cached_plan(para_fname, sh.gpu, (int)sh.ids.size(), sh.ids.data());
      sh.call = std::unique_lock<std::mutex>(sh.plan->call_mu);
      check(fwi_b200_plan_set_layout(sh.plan.get(), 0));
    }
    // one host thread per device: inputs H2D, observations, forward + backward enqueued on the plan's stream.
    // Nothing synchronises: the gradient of the shard stays in the plan's packed result buffer.
    std::vector<std::thread> workers;
    for (size_t r = 0; r < shards.size(); r++) {
      Shard &sh = shards[r];
      workers.emplace_back([&sh, Lambda, Mu, Den, stf] {
        sh.rc = guarded([&] { host_eval(sh.plan.get(), Lambda, Mu, Den, stf, 1, false); });
        if (sh.rc != FWI_B200_OK) sh.err = last_error_cstr();   // the error text is thread-local
      });
    }
    for (auto &w : workers) w.join();
    for (const Shard &sh : shards)
      if (sh.rc != FWI_B200_OK) throw Error(sh.rc, "gpu " + std::to_string(sh.gpu) + ": " + sh.err);
    if (reduce) {
      // ONE ncclAllReduce(sum, float32, 3 nz nx + 1) over [grad_Lambda | grad_Mu | grad_Den | misfit], in place in the
      // buffer finalize_kernel wrote, on each plan's own stream right behind its last kernel (NVLink / NVSwitch)
      std::vector<int> devs;
      for (const Shard &sh : shards) devs.push_back(sh.gpu);
      std::lock_guard<std::mutex> lk(g_comm_mu);
      CommSet &cs = comm_set_locked(devs);
      NcclApi &nc = nccl_api();
      NCCL_OK(nc.GroupStart());
      for (size_t r = 0; r < shards.size(); r++) {
        fwi_b200_plan *pl = shards[r].plan.get();
        NCCL_OK(nc.AllReduce(pl->result.p, pl->result.p, 3 * n + 1, ncclFloat32, ncclSum, cs.comms[r], pl->stream));
      }
      NCCL_OK(nc.GroupEnd());
    }
    // a single D2H of the reduced buffer from the first device; the per-shot grad_stf rows are gathered from their owners
    plan_get_result_impl(shards[0].plan.get(), misfit, gl, gm, gd, nullptr);
    for (Shard &sh : shards) {
      fwi_b200_plan *pl = sh.plan.get();
      if (gs || pl->para.save_scratch || &sh != &shards[0]) {
        std::lock_guard<std::mutex> lk(pl->mu);
        use_device(pl->gpu);
        CUDA_OK(cudaStreamSynchronize(pl->stream));
        if (gs) {
          std::vector<float> h(sh.ids.size() * (size_t)N);
          CUDA_OK(cudaMemcpy(h.data(), pl->stf_grad.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
          for (size_t k = 0; k < sh.ids.size(); k++)
            for (int t = 0; t < N; t++) gs[(size_t)sh.pos[k] * N + t] = h[k * N + t];
        }
      }
      if (pl->para.save_scratch) write_scratch(pl);
      sh.call.unlock();
    }
  });
}

// =================================================================================================
// time-lapse surveys (baseline + monitors) in one call
// =================================================================================================
extern "C" int fwi_b200_timelapse(int nsurveys, const char *const *para_fnames, const double *const *Lambda,
                                  const double *const *Mu, const double *const *Den, const double *stf, int ngpu,
                                  const int *gpu_ids, int group_size, const int *shot_ids, double *misfit,
                                  double *const *gl, double *const *gm, double *const *gd) {
  return guarded([&] {
    if (nsurveys <= 0 || !para_fnames || !Lambda || !Mu || !Den || !stf || ngpu <= 0 || !gpu_ids || group_size <= 0 ||
        !shot_ids || !misfit)
      throw Error(FWI_B200_ERR_ARG, "timelapse: bad arguments");
    struct Job { int rc = FWI_B200_OK; std::string err; };
    std::vector<Job> jobs(nsurveys);
    const int nw = std::min(ngpu, nsurveys);
    std::vector<std::thread> workers;
    for (int w = 0; w < nw; w++)
      workers.emplace_back([&, w] {   // survey i on gpu_ids[i % ngpu] (main_two_phase_flow_inversion.jl:84-93)
        for (int i = w; i < nsurveys; i += nw) {
          jobs[i].rc = host_call(&misfit[i], gl ? gl[i] : nullptr, gm ? gm[i] : nullptr, gd ? gd[i] : nullptr, nullptr,
                                 Lambda[i], Mu[i], Den[i], stf, 1, true, gpu_ids[w], group_size, shot_ids, para_fnames[i]);
          if (jobs[i].rc != FWI_B200_OK) jobs[i].err = last_error_cstr();
        }
      });
    for (auto &t : workers) t.join();
    for (int i = 0; i < nsurveys; i++)
      if (jobs[i].rc != FWI_B200_OK) throw Error(jobs[i].rc, "survey " + std::to_string(i) + ": " + jobs[i].err);
  });
}

extern "C" int fwi_b200_set_option(const char *name, int value) {
  return guarded([&] {
    if (!name) throw Error(FWI_B200_ERR_ARG, "set_option: null name");
    const std::string k(name);
    if (k == "rev_lean") set_rev_lean(value);
    else if (k == "dyn_units") g_dyn_units.store(value != 0, std::memory_order_relaxed);
    else if (k == "acc_group") {   // shots per accumulator slot of the reverse step: 0 automatic, 1 a slot per shot, k forced
      if (value < 0) throw Error(FWI_B200_ERR_ARG, "set_option: acc_group must be >= 0");
      set_acc_group(value);
    }
    else if (k == "merged_bwd") g_merged_bwd.store(value != 0, std::memory_order_relaxed);
    else if (k == "frame_ring") {   // takes effect for plans created afterwards
      if (value != 2 && value != 5) throw Error(FWI_B200_ERR_ARG, "set_option: frame_ring must be 2 or 5");
      g_frame_ring.store(value, std::memory_order_relaxed);
    }
    else throw Error(FWI_B200_ERR_ARG, "set_option: unknown option '" + k + "'");
  });
}

extern "C" int fwi_b200_para_info(const char *para_fname, int *out) {
  return guarded([&] {
    if (!para_fname || !out) throw Error(FWI_B200_ERR_ARG, "para_info: null pointer");
    const Para p = read_para(para_fname);
    const int v[8] = {p.nz, p.nx, p.nSteps, p.nPml, p.nPad, p.if_win ? 1 : 0, p.save_scratch ? 1 : 0, 0};
    std::copy(v, v + 8, out);
  });
}

extern "C" void fwi_b200_release(void) {
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache.clear();
  }
  if (!g_comms.empty()) destroy_comms();
}
