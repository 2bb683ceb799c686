// Device-side data layout and kernel launchers of the FWI hot path (sm_100a).
//
// Layout in HBM (all float32, z fastest like the reference's a[x*nz+z]):
//   plane  = one 2-D array, column pitch P (multiple of 32 floats), XM margin columns on
//            either side of x and SLACK floats before/after, so halo reads never leave the
//            allocation.  A "plane pointer" always points at (z=0, x=0).
//   state  = [shot][slot][plane]: 36 planes per concurrent shot (slots below)
//   model  = lambda, mu, mu_bar, byc_a, byc_b planes shared by all shots
//   frames = [shot][step][field 0..4][flen]  saved boundary frames (the quads covering the 5-cell ring)
//   traces = [shot][step][nrp]  (receiver fastest -> coalesced record / inject)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace fwi {

constexpr int XM = 4;        // margin columns in x
constexpr int SLACK = 256;   // floats before / after each plane
constexpr int TILE_Z = 56;   // owner tile (z fastest): 14 float4 quads, 7 sectors of 32 B
#ifndef FWI_TILE_X
#define FWI_TILE_X 28        // owner tile columns.  28: one 512-thread CTA per SM; 12: 256-thread CTAs, two per SM (FWI_CTAS_PER_SM = 2)
#endif
#ifndef FWI_CTAS_PER_SM
#define FWI_CTAS_PER_SM 1
#endif
constexpr int TILE_X = FWI_TILE_X;   // owner tile columns: stress region TILE_X + 4 columns, velocity-input region TILE_X + 6
constexpr int NT_STEP = 16 * (TILE_X + 4);   // persistent TMA-fed step kernels: 16 quads x (TILE_X + 4) columns, one quad per thread
constexpr int CTAS_PER_SM = FWI_CTAS_PER_SM;

// state slots
enum Slot : int {
  S_FA = 0,    // forward fields, buffer A: vz vx szz sxx sxz
  S_FB = 5,    // forward fields, buffer B
  S_AA = 10,   // adjoint fields, buffer A
  S_AB = 15,   // adjoint fields, buffer B
  S_PSI_A = 20,  // memory of velocity derivatives: dvz_dz dvx_dx dvx_dz dvz_dx (buffer A)
  S_PSI_B = 24,
  S_PHI_A = 28,  // memory of stress derivatives: dszz_dz dsxz_dx dsxz_dz dsxx_dx (buffer A)
  S_PHI_B = 32,
  S_COUNT = 36
};
// gradient accumulator planes (per concurrent shot): lambda, mu (direct term), mu (spray amplitude, gathered by
// finalize_kernel), rho (spray gathered in the reverse kernel)
enum Grad : int { G_LAM = 0, G_MU = 1, G_MUS = 2, G_RHO = 3, G_COUNT = 4 };
enum Field : int { F_VZ = 0, F_VX = 1, F_SZZ = 2, F_SXX = 3, F_SXZ = 4 };
enum Psi : int { PSI_VZ_Z = 0, PSI_VX_X = 1, PSI_VX_Z = 2, PSI_VZ_X = 3 };
enum Phi : int { PHI_SZZ_Z = 0, PHI_SXZ_X = 1, PHI_SXZ_Z = 2, PHI_SXX_X = 3 };
// z / x profile rows inside the packed profile arrays
enum Prof : int { PR_RK = 0, PR_A = 1, PR_B = 2, PR_RKH = 3, PR_AH = 4, PR_BH = 5, PR_COUNT = 6 };

struct Grid {
  int nz, nx, nPml, nPad, nSteps;
  int P;               // column pitch
  long long plane;     // floats per plane (incl. margins and slack)
  long long origin;    // offset of (z=0,x=0) inside a plane allocation
  int az_hi, ax_hi;    // last active cell: nz-nPad-3, nx-3 (first active = 2)
  int zlive;           // rows >= zlive (az_hi + 1 rounded up to a quad) are never updated: every field, CPML memory and
                       // dt-scaled coefficient is identically zero there, so they are neither loaded (the TMA tensor
                       // ends at zlive and zero-fills beyond it) nor stored
  int zlo, zhi, xlo, xhi;  // inner box (reconstruction / imaging region)
  int tiles_z, tiles_x;    // tile grid of the forward / adjoint kernels covering [z_off, zlive) x [0, nx)
  int z_off;               // <= 0, multiple of 4: row of the first tile.  Chosen so that a tile boundary falls just below
                           // the top absorbing layer (+ its 2-cell fringe + the 4-row halo): the tile rows between the
                           // layers then carry no CPML code at all (C2: 2 of 4 tile rows instead of 1 of 4)
  float dt, rdz, rdx;      // 1/dz, 1/dx
  float dz, dx;            // the spacings themselves
  // boundary frames
  // boundary frames, stored at float4-quad granularity (every quad that intersects the 5-cell ring)
  int f_in;            // ring cells kept INSIDE the box on each side: 3 = the reference's 5-deep ring (Boundary.cu:17-27),
                       // 0 = only the two cells outside the box that the 4th-order stencils of the box cells reach
  int f_zq0;           // first row of the first ring quad: (zlo-2) & ~3
  int f_nqB;           // quads per column of the left / right bands (2 + f_in columns each)
  int f_tq0, f_bq0;    // z >> 2 of the first quad of the top / bottom bands
  int f_ntq, f_nbq;    // quads per column of the top / bottom bands (1 or 2)
  int f_len;           // floats per field per step
};

struct Model {
  const float *lam, *mu, *amu, *bya, *byb;  // plane pointers
  const float *ldt;  // first of the 5 consecutive dt-scaled planes (M_LDT .. M_BYBDT), zero outside the active region
};
// model planes inside the plan's model buffer
enum ModelPlane : int {
  M_LAM = 0, M_MU = 1, M_DEN = 2, M_AMU = 3, M_BYA = 4, M_BYB = 5,
  // time-step-scaled coefficients the step kernels consume (TMA boxes over planes 6..8 and 9..10)
  M_LDT = 6, M_L2MDT = 7, M_AMUDT = 8, M_BYADT = 9, M_BYBDT = 10, M_COUNT = 11
};

// TMA descriptors (cuTensorMapEncodeTiled): 3-D tensors (z, x + XM, plane), float32, no swizzle, zero OOB fill.
struct alignas(64) TmaMaps {
  CUtensorMap v;    // state planes, box (TILE_Z+16, TILE_X+6, 2): velocity pair with halo 8 / 3
  CUtensorMap s;    // state planes, box (TILE_Z+8,  TILE_X+4, 3): stress triple with halo 4 / 2
  CUtensorMap sw;   // state planes, box (TILE_Z+16, TILE_X+8, 3): stress triple with halo 8 / 4 (reverse step)
  CUtensorMap vn;   // state planes, box (TILE_Z+8,  TILE_X+4, 2): velocity pair with halo 4 / 2 (reverse / adjoint step)
  CUtensorMap s3;   // state planes, box (TILE_Z+16, TILE_X+6, 3): stress triple with halo 8 / 3 (adjoint step)
  // L2 prefetch boxes (cp.async.bulk.prefetch.tensor): operands the threads read with direct loads
  CUtensorMap o5;   // state planes, box (TILE_Z, TILE_X, 5): the five adjoint fields of the owner tile (reverse step)
  CUtensorMap r1;   // state planes, box (TILE_Z+8, TILE_X+4, 1): one CPML memory plane of the tile region
  CUtensorMap g4;   // imaging accumulators [batch][G_COUNT][plane], box (TILE_Z, TILE_X, 4)
  CUtensorMap m5;   // model planes, box (TILE_Z+8, TILE_X+4, 5): the five dt-scaled coefficient planes of the tile region
};

struct Profiles {
  const float *z;  // [PR_COUNT][P]      index by z (valid for z < nz-nPad)
  const float *x;  // [PR_COUNT][nxp]    index by x + XM
  int nxp;
};

struct ShotTables {
  const int *src_z, *src_x;     // [batch]
  const float *stf;             // [batch][nSteps]  tapered source, float
  const int *rec_ptr;           // [batch][ntiles+1] CSR over tiles
  const int *rec_loc;           // [batch][nrp]  (lz | lx << 16) inside the tile
  const int *rec_id;            // [batch][nrp]  receiver index
  int nrp;                      // padded receiver count (row pitch of traces)
};

struct FwdArgs {
  Grid g;
  Model m;
  Profiles pr;
  ShotTables st;
  float *state;        // [batch][S_COUNT][plane] (allocation base)
  float *traces;       // [batch][nSteps][nrp]
  float *frames;       // [batch][nSteps][5][f_len] or nullptr
  int batch;
  int it;              // time index of the state being advanced (it -> it+1)
  int cur;             // 0: read buffer A, write B; 1: the reverse
  int order;           // 0: items in ascending order, 1: descending (see launch_forward_step)
  TmaMaps tm;
};

struct BwdArgs {
  Grid g;
  Model m;
  Profiles pr;
  ShotTables st;
  float *state;
  const float *res;    // [batch][nSteps][nrp] tapered residual
  const float *frames;
  float *gacc;         // [ceil(batch / acc_group)][G_COUNT][plane] imaging accumulators
  float *stf_grad;     // [batch][nSteps]
  int batch;
  int it;
  int cur_f;           // forward-field buffer holding state it+1
  int cur_a;           // adjoint buffer holding the pre-update adjoint state
  int order;           // item order of this launch (0 ascending, 1 descending)
  int *unit_counter;   // reverse step with shot groups: {next unit, CTAs done} in global memory, zero between launches
                       // (nullptr: units are dealt round-robin)
  int acc_group;       // reverse step: shots that share one accumulator slot of gacc (reverse_acc_group; 1 = a slot per shot)
  int indep;           // adjoint step only: the launch before it in the stream is the reverse step of the same time
                       // index, whose output it does not touch -> no wait in the prologue (see adj_step_kernel)
  TmaMaps tm;
};

// one forward time step for `batch` shots: stress + source + velocity + record (+ frame save)
void launch_forward_step(const FwdArgs &a, bool save_frames, cudaStream_t s);
// reverse-time reconstruction it+1 -> it with frame restore + imaging condition
void launch_reverse_imaging(const BwdArgs &a, cudaStream_t s);
// shots per accumulator slot the reverse step should use for this grid and batch (the caller passes it in
// BwdArgs::acc_group and sums ceil(batch / acc_group) slots in launch_finalize); set_acc_group: 0 automatic, k forced
int reverse_acc_group(const Grid &g, int batch, bool dynamic_units);
void set_acc_group(int v);
// adjoint step: source_grad, adjoint velocity, residual injection, adjoint stress
void launch_adjoint_step(const BwdArgs &a, cudaStream_t s);
// merged backward launch: adjoint step of time index a.it + 1 (adjoint buffer cur_a -> the other), then reverse step
// a.it + 1 -> a.it with imaging (forward buffer cur_f -> the other); same accumulators as the two-launch form
void launch_backward_merged(const BwdArgs &a, cudaStream_t s);

// model preparation: double row-major MPa -> float planes (Pa), derived coefficients, max cp
// `model` = base of the M_COUNT model planes
// column_major: the caller's (nz, nx) arrays are column-major (z fastest, Julia) instead of row-major [z][x] (TF)
void launch_model_prep(const Grid &g, const double *d_lam, const double *d_mu, const double *d_den, float *model,
                       unsigned int *cpmax_bits, int column_major, cudaStream_t s);

// velocity-space front end: `in` = [cp | cs | rho | cp_ref | cs_ref | rho_ref] doubles of n_in elements each (unpadded
// (nz0, nx0) or padded grids in the caller's layout; the refs only when !is_masked) -> vel [3][nz nx] + model_in [3][nz nx]
void launch_velocity_prep(const Grid &g, int column_major, int padded, int is_masked, const double *in, long long n_in,
                          double *vel, double *model_in, cudaStream_t s);
// packed float result [gl | gm | gd] + vel -> out [g_cp | g_cs | g_rho] doubles on the padded grid, caller's layout
void launch_velocity_grad(const Grid &g, int column_major, int is_masked, const float *result, const double *vel, double *out,
                          cudaStream_t s);

// residual: taper obs & syn, res = obs - syn (t=0 -> 0), partial sums of res^2, taper res
struct ResidualArgs {
  const float *syn_tr;   // [nSteps][nrp]  raw synthetic (receiver fastest)
  const float *obs_rt;   // [nrec][nSteps] observed (time fastest, file layout)
  const float *w2;       // [nSteps] taper multipliers
  // per-trace windows (para "if_win", libCUFD.cu:257-266,304-309): seconds / trace weights, null without if_win
  const float *win_start, *win_end, *weights;   // [nrec]
  float dt;
  float *res_tr;         // [nSteps][nrp]  tapered residual for injection
  float *syn_rt;         // [nrec][nSteps] conditioned synthetic   (may be null)
  float *res_rt;         // [nrec][nSteps] tapered residual        (may be null)
  float *obs_cond_rt;    // [nrec][nSteps] conditioned observed    (may be null)
  double *partial;       // [gridDim] block partial sums
  int nrec, nrp, nSteps;
};
void launch_residual(const ResidualArgs &a, int *nblocks_out, cudaStream_t s);
void launch_sum_partials(const double *partial, int n, float *out_j, cudaStream_t s);
void launch_misfit(const float *j_shot, int n, float *misfit_half, cudaStream_t s);
// traces [nSteps][nrp] -> [nrec][nSteps]
void launch_traces_to_rt(const float *tr, float *rt, int nrec, int nrp, int nSteps, cudaStream_t s);

// result = [gl|gm|gd|misfit] row-major [z][x] float: sums the per-slot accumulators
void launch_finalize(const Grid &g, const float *gacc, int nslots, const float *mu, const float *misfit_half,
                     float *result, int column_major, cudaStream_t s);

// host: encode the TMA descriptors for a state buffer of `nplanes` planes and the model buffer
// `gacc` may be null (no gradient): its descriptor is then left untouched
void encode_tma_maps(const Grid &g, float *state, long long nplanes, float *gacc, long long gacc_planes, float *model,
                     TmaMaps *out);
size_t forward_smem_bytes();
size_t reverse_smem_bytes();
size_t adjoint_smem_bytes();
size_t merged_smem_bytes();
void configure_kernels();  // cudaFuncSetAttribute for dynamic smem
void set_rev_lean(int v);   // A/B switch of the reverse kernel build: -1 auto, 0 double-buffered, 1 LEAN

}  // namespace fwi
