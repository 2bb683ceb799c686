// Forward time step of the 2-D elastic propagator for sm_100a: ONE persistent kernel per time step advances every
// shot of the batch:  stress update + CPML  ->  source injection  ->  velocity update + CPML  ->  receiver
// recording (+ boundary-frame save when a gradient will follow).
//   replaces el_stress(isFor) / add_source / el_velocity(isFor) / recording / from_bnd x5
//   (reference: deps/CustomOps/FWI/Src/libCUFD.cu:202-240, el_stress.cu:50-88, el_velocity.cu:45-82,
//    utilities.cu:361-392,521-567)
//
// Structure (B200): one CTA of 16 warps per SM, looping over (shot, tile) work items.
//   * producer: one lane of the warp that owns the two right-hand halo columns (it has no velocity work) builds, for
//     the item two iterations ahead, a small tile descriptor in shared memory and asks the TMA unit for the item's
//     halo tiles -- velocity pair (72 x 34) and stress triple (64 x 32), two cp.async.bulk.tensor boxes, 44 KB --
//     into a 2-stage shared-memory ring; completion is counted in bytes on an mbarrier.  The slot it refills is
//     the one every warp finished reading before the block barrier of the current item.
//   * compute: every thread owns ONE float4 quad (4 consecutive z cells) of the 64 x 32 stress region for
//     both half-steps.  Stress: derivatives of the velocity tile, CPML, update, source; the new stresses go to a
//     double-buffered shared tile (one block barrier per item).  Velocity: the same thread updates the velocities
//     of the same quad from its own registers + the neighbours' stresses in shared memory.
//   * the dt-scaled coefficient planes are ZERO outside the reference's active region (fwi_kernels.cu,
//     model_derive_kernel), so inactive cells keep their value without a single predicate, and the CPML
//     recursion is applied to whole quads (its profiles are the identity outside the layers).
// HBM sees each field once per step (halo re-reads are L2 hits); all global accesses are 16-byte.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "fwi_device.cuh"
#include "fwi_host.hpp"

namespace fwi {
using namespace dev;

namespace {

#ifndef FWD_NS
#define FWD_NS 2
#endif
#ifndef FWI_L2PF
#define FWI_L2PF 1
#endif
#ifndef FWI_TMA_L2PROMO
#define FWI_TMA_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif
#ifndef FWD_SNEW_SINGLE
#define FWD_SNEW_SINGLE 1   // one new-stress tile handed over with an arrive / wait pair: 113 KB of shared memory -> 124 KB of L1
#endif
#ifndef FWI_ZIGZAG
#define FWI_ZIGZAG 1
#endif
#ifndef FWD_PF_MODEL
#define FWD_PF_MODEL 1   // coefficient tile -> L2 two items ahead: 0 off, 1 every item, 2 only the item of shot 0
#endif
constexpr int NS = FWD_NS;               // ring stages
constexpr int NTHREADS_FWD = NCOMPUTE;
constexpr int V_BYTES = 2 * VCOLS * VPITCH * 4;
constexpr int S_BYTES = 3 * SCOLS * SPITCH * 4;
constexpr int STAGE_BYTES = V_BYTES + S_BYTES;
constexpr int SNEW_BYTES = 3 * SCOLS * SPITCH * 4;
constexpr int DESC_BYTES = 64;
constexpr int NSNEW = FWD_SNEW_SINGLE ? 1 : 2;
constexpr size_t FWD_SMEM = (size_t)NS * STAGE_BYTES + NSNEW * SNEW_BYTES + (NS + 1) * DESC_BYTES + (NS + 1) * 8 + 128;
static_assert(V_BYTES % 128 == 0 && S_BYTES % 128 == 0, "TMA destination alignment");

template <bool SAVE>
__global__ void __launch_bounds__(NTHREADS_FWD, CTAS_PER_SM) fwd_step_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float *s_new_base = reinterpret_cast<float *>(base + NS * STAGE_BYTES);                     // [2][3][SCOLS][SPITCH]
  TileDesc *sdesc = reinterpret_cast<TileDesc *>(base + NS * STAGE_BYTES + NSNEW * SNEW_BYTES);   // [NS + 1]
  uint64_t *full = reinterpret_cast<uint64_t *>(base + NS * STAGE_BYTES + NSNEW * SNEW_BYTES + (NS + 1) * DESC_BYTES);   // [NS] ring + [1] "s_new free"

  const Grid &g = a.g;
  const int tid = threadIdx.x;
  const int ntiles = g.tiles_z * g.tiles_x;
  const int nitems = a.batch * ntiles;
  // Items are dealt round-robin: at any instant the CTAs work on ~148 consecutive items = a few z-adjacent tiles of
  // every shot, i.e. whole grid columns -- contiguous HBM pages.  (A contiguous chunk of items per CTA, which would keep
  // a tile's coefficients in registers across shots, was measured 15-45 % SLOWER: the accesses scatter over HBM pages.)
  const int stride = gridDim.x;
  const int fin = a.cur ? S_FB : S_FA, fout = a.cur ? S_FA : S_FB;
  const int P = g.P;
  const long long pl = g.plane;
  const int zp_hi = g.nz - g.nPml - g.nPad - 1;  // z > zp_hi is bottom PML
  const int pin_ = a.cur ? S_PSI_B : S_PSI_A;

  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < NS; s++) mbar_init(&full[s], 1);
    mbar_init(&full[NS], NCOMPUTE / 32);   // "every warp has finished the velocity half of the previous item"
    fence_barrier_init();
  }
  __syncthreads();

  // producer step: descriptor (slot `ds` of NS + 1, so that it never overwrites the one in use) + TMA requests of
  // one item into ring slot `stage`
  auto produce = [&](int item, int stage, int ds, bool first = false) {
    const int io = a.order ? nitems - 1 - item : item;             // zig-zag over launches (launch_forward_step)
    const int tile = io / a.batch, shot = io - tile * a.batch;      // shot fastest: the shots of a tile share its coefficients in L2
    const int tz = tile % g.tiles_z, tx = tile / g.tiles_z;
    const int z0 = tz * TILE_Z + g.z_off, x0 = tx * TILE_X;
    const int sz = a.st.src_z[shot], sx = a.st.src_x[shot];
    TileDesc d;
    d.soff = (long long)shot * S_COUNT * pl + (long long)x0 * P + z0;
    d.moff = x0 * P + z0;
    d.z0 = z0; d.x0 = x0; d.shot = shot; d.tile = tile; d.sz = sz; d.sx = sx;
    d.r0 = a.st.rec_ptr[shot * (ntiles + 1) + tile];
    d.r1 = a.st.rec_ptr[shot * (ntiles + 1) + tile + 1];
    int fl = 0;
    if ((z0 - 4 < g.nPml) || (z0 + TILE_Z + 3 > zp_hi) || (x0 - 2 < g.nPml) || (x0 + TILE_X + 1 > g.nx - g.nPml - 1)) fl |= TF_PML;
    if (tile_touches_frame(g, z0, x0)) fl |= TF_FRAME;
    if (sz >= z0 - 4 && sz < z0 + TILE_Z + 4 && sx >= x0 - 2 && sx < x0 + TILE_X + 2) fl |= TF_SRC;
    d.flags = fl;
    d.pad[0] = d.pad[1] = d.pad[2] = d.pad[3] = 0;
    if (item + stride < nitems) {   // tile origin of this CTA's NEXT item: its coefficient quads are fetched one item ahead,
      const int ion = a.order ? nitems - 1 - (item + stride) : item + stride;   // and 512 threads need not divide for it
      const int tn = ion / a.batch;
      d.pad[0] = (tn % g.tiles_z) * TILE_Z + g.z_off;
      d.pad[1] = (tn / g.tiles_z) * TILE_X;
    }
    sdesc[ds] = d;
    unsigned char *sb = base + stage * STAGE_BYTES;
    const int p0 = shot * S_COUNT + fin;
    if (first) pdl_wait();   // everything above reads static tables only; the wavefields belong to the previous launch
    mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);   // release: the descriptor is visible to whoever sees the phase flip
    tma_load_3d(sb, &a.tm.v, z0 - 8, x0 - 3 + XM, p0 + F_VZ, &full[stage]);
    tma_load_3d(sb + V_BYTES, &a.tm.s, z0 - 4, x0 - 2 + XM, p0 + F_SZZ, &full[stage]);
#if FWD_PF_MODEL
    if (FWD_PF_MODEL == 1 || shot == 0) tma_prefetch_3d(&a.tm.m5, z0 - 4, x0 - 2 + XM, M_LDT);
#endif
#if FWI_L2PF
    if (fl & TF_PML) {  // CPML memory of the layers this tile touches: HBM -> L2 now, direct loads two items later
      const int ps = shot * S_COUNT;
      if ((z0 - 4 < g.nPml) || (z0 + TILE_Z + 3 > zp_hi)) {
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + pin_ + PSI_VZ_Z);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + pin_ + PSI_VX_Z);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + S_PHI_A + PHI_SZZ_Z);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + S_PHI_A + PHI_SXZ_Z);
      }
      if ((x0 - 2 < g.nPml) || (x0 + TILE_X + 1 > g.nx - g.nPml - 1)) {
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + pin_ + PSI_VX_X);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + pin_ + PSI_VZ_X);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + S_PHI_A + PHI_SXZ_X);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + S_PHI_A + PHI_SXX_X);
      }
    }
#endif
  };
  if (tid == PRODUCER_TID)
    for (int s = 0; s < NS; s++)
      if (blockIdx.x + s * stride < nitems) produce(blockIdx.x + s * stride, s, s, s == 0);

  const float dt = g.dt;
  const float kz1 = C1 * g.rdz, kz2 = C2 * g.rdz, kx1 = C1 * g.rdx, kx2 = C2 * g.rdx;
  const int q = tid & 15, c = tid >> 4;
  const bool inner = q >= 1 && q <= TILE_Z / 4 && c >= 2 && c < TILE_X + 2;
  const int sj = c * SPITCH + 4 * q;                 // this thread's quad inside a stress tile
  const float *zprof = a.pr.z;
  const int gx_max = g.nx + XM - 1;                  // last allocated (margin) column
  const int pin = a.cur ? S_PSI_B : S_PSI_A, pout = a.cur ? S_PSI_A : S_PSI_B;

  // dt-scaled coefficients of the quad (5 consecutive model planes).  They are fetched ONE ITEM AHEAD, straight into
  // registers, so that their latency hides behind the velocity half-step of the previous item.  Quads of the halo
  // that fall outside the grid read the zero margins of the planes (or other finite values nobody uses).
  auto coef_ptr = [&](const TileDesc &d) {
    const int gz = d.z0 - 4 + 4 * q, gx = min(d.x0 - 2 + c, gx_max);
    return a.m.ldt + ((long long)gx * P + gz);
  };
  F4 ldt, l2mdt, amudt, byadt, bybdt;
  {
    __syncthreads();                                 // the producer's first descriptors are in shared memory
    const float *mq = coef_ptr(sdesc[0]);
    ldt = ld4(mq); l2mdt = ld4(mq + pl); amudt = ld4(mq + 2 * pl); byadt = ld4(mq + 3 * pl); bybdt = ld4(mq + 4 * pl);
  }

  pdl_wait();
  int stage = 0, phase = 0, nb = 0, ds = 0, kdone = 0;
  for (int item = blockIdx.x; item < nitems; item += stride) {
    mbar_wait(&full[stage], phase);
    const TileDesc d = sdesc[ds];
    const int gz = d.z0 - 4 + 4 * q, gx = d.x0 - 2 + c;
    const bool inb = (unsigned)gx < (unsigned)g.nx && (unsigned)gz < (unsigned)g.zlive;
    const bool owner = inner && inb;
    float *sq = a.state + g.origin + d.soff + ((long long)(c - 2) * P + 4 * q - 4);   // + slot * pl
    const bool pml = (d.flags & TF_PML) && inb;
    const bool zq = pml && (gz < g.nPml || gz + 3 > zp_hi);
    const bool xq_s = pml && (gx < g.nPml || gx > g.nx - g.nPml - 1);   // stress flavour    (el_stress.cu:61)
    const bool xq_v = pml && (gx < g.nPml || gx > g.nx - g.nPml);       // velocity flavour  (el_velocity.cu:56)
    F4 pz1, pz2, px1, px2;
    if (zq) {
      pz1 = ld4s(sq + (pin + PSI_VZ_Z) * pl);
      pz2 = ld4s(sq + (pin + PSI_VX_Z) * pl);
    }
    if (xq_s) {
      px1 = ld4s(sq + (pin + PSI_VX_X) * pl);
      px2 = ld4s(sq + (pin + PSI_VZ_X) * pl);
    }

    const unsigned char *sb = base + stage * STAGE_BYTES;
    const float *sv = reinterpret_cast<const float *>(sb);             // [2][VCOLS][VPITCH]
    const float *so = reinterpret_cast<const float *>(sb + V_BYTES);   // [3][SCOLS][SPITCH]
    float *s_new = s_new_base + (FWD_SNEW_SINGLE ? 0 : nb) * (SNEW_BYTES / 4);

    // ---- stress on 16 quads x 32 columns (el_stress.cu:50-88) ----
    const float *vzc = sv + (c + 1) * VPITCH + 4 * (q + 1);
    const float *vxc = vzc + VCOLS * VPITCH;
    float dvz_dz[4], dvx_dz[4], dvx_dx[4], dvz_dx[4];
    const F4 zB = ld4(vzc), xB = ld4(vxc);
    dz_minus4(ld4(vzc - 4), zB, ld4(vzc + 4), kz1, kz2, dvz_dz);
    dz_plus4(ld4(vxc - 4), xB, ld4(vxc + 4), kz1, kz2, dvx_dz);
    // the outermost halo columns need only one of the two x-derivatives: keep their reads inside the tile
    dx4(ld4(vxc - (c > 0 ? 2 : 1) * VPITCH), ld4(vxc - VPITCH), xB, ld4(vxc + VPITCH), kx1, kx2, dvx_dx);
    dx4(ld4(vzc - VPITCH), zB, ld4(vzc + VPITCH), ld4(vzc + (c < SCOLS - 1 ? 2 : 1) * VPITCH), kx1, kx2, dvz_dx);
    F4 szz = ld4(so + sj), sxx = ld4(so + SCOLS * SPITCH + sj), sxz = ld4(so + 2 * SCOLS * SPITCH + sj);

    if (SAVE && (d.flags & TF_FRAME) && owner) {  // from_bnd x5: state at time `it`, before the update (libCUFD.cu:206)
      const int fq = frame_quad(g, gz, gx);
      if (fq >= 0) {
        float *frm = a.frames + ((long long)d.shot * g.nSteps + a.it) * 5 * g.f_len + 4 * fq;
        st4(frm + F_SZZ * g.f_len, szz);
        st4(frm + F_SXX * g.f_len, sxx);
        st4(frm + F_SXZ * g.f_len, sxz);
        st4(frm + F_VZ * g.f_len, zB);
        st4(frm + F_VX * g.f_len, xB);
      }
    }
    if (zq) {  // z-CPML on the whole quad: a = 0, b = 1, 1/K = 1 outside the layer (el_stress.cu:57-60,74-77)
      const F4 b = ld4(zprof + PR_B * P + gz), aa = ld4(zprof + PR_A * P + gz), rk = ld4(zprof + PR_RK * P + gz);
      const F4 bh = ld4(zprof + PR_BH * P + gz), ah = ld4(zprof + PR_AH * P + gz), rkh = ld4(zprof + PR_RKH * P + gz);
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        pz1.v[kk] = fmaf(b.v[kk], pz1.v[kk], aa.v[kk] * dvz_dz[kk]);
        dvz_dz[kk] = fmaf(dvz_dz[kk], rk.v[kk], pz1.v[kk]);
        pz2.v[kk] = fmaf(bh.v[kk], pz2.v[kk], ah.v[kk] * dvx_dz[kk]);
        dvx_dz[kk] = fmaf(dvx_dz[kk], rkh.v[kk], pz2.v[kk]);
      }
      if (owner) {
        st4(sq + (pout + PSI_VZ_Z) * pl, pz1);
        st4(sq + (pout + PSI_VX_Z) * pl, pz2);
      }
    }
    if (xq_s) {  // x-CPML (el_stress.cu:61-64,78-81)
      const float *xp = a.pr.x + gx + XM;
      const int n = a.pr.nxp;
      const float b = xp[PR_B * n], aa = xp[PR_A * n], rk = xp[PR_RK * n];
      const float bh = xp[PR_BH * n], ah = xp[PR_AH * n], rkh = xp[PR_RKH * n];
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        px1.v[kk] = fmaf(b, px1.v[kk], aa * dvx_dx[kk]);
        dvx_dx[kk] = fmaf(dvx_dx[kk], rk, px1.v[kk]);
        px2.v[kk] = fmaf(bh, px2.v[kk], ah * dvz_dx[kk]);
        dvz_dx[kk] = fmaf(dvz_dx[kk], rkh, px2.v[kk]);
      }
      if (owner) {
        st4(sq + (pout + PSI_VX_X) * pl, px1);
        st4(sq + (pout + PSI_VZ_X) * pl, px2);
      }
    }
    // FWI_F64_UPDATE = R > 1: double-precision increments only for quads within R cells of the shot's source (a per-cell
    // criterion, so every tile that computes the quad agrees)
    const bool near_src = FWI_F64_UPDATE > 1 && abs(gx - d.sx) <= FWI_F64_UPDATE && gz + 3 >= d.sz - FWI_F64_UPDATE &&
                          gz <= d.sz + FWI_F64_UPDATE;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {  // el_stress.cu:66-67,83; coefficients carry dt and are 0 on inactive cells
      // the reference's (lambda + 2.0 mu) promotes the whole increment to double: ONE rounding per update
      if (FWI_F64_UPDATE == 1 || (FWI_F64_UPDATE > 1 && near_src)) {
        szz.v[kk] = (float)((double)szz.v[kk] + ((double)l2mdt.v[kk] * (double)dvz_dz[kk] + (double)ldt.v[kk] * (double)dvx_dx[kk]));
        sxx.v[kk] = (float)((double)sxx.v[kk] + ((double)ldt.v[kk] * (double)dvz_dz[kk] + (double)l2mdt.v[kk] * (double)dvx_dx[kk]));
      } else {
        szz.v[kk] = fmaf(l2mdt.v[kk], dvz_dz[kk], fmaf(ldt.v[kk], dvx_dx[kk], szz.v[kk]));
        sxx.v[kk] = fmaf(l2mdt.v[kk], dvx_dx[kk], fmaf(ldt.v[kk], dvz_dz[kk], sxx.v[kk]));
      }
      sxz.v[kk] = fmaf(amudt.v[kk], dvx_dz[kk] + dvz_dx[kk], sxz.v[kk]);
    }
    if ((d.flags & TF_SRC) && gx == d.sx && (unsigned)(d.sz - gz) < 4u) {  // add_source (utilities.cu:521-537): point stamp
      const float amp = a.st.stf[d.shot * g.nSteps + a.it];
      const float azz = SRC_SCALE * amp * dt;
      const double axx = 3.0 * (double)SRC_SCALE * (double)amp * (double)dt;
      const int ks = d.sz - gz;
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        szz.v[kk] += (kk == ks) ? azz : 0.0f;
        sxx.v[kk] = (kk == ks) ? (float)((double)sxx.v[kk] + axx) : sxx.v[kk];
      }
    }
    // single-buffered tile: the velocity half of the previous item (and its recording) has read it everywhere; the
    // arrive was posted at the end of that item, a whole stress half ago
    if (FWD_SNEW_SINGLE && kdone > 0) mbar_wait(&full[NS], (kdone - 1) & 1);
    st4(s_new + sj, szz);
    st4(s_new + SCOLS * SPITCH + sj, sxx);
    st4(s_new + 2 * SCOLS * SPITCH + sj, sxz);
    float *fo = sq + fout * pl;
    if (owner) {
      st4(fo + F_SZZ * pl, szz);
      st4(fo + F_SXX * pl, sxx);
      st4(fo + F_SXZ * pl, sxz);
    }
    // CPML memory of the velocity half-step and the NEXT item's stress coefficients: requested now, used after the barrier
    F4 fz1, fz2, fx1, fx2;
    if (zq && owner) {
      fz1 = ld4s(sq + (S_PHI_A + PHI_SZZ_Z) * pl);
      fz2 = ld4s(sq + (S_PHI_A + PHI_SXZ_Z) * pl);
    }
    if (xq_v && owner) {
      fx1 = ld4s(sq + (S_PHI_A + PHI_SXZ_X) * pl);
      fx2 = ld4s(sq + (S_PHI_A + PHI_SXX_X) * pl);
    }
    // the next item's stress coefficients: requested now, used after the barrier.  Its tile origin travels in THIS
    // item's descriptor (the next item's own descriptor is written after the previous block barrier and is therefore
    // only safe to read on the far side of this one).
    const bool more = item + stride < nitems;
    const float *mq_next = a.m.ldt;
    if (more) {
      const int gzn = d.pad[0] - 4 + 4 * q, gxn = min(d.pad[1] - 2 + c, gx_max);
      mq_next = a.m.ldt + ((long long)gxn * P + gzn);
      ldt = ld4(mq_next); l2mdt = ld4(mq_next + pl); amudt = ld4(mq_next + 2 * pl);
    }
    __syncthreads();  // s_new is complete; nobody reads ring slot `stage` any more
    if (tid == PRODUCER_TID && item + NS * stride < nitems) produce(item + NS * stride, stage, ds == 0 ? NS : ds - 1);

    // ---- recording at time index it+1 (utilities.cu:557-567) ----
    for (int r = d.r0 + tid; r < d.r1; r += NCOMPUTE) {
      const int loc = a.st.rec_loc[d.shot * a.st.nrp + r];
      const int lz = loc & 0xffff, lx = loc >> 16;
      const int j = (lx + 2) * SPITCH + lz + 4;
      a.traces[((long long)d.shot * g.nSteps + a.it + 1) * a.st.nrp + a.st.rec_id[d.shot * a.st.nrp + r]] =
          (float)((double)s_new[j] + 3.0 * (double)s_new[SCOLS * SPITCH + j]);
    }

    // ---- velocity of the same quad, owner threads (el_velocity.cu:45-82) ----
    if (owner) {
      const float *zz = s_new + sj;
      const float *xx = zz + SCOLS * SPITCH;
      const float *xz = xx + SCOLS * SPITCH;
      float dszz_dz[4], dsxz_dz[4], dsxz_dx[4], dsxx_dx[4];
      dz_plus4(ld4(zz - 4), szz, ld4(zz + 4), kz1, kz2, dszz_dz);
      dz_minus4(ld4(xz - 4), sxz, ld4(xz + 4), kz1, kz2, dsxz_dz);
      dx4(ld4(xz - 2 * SPITCH), ld4(xz - SPITCH), sxz, ld4(xz + SPITCH), kx1, kx2, dsxz_dx);
      dx4(ld4(xx - SPITCH), sxx, ld4(xx + SPITCH), ld4(xx + 2 * SPITCH), kx1, kx2, dsxx_dx);
      if (zq) {  // el_velocity.cu:52-55,67-70
        const F4 b = ld4(zprof + PR_B * P + gz), aa = ld4(zprof + PR_A * P + gz), rk = ld4(zprof + PR_RK * P + gz);
        const F4 bh = ld4(zprof + PR_BH * P + gz), ah = ld4(zprof + PR_AH * P + gz), rkh = ld4(zprof + PR_RKH * P + gz);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          fz1.v[kk] = fmaf(bh.v[kk], fz1.v[kk], ah.v[kk] * dszz_dz[kk]);
          dszz_dz[kk] = fmaf(dszz_dz[kk], rkh.v[kk], fz1.v[kk]);
          fz2.v[kk] = fmaf(b.v[kk], fz2.v[kk], aa.v[kk] * dsxz_dz[kk]);
          dsxz_dz[kk] = fmaf(dsxz_dz[kk], rk.v[kk], fz2.v[kk]);
        }
        st4(sq + (S_PHI_A + PHI_SZZ_Z) * pl, fz1);
        st4(sq + (S_PHI_A + PHI_SXZ_Z) * pl, fz2);
      }
      if (xq_v) {
        const float *xp = a.pr.x + gx + XM;
        const int n = a.pr.nxp;
        const float b = xp[PR_B * n], aa = xp[PR_A * n], rk = xp[PR_RK * n];
        const float bh = xp[PR_BH * n], ah = xp[PR_AH * n], rkh = xp[PR_RKH * n];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          fx1.v[kk] = fmaf(b, fx1.v[kk], aa * dsxz_dx[kk]);
          dsxz_dx[kk] = fmaf(dsxz_dx[kk], rk, fx1.v[kk]);
          fx2.v[kk] = fmaf(bh, fx2.v[kk], ah * dsxx_dx[kk]);
          dsxx_dx[kk] = fmaf(dsxx_dx[kk], rkh, fx2.v[kk]);
        }
        st4(sq + (S_PHI_A + PHI_SXZ_X) * pl, fx1);
        st4(sq + (S_PHI_A + PHI_SXX_X) * pl, fx2);
      }
      F4 vz = zB, vx = xB;
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {  // el_velocity.cu:60-61,75-76; buoyancies carry dt, 0 on inactive cells
        vz.v[kk] = fmaf(dszz_dz[kk] + dsxz_dx[kk], byadt.v[kk], vz.v[kk]);
        vx.v[kk] = fmaf(dsxz_dz[kk] + dsxx_dx[kk], bybdt.v[kk], vx.v[kk]);
      }
      st4(fo + F_VZ * pl, vz);
      st4(fo + F_VX * pl, vx);
    }
    if (more) {
      byadt = ld4(mq_next + 3 * pl);
      bybdt = ld4(mq_next + 4 * pl);
    }
    if (FWD_SNEW_SINGLE) {   // this warp is done with the new-stress tile
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&full[NS]);
      kdone++;
    }
    nb ^= 1;
    if (++ds == NS + 1) ds = 0;
    if (++stage == NS) { stage = 0; phase ^= 1; }
  }
}

}  // namespace

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
      throw Error(FWI_B200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this CUDA driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

void encode_one(CUtensorMap *m, const Grid &g, float *plane0, long long nplanes, int bz, int bx, int bp) {
  // plane0 points at the allocation base of the first plane; element (z, x, p) lives at plane0 + p*plane + SLACK + (x+XM)*P + z
  // the tensor ends at zlive: rows beyond it hold zeros for ever and are zero-filled by the TMA unit without a read
  const cuuint64_t dims[3] = {(cuuint64_t)g.zlive, (cuuint64_t)(g.nx + 2 * XM), (cuuint64_t)nplanes};
  const cuuint64_t strides[2] = {(cuuint64_t)g.P * sizeof(float), (cuuint64_t)g.plane * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)bz, (cuuint32_t)bx, (cuuint32_t)bp};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, plane0 + SLACK, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, FWI_TMA_L2PROMO,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(FWI_B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

}  // namespace

void encode_tma_maps(const Grid &g, float *state, long long nplanes, float *gacc, long long gacc_planes, float *model,
                     TmaMaps *out) {
  encode_one(&out->v, g, state, nplanes, VPITCH, VCOLS, 2);
  encode_one(&out->s, g, state, nplanes, SPITCH, SCOLS, 3);
  encode_one(&out->sw, g, state, nplanes, TILE_Z + 16, TILE_X + 8, 3);
  encode_one(&out->vn, g, state, nplanes, TILE_Z + 8, TILE_X + 4, 2);
  encode_one(&out->s3, g, state, nplanes, TILE_Z + 16, TILE_X + 6, 3);
  encode_one(&out->o5, g, state, nplanes, TILE_Z, TILE_X, 5);
  encode_one(&out->r1, g, state, nplanes, SPITCH, SCOLS, 1);
  if (gacc) encode_one(&out->g4, g, gacc, gacc_planes, TILE_Z, TILE_X, 4);
  encode_one(&out->m5, g, model, M_COUNT, SPITCH, SCOLS, 5);
}

size_t forward_smem_bytes() { return FWD_SMEM; }

void configure_forward_kernels() {
  cudaFuncSetAttribute(fwd_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
  cudaFuncSetAttribute(fwd_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
}

// Item order alternates from one time step to the next (FWI_ZIGZAG): the wavefields a step reads are the ones the
// previous step wrote, and what it wrote LAST is what is still in the 126 MB L2 -- so the next step starts there.
void launch_forward_step(const FwdArgs &a_in, bool save_frames, cudaStream_t s) {
  FwdArgs a = a_in;
  a.order = FWI_ZIGZAG ? (a.it & 1) : 0;
  const int nitems = a.batch * a.g.tiles_z * a.g.tiles_x;
  const int blocks = nitems < sm_count() * CTAS_PER_SM ? nitems : sm_count() * CTAS_PER_SM;
  if (save_frames)
    launch_step(fwd_step_kernel<true>, blocks, NTHREADS_FWD, FWD_SMEM, s, a);
  else
    launch_step(fwd_step_kernel<false>, blocks, NTHREADS_FWD, FWD_SMEM, s, a);
}

}  // namespace fwi
