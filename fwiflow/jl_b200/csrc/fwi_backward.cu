// Backward (reverse-time) half of the gradient for sm_100a, same persistent TMA-fed structure as fwi_forward.cu.
//
// rev_image_kernel: wavefield reconstruction it+1 -> it inside the inner box + frame restore + imaging condition
//   replaces el_velocity(isFor=false) / to_bnd x2 / add_source(isFor=false) / el_stress(isFor=false) / to_bnd x3
//   (reference: deps/CustomOps/FWI/Src/libCUFD.cu:380-403, el_velocity.cu:84-117, el_stress.cu:90-131,
//    utilities.cu:394-424,538-551)
//   The imaging condition accumulates four per-cell planes per concurrent shot: lambda, mu-direct, mu-spray amplitude S
//   and rho.  The reference's 4-point atomic "sprays" (el_stress.cu:113-124, el_velocity.cu:101-110) are linear in
//   per-cell amplitudes and are applied as deterministic gathers: the density one here, per time step (shuffle for the
//   row above, the left column evaluated in place), the mu one once per gradient by finalize_kernel.
//
// One CTA of 16 warps per SM loops over (shot, tile) items of the tile range that covers the inner box + frame ring.
// The producer lane streams, two items ahead, the stress triple of time it+1 with halo 8 / 4 (72 x 36) and the
// velocity pair with halo 4 / 2 (64 x 32) through a 2-stage TMA ring.  Every thread owns one float4 quad of the
// 64 x 32 region: it rewinds the velocities of its quad (all threads; the result goes to a double-buffered shared
// tile), then -- owner threads -- rewinds the stresses of the same quad from the neighbours' rewound velocities.
#include <atomic>
#include <cstdlib>

#include "fwi_device.cuh"
#include "fwi_host.hpp"

namespace fwi {
using namespace dev;

namespace {

#ifndef REV_NS
#define REV_NS 2
#endif
#ifndef ADJ_NS
#define ADJ_NS 2
#endif
#ifndef FWI_L2PF
#define FWI_L2PF 1
#endif
// coefficient tile -> L2 two items ahead: 0 off, 1 every item, 2 only the item of shot 0.  Off in the reverse kernel:
// its TMA queue already carries the adjoint / accumulator prefetches, and a third box per item delayed the ring loads
// (C3, 8 shots: 448 -> 628 us).
#ifndef REV_PF_MODEL
#define REV_PF_MODEL 0
#endif
#ifndef ADJ_PF_MODEL
#define ADJ_PF_MODEL 1
#endif
#ifndef FWI_ZIGZAG
#define FWI_ZIGZAG 1
#endif
#ifndef REV_LEAN_AUTO
#define REV_LEAN_AUTO 1
#endif
#ifndef ADJ_SPLIT
#define ADJ_SPLIT 1   // single-buffered phi / injection tiles are handed over with an arrive / wait pair instead of a block barrier
#endif
#ifndef ADJ_DB
#define ADJ_DB 0   // 1: double-buffer the phi / injection tiles instead of a second block barrier per item
#endif
constexpr int RW_BYTES = 3 * WCOLS * VPITCH * 4;   // stress triple, rows z0-8.., columns x0-4..
constexpr int RV_BYTES = 2 * SCOLS * SPITCH * 4;   // velocity pair, rows z0-4.., columns x0-2..
constexpr int RSTAGE_BYTES = RW_BYTES + RV_BYTES;
constexpr int SV_BYTES = 2 * SCOLS * SPITCH * 4;   // rewound velocities
constexpr int NS = REV_NS;   // ring stages of the reverse kernel
constexpr int NOWN = (TILE_Z / 4) * TILE_X;        // 392 owner quads per tile
// Two builds of the reverse kernel.  LEAN = false: double-buffered rewound-velocity tile, one landing slot per thread
// for the saved stress frame values (152.7 KB of shared memory, 92 KB of L1).  LEAN = true: ONE velocity tile handed
// over with an arrive / wait pair at the top of every item and landing slots for the owner quads only (130.5 KB ->
// the 132 KB carve-out -> 124 KB of L1).  The wait at the top costs a little where the kernel is latency-bound
// (C2: 42.7 -> 44.0 us), the larger L1 wins where it is bound by the streams of its direct loads (C3: 482 -> 461 us);
// launch_reverse_imaging picks by the size of the working set.
template <bool LEAN> constexpr int frm_bytes() { return (LEAN ? NOWN : NCOMPUTE) * 3 * 16; }
template <bool LEAN> constexpr size_t rev_smem() {
  return (size_t)NS * RSTAGE_BYTES + (LEAN ? 1 : 2) * SV_BYTES + frm_bytes<LEAN>() + (NS + 1) * sizeof(TileDesc) + (NS + 1) * 8 + 128;
}
// Shot groups (BwdArgs::acc_group > 1): the imaging accumulators of a tile stay in shared memory while the CTA takes
// the shots of the group one after the other -- one accumulator slot per GROUP in HBM, read at the group's first shot
// and written at its last, instead of 32 B per cell, shot and time index.
// Same build: the density term g_b of a quad is handed to the thread one column to the right through shared memory
// (one quad per thread; double-buffered unless LEAN, whose "velocity tile free" wait also orders these writes) instead of
// being evaluated a second time by that thread.
constexpr int RACC_ONLY = G_COUNT * NOWN * 16 + 32;
template <bool LEAN> constexpr int racc_bytes() { return RACC_ONLY + (LEAN ? 1 : 2) * NCOMPUTE * 16; }
enum : int { TF_ACC_FIRST = 64, TF_ACC_LAST = 128 };
constexpr size_t REV_SMEM = rev_smem<false>();
static_assert(RW_BYTES % 128 == 0 && RV_BYTES % 128 == 0, "TMA destination alignment");

template <bool LEAN, bool GROUPED>
__global__ void __launch_bounds__(NCOMPUTE, CTAS_PER_SM)
rev_image_kernel(const __grid_constant__ BwdArgs a, int tz_first, int tx_first, int ntz, int ntiles) {
  constexpr int NSV = LEAN ? 1 : 2;
  constexpr int NFRM = LEAN ? NOWN : NCOMPUTE;
  constexpr int FRM_BYTES = NFRM * 3 * 16;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float *s_v_base = reinterpret_cast<float *>(base + NS * RSTAGE_BYTES);                            // [2][2][SCOLS][SPITCH]
  float *s_frm = reinterpret_cast<float *>(base + NS * RSTAGE_BYTES + NSV * SV_BYTES);              // [3][NFRM] quads: szz sxx sxz
  unsigned char *rtail = base + NS * RSTAGE_BYTES + NSV * SV_BYTES + FRM_BYTES;
  TileDesc *sdesc = reinterpret_cast<TileDesc *>(rtail);   // [NS + 1]
  uint64_t *full = reinterpret_cast<uint64_t *>(rtail + (NS + 1) * sizeof(TileDesc));   // [NS] ring + [1] "velocity tile free"

  const Grid &g = a.g;
  const int tid = threadIdx.x;
  // Work is dealt round-robin in units of (tile, shot group): CTA b takes units b, b + gridDim.x, ... and, inside a
  // unit, the shots of the group one after the other.  With acc_group == 1 a unit is one (tile, shot) item, shot
  // fastest -- the order of the other step kernels.  With larger groups the CTAs walk the shots in step, each on its
  // own tile, so that at any instant they still read neighbouring tiles of the same few shots (contiguous HBM pages).
  const int G = GROUPED ? a.acc_group : 1;
  const int ngrp = (a.batch + G - 1) / G;
  const int nunits = ntiles * ngrp;
  const int stride = gridDim.x;
  constexpr int TAIL_OFF = NS * RSTAGE_BYTES + NSV * SV_BYTES + FRM_BYTES + (NS + 1) * (int)sizeof(TileDesc) + (NS + 1) * 8;
  int *pst = reinterpret_cast<int *>(base + TAIL_OFF);   // producer's position {unit, shot within the unit's group}: its lane only
  float *s_acc = reinterpret_cast<float *>(base + (TAIL_OFF + 8 + 15) / 16 * 16);   // [G_COUNT][NOWN] quads (acc_group > 1 only)
  float *s_gb_base = s_acc + G_COUNT * NOWN * 4;                                      // [1 or 2][NCOMPUTE] quads (GROUPED only)
  // GROUPED: units are claimed from a counter in global memory (a.unit_counter) instead of being dealt statically, so
  // that the units in flight are always ~gridDim.x consecutive ones however far the CTAs drift apart over the launch;
  // the loop below ends on a sentinel descriptor.  Otherwise: static round-robin, item count known up front.
  const bool dyn = GROUPED && a.unit_counter != nullptr;
  const int n_my = GROUPED ? 0 : (nunits - (int)blockIdx.x + stride - 1) / stride;   // (!GROUPED: a unit is one item)
  const int fin = a.cur_f ? S_FB : S_FA, fout = a.cur_f ? S_FA : S_FB;
  const int ain = a.cur_a ? S_AB : S_AA;
  const int P = g.P;
  const long long pl = g.plane;

  // (griddepcontrol.launch_dependents is issued AFTER the wait below: the adjoint step that follows does not wait in its
  //  prologue, so it may only be let loose once this grid knows that the previous adjoint step has completed)
  if (tid == 0) {
    for (int s = 0; s < NS; s++) mbar_init(&full[s], 1);
    mbar_init(&full[NS], NCOMPUTE / 32);   // LEAN: "every warp has finished reading the velocity tile of the previous item"
    fence_barrier_init();
  }
  __syncthreads();

  // producer: next item of this CTA (only the producer lane calls it, in sequence; p_u < nunits on entry)
  auto produce = [&](int stage, int ds, bool first = false) {
    int p_u = pst[0], p_j = pst[1];
    if (GROUPED && p_u >= nunits) {   // nothing left: sentinel (the ring slot is not armed, nobody waits on it)
      sdesc[ds].tile = -1;
      return;
    }
    const int uo = a.order ? nunits - 1 - p_u : p_u;         // reverse launches run descending, adjoint launches ascending
    const int t = uo / ngrp, grp = uo - t * ngrp;             // group fastest: the shots of a tile share its coefficients
    const int len = min(G, a.batch - grp * G);
    const int shot = grp * G + (a.order ? len - 1 - p_j : p_j);
    const int z0 = (tz_first + t % ntz) * TILE_Z, x0 = (tx_first + t / ntz) * TILE_X;
    const int sz = a.st.src_z[shot], sx = a.st.src_x[shot];
    TileDesc d;
    d.soff = (long long)shot * S_COUNT * pl + (long long)x0 * P + z0;
    d.moff = x0 * P + z0;
    d.z0 = z0; d.x0 = x0; d.shot = shot; d.tile = t; d.sz = sz; d.sx = sx; d.r0 = grp; d.r1 = 0;
    int fl = (p_j == 0 ? TF_ACC_FIRST : 0) | (p_j == len - 1 ? TF_ACC_LAST : 0);
    if (++p_j == len) { p_j = 0; p_u = dyn ? atomicAdd(a.unit_counter, 1) : p_u + stride; }
    pst[0] = p_u; pst[1] = p_j;
    // every tile whose owner cells or their +-4 halo can touch the ring (the old launch's frame_tile test)
    if (!(z0 - 4 > g.zlo - 1 + g.f_in && z0 + TILE_Z + 3 < g.zhi + 1 - g.f_in && x0 - 2 > g.xlo - 1 + g.f_in &&
          x0 + TILE_X + 1 < g.xhi + 1 - g.f_in))
      fl |= TF_FRAME;
    if (sz >= z0 && sz < z0 + TILE_Z && sx >= x0 && sx < x0 + TILE_X) fl |= TF_SRC;
    d.flags = fl;
    d.pad[0] = d.pad[1] = d.pad[2] = d.pad[3] = 0;
    sdesc[ds] = d;
    unsigned char *sb = base + stage * RSTAGE_BYTES;
    const int p0 = shot * S_COUNT + fin;
    if (first && !dyn) pdl_wait();   // everything above reads static tables only
    mbar_arrive_expect_tx(&full[stage], RSTAGE_BYTES);
    tma_load_3d(sb, &a.tm.sw, z0 - 8, x0 - 4 + XM, p0 + F_SZZ, &full[stage]);
    tma_load_3d(sb + RW_BYTES, &a.tm.vn, z0 - 4, x0 - 2 + XM, p0 + F_VZ, &full[stage]);
#if REV_PF_MODEL
    if (REV_PF_MODEL == 1 || shot == 0) tma_prefetch_3d(&a.tm.m5, z0 - 4, x0 - 2 + XM, M_LDT);
#endif
#if FWI_L2PF
    // operands of the owner quads that are fetched with direct loads: adjoint fields and imaging accumulators -> L2
    tma_prefetch_3d(&a.tm.o5, z0, x0 + XM, shot * S_COUNT + ain);
    if (fl & TF_ACC_FIRST) tma_prefetch_3d(&a.tm.g4, z0, x0 + XM, grp * G_COUNT);
#endif
  };
  if (tid == PRODUCER_TID) {
    if (dyn) {   // the counter belongs to the previous reverse launch until every earlier grid has completed
      pdl_wait();
      pst[0] = atomicAdd(a.unit_counter, 1);
    } else {
      pst[0] = blockIdx.x;
    }
    pst[1] = 0;
    for (int s = 0; s < NS; s++)
      if (GROUPED || pst[0] < nunits) produce(s, s, s == 0);
  }
  __syncthreads();   // the first descriptors are visible: per-item global loads may start before the TMA data lands
  pdl_wait();
  pdl_launch_dependents();

  const float dt = g.dt;
  const float kz1 = C1 * g.rdz, kz2 = C2 * g.rdz, kx1 = C1 * g.rdx, kx2 = C2 * g.rdx;
  const float half_rdt = 0.5f / dt;            // 0.5 byc^2 dt      = (0.5 / dt) (byc dt)^2
  const float q_rdt = 250000.0f / dt;          // 1e6 mu_bar^2 dt/4 = (250000 / dt) (mu_bar dt)^2
  const float dt6 = dt * 1e6f;
  const int q = tid & 15, c = tid >> 4;
  const bool inner = q >= 1 && q <= TILE_Z / 4 && c >= 2 && c < TILE_X + 2;
  const int sj = c * SPITCH + 4 * q;
  const int gx_max = g.nx + XM - 1;

  int stage = 0, phase = 0, nb = 0, ds = 0, kdone = 0;
  for (int item = 0; GROUPED || item < n_my; item++) {
    const TileDesc d = sdesc[ds];   // written by the producer >= 1 block barrier ago
    if (GROUPED && d.tile < 0) break;
    const int gz = d.z0 - 4 + 4 * q, gx = d.x0 - 2 + c;
    const bool inb = (unsigned)gx < (unsigned)g.nx && (unsigned)gz < (unsigned)g.zlive;
    const bool owner = inner && inb;
    const long long toff = d.soff + ((long long)(c - 2) * P + 4 * q - 4);
    float *sq = a.state + g.origin + toff;                                       // + slot * pl
    // accumulator slot of the shot's group: group * G_COUNT * pl + cell
    float *acc = a.gacc + g.origin + ((long long)d.r0 * G_COUNT - (long long)d.shot * S_COUNT) * pl + toff;
    const bool acc_first = d.flags & TF_ACC_FIRST, acc_last = d.flags & TF_ACC_LAST;
    float *my_acc = s_acc + 4 * ((q - 1) + (TILE_Z / 4) * (c - 2));   // owner quads only; planes 4 * NOWN floats apart
    const float *mq = a.m.ldt + ((long long)min(gx, gx_max) * P + gz);
    const bool colbox = gx >= g.xlo && gx <= g.xhi;
    bool bx[4];   // cell inside the inner box (reconstruction / imaging region)
#pragma unroll
    for (int kk = 0; kk < 4; kk++) bx[kk] = colbox && (unsigned)(gz + kk - g.zlo) <= (unsigned)(g.zhi - g.zlo);
    const bool in_rect = gx >= g.xlo - 2 && gx <= g.xhi + 2 && gz + 3 >= g.zlo - 2 && gz <= g.zhi + 2;
    // saved frame values of the quad (to_bnd, libCUFD.cu:388,403), copied asynchronously (LDGSTS, no registers):
    // the velocities straight into the quad's place in the rewound-velocity tile, the stresses into this thread's
    // landing zone.  State slots hold F_VZ F_VX F_SZZ F_SXX F_SXZ in that order.
    int fq = -1;
    if ((d.flags & TF_FRAME) && in_rect) fq = frame_quad(g, gz, gx);
    float *s_v = s_v_base + (LEAN ? 0 : nb) * (SV_BYTES / 4);
    float *my_frm = s_frm + 4 * (LEAN ? (q - 1) + (TILE_Z / 4) * (c - 2) : tid);   // LEAN: owner quads only (nobody else uses them)
    // LEAN: the frame copies below are the first writes into the single velocity tile
    if (LEAN && kdone > 0) mbar_wait(&full[NS], (kdone - 1) & 1);
    if (fq >= 0) {
      const float *frm = a.frames + ((long long)d.shot * g.nSteps + a.it) * 5 * g.f_len + 4 * fq;
      cp_async16(s_v + sj, frm + F_VZ * g.f_len);
      cp_async16(s_v + SCOLS * SPITCH + sj, frm + F_VX * g.f_len);
      if (!LEAN || inner) {
#pragma unroll
        for (int f = 0; f < 3; f++) cp_async16(my_frm + f * 4 * NFRM, frm + (F_SZZ + f) * g.f_len);
      }
    }

    const bool wr = owner && in_rect;
    const bool colrho = gx >= g.xlo && gx <= g.xhi + 1;   // the x+1 spray also lands in column xhi + 1
    const bool rowbox = gz + 3 >= g.zlo && gz <= g.zhi;
    if (GROUPED && acc_first && wr) {   // the group's accumulator slot -> this thread's shared slots (LDGSTS, no registers)
#pragma unroll
      for (int k = 0; k < 3; k++) cp_async16(my_acc + k * 4 * NOWN, acc + k * pl);
      if (rowbox && colrho) cp_async16(my_acc + G_RHO * 4 * NOWN, acc + G_RHO * pl);
    }

    // global operands of the velocity half, requested before waiting for the ring: buoyancies, adjoint velocities.
    // The density spray (el_velocity.cu:105-110) is applied here as a gather -- cell (z, x) receives g_a(z, x) +
    // g_b(z, x) + g_a(z-1, x) + g_b(z, x-1) -- so the quad row above the owner rows needs its adjoint vz as well (its
    // g_a comes down by shuffle), and every owner thread evaluates g_b of the column to its left itself (one more
    // adjoint quad and buoyancy quad, both L2 hits: the neighbour thread loads the same lines).
    const F4 byadt = ld4(mq + 3 * pl), bybdt = ld4(mq + 4 * pl);
    F4 vza = zero4(), vxa = zero4(), vxaL = zero4(), bybL = zero4();
    const bool above = q == 0 && c >= 2 && c < TILE_X + 2 && inb;   // halo quad right above an owner quad
    if (owner || above) vza = ld4s(sq + (ain + F_VZ) * pl);
    if (GROUPED) {   // g_b of column x-1 comes from the thread of that column: the halo column left of the tile computes it too
      const bool leftcol = c == 1 && q >= 1 && q <= TILE_Z / 4 && inb;
      if (owner || leftcol) vxa = ld4s(sq + (ain + F_VX) * pl);
    } else if (owner) {
      vxa = ld4s(sq + (ain + F_VX) * pl);
      vxaL = ld4s(sq + (ain + F_VX) * pl - P);
      bybL = ld4(mq + 4 * pl - P);
    }
    float *s_gb = s_gb_base + (LEAN ? 0 : nb) * (NCOMPUTE * 4) + 4 * tid;
    mbar_wait(&full[stage], phase);

    const unsigned char *sb = base + stage * RSTAGE_BYTES;
    const float *sw = reinterpret_cast<const float *>(sb);              // [3][WCOLS][VPITCH]: szz sxx sxz of time it+1
    const float *sv = reinterpret_cast<const float *>(sb + RW_BYTES);   // [2][SCOLS][SPITCH]: vz vx of time it+1

    // ---- v^{it} = v^{it+1} - velocity(sigma^{it+1}) on 16 quads x 32 columns; rho imaging terms (el_velocity.cu:84-110) ----
    const float *zz = sw + (c + 2) * VPITCH + 4 * (q + 1);
    const float *xx = zz + WCOLS * VPITCH;
    const float *xz = xx + WCOLS * VPITCH;
    float ea[4], eb[4];
    const F4 szzB = ld4(zz), sxxB = ld4(xx), sxzB = ld4(xz);
    {
      float d1[4], d2[4];
      dz_plus4(ld4(zz - 4), szzB, ld4(zz + 4), kz1, kz2, d1);                       // dszz_dz
      dx4(ld4(xz - 2 * VPITCH), ld4(xz - VPITCH), sxzB, ld4(xz + VPITCH), kx1, kx2, d2);   // dsxz_dx
#pragma unroll
      for (int kk = 0; kk < 4; kk++) ea[kk] = d1[kk] + d2[kk];
      dz_minus4(ld4(xz - 4), sxzB, ld4(xz + 4), kz1, kz2, d1);                      // dsxz_dz
      dx4(ld4(xx - VPITCH), sxxB, ld4(xx + VPITCH), ld4(xx + 2 * VPITCH), kx1, kx2, d2);   // dsxx_dx
#pragma unroll
      for (int kk = 0; kk < 4; kk++) eb[kk] = d1[kk] + d2[kk];
    }
    F4 vz = ld4(sv + sj), vx = ld4(sv + SCOLS * SPITCH + sj);
    F4 ga = zero4(), gb = zero4();
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      if (bx[kk]) {
        vz.v[kk] = fmaf(-ea[kk], byadt.v[kk], vz.v[kk]);
        vx.v[kk] = fmaf(-eb[kk], bybdt.v[kk], vx.v[kk]);
        // g = -v_adj (d sigma) dt * (-byc^2 / 2)     (el_velocity.cu:101-104); accumulated with the other planes below
        ga.v[kk] = (vza.v[kk] * ea[kk]) * (half_rdt * byadt.v[kk] * byadt.v[kk]);
        gb.v[kk] = (vxa.v[kk] * eb[kk]) * (half_rdt * bybdt.v[kk] * bybdt.v[kk]);
      }
    }
    // g_a of the cell right above the quad: from the thread above in the same half-warp (all 32 lanes take part)
    const float ga_up = __shfl_up_sync(0xffffffffu, ga.v[3], 1, 16);
    F4 grho = zero4();   // this step's density term of the quad, gathered
    if (GROUPED) {
      st4(s_gb, gb);   // read by the thread one column to the right after the block barrier
      if (wr && rowbox && colrho) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const bool rowin = (unsigned)(gz + kk - g.zlo) <= (unsigned)(g.zhi - g.zlo);
          const float up = kk == 0 ? ga_up : ga.v[kk - 1];
          grho.v[kk] = rowin ? (ga.v[kk] + gb.v[kk]) + up : 0.0f;   // + g_b(z, x-1) after the barrier
        }
      }
    } else if (wr && rowbox && colrho) {
      // g_b of column x-1 (zero outside the box): D-z(sxz) + D+x(sxx) one column to the left
      float ebL[4] = {0.f, 0.f, 0.f, 0.f};
      if (gx - 1 >= g.xlo && gx - 1 <= g.xhi) {
        float d1[4], d2[4];
        const float *xzL = xz - VPITCH;
        dz_minus4(ld4(xzL - 4), ld4(xzL), ld4(xzL + 4), kz1, kz2, d1);
        dx4(ld4(xx - 2 * VPITCH), ld4(xx - VPITCH), sxxB, ld4(xx + VPITCH), kx1, kx2, d2);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) ebL[kk] = d1[kk] + d2[kk];
      }
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const bool rowin = (unsigned)(gz + kk - g.zlo) <= (unsigned)(g.zhi - g.zlo);
        const float gbl = (vxaL.v[kk] * ebL[kk]) * (half_rdt * bybL.v[kk] * bybL.v[kk]);
        const float up = kk == 0 ? ga_up : ga.v[kk - 1];
        // rows outside the box receive nothing; g_a / g_b are zero outside it, and the z+1 spray stops at zhi (el_velocity.cu:107)
        grho.v[kk] = rowin ? (ga.v[kk] + gb.v[kk]) + (up + gbl) : 0.0f;
      }
    }
    if (fq >= 0) {  // exact values of time `it` on the ring: already in the shared tile
      cp_async_wait_all();
      vz = ld4(s_v + sj);
      vx = ld4(s_v + SCOLS * SPITCH + sj);
    } else {
      st4(s_v + sj, vz);
      st4(s_v + SCOLS * SPITCH + sj, vx);
    }
    float *fo = sq + fout * pl;
    if (wr) {
      st4(fo + F_VZ * pl, vz);
      st4(fo + F_VX * pl, vx);
    }
    // global operands of the stress half, requested before the barrier
    F4 ldt, l2mdt, amudt, za, xa, xza, gl, gm, gs, gd;
    if (wr) {
      ldt = ld4(mq); l2mdt = ld4(mq + pl); amudt = ld4(mq + 2 * pl);
      za = ld4s(sq + (ain + F_SZZ) * pl); xa = ld4s(sq + (ain + F_SXX) * pl); xza = ld4s(sq + (ain + F_SXZ) * pl);
      if (!GROUPED) {
        gl = ld4s(acc + G_LAM * pl); gm = ld4s(acc + G_MU * pl); gs = ld4s(acc + G_MUS * pl);
        if (rowbox && colrho) gd = ld4s(acc + G_RHO * pl);
      }
    }
    if (GROUPED && acc_first) cp_async_wait_all();   // (thread-private slots: no barrier needed for them)
    __syncthreads();  // s_v is complete; nobody reads ring slot `stage` any more
    if (tid == PRODUCER_TID && (GROUPED || pst[0] < nunits)) produce(stage, ds == 0 ? NS : ds - 1);

    // ---- sigma^{it} = sigma^{it+1} - source - stress(v^{it}) on the owner quads; lambda / mu imaging (el_stress.cu:90-124) ----
    if (wr) {
      F4 szz = szzB, sxx = sxxB, sxz = sxzB;
      if (colbox && gz + 3 >= g.zlo && gz <= g.zhi) {
        const float *pz = s_v + sj;
        const float *px = pz + SCOLS * SPITCH;
        float dvz_dz[4], dvx_dz[4], dvx_dx[4], dvz_dx[4];
        dz_minus4(ld4(pz - 4), vz, ld4(pz + 4), kz1, kz2, dvz_dz);
        dz_plus4(ld4(px - 4), vx, ld4(px + 4), kz1, kz2, dvx_dz);
        dx4(ld4(px - 2 * SPITCH), ld4(px - SPITCH), vx, ld4(px + SPITCH), kx1, kx2, dvx_dx);
        dx4(ld4(pz - SPITCH), vz, ld4(pz + SPITCH), ld4(pz + 2 * SPITCH), kx1, kx2, dvz_dx);
        if ((d.flags & TF_SRC) && gx == d.sx && (unsigned)(d.sz - gz) < 4u) {  // add_source(isFor=false): utilities.cu:538-551
          const float amp = a.st.stf[d.shot * g.nSteps + a.it];
          const float azz = SRC_SCALE * amp * dt;
          const double axx = 3.0 * (double)SRC_SCALE * (double)amp * (double)dt;
          const int ks = d.sz - gz;
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            szz.v[kk] -= (kk == ks) ? azz : 0.0f;
            sxx.v[kk] = (kk == ks) ? (float)((double)sxx.v[kk] - axx) : sxx.v[kk];
          }
        }
        if (GROUPED) {   // running sums of the group's earlier shots (or the slot's contents, copied at the top of the item)
          gl = ld4(my_acc + G_LAM * 4 * NOWN); gm = ld4(my_acc + G_MU * 4 * NOWN); gs = ld4(my_acc + G_MUS * 4 * NOWN);
        }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          if (bx[kk]) {
#if FWI_F64_UPDATE == 1
            szz.v[kk] = (float)((double)szz.v[kk] - ((double)l2mdt.v[kk] * (double)dvz_dz[kk] + (double)ldt.v[kk] * (double)dvx_dx[kk]));
            sxx.v[kk] = (float)((double)sxx.v[kk] - ((double)ldt.v[kk] * (double)dvz_dz[kk] + (double)l2mdt.v[kk] * (double)dvx_dx[kk]));
#else
            szz.v[kk] = fmaf(-l2mdt.v[kk], dvz_dz[kk], fmaf(-ldt.v[kk], dvx_dx[kk], szz.v[kk]));
            sxx.v[kk] = fmaf(-l2mdt.v[kk], dvx_dx[kk], fmaf(-ldt.v[kk], dvz_dz[kk], sxx.v[kk]));
#endif
            const float e = dvx_dz[kk] + dvz_dx[kk];
            sxz.v[kk] = fmaf(-amudt.v[kk], e, sxz.v[kk]);
            // el_stress.cu:109-116
            gl.v[kk] += -(za.v[kk] + xa.v[kk]) * (dvz_dz[kk] + dvx_dx[kk]) * dt6;
            gm.v[kk] += (-2.0f * za.v[kk] * dvz_dz[kk] - 2.0f * xa.v[kk] * dvx_dx[kk]) * dt6;
            //  s = -sxz_adj (exz + ezx) dt mu_bar / sum(1/mu) 1e6, mu_bar / sum(1/mu) == mu_bar^2 / 4; zero where mu_bar == 0
            gs.v[kk] += -xza.v[kk] * e * (q_rdt * amudt.v[kk] * amudt.v[kk]);
          }
        }
        if (!GROUPED || acc_last) {
          st4(acc + G_LAM * pl, gl);
          st4(acc + G_MU * pl, gm);
          st4(acc + G_MUS * pl, gs);
        } else {
          st4(my_acc + G_LAM * 4 * NOWN, gl);
          st4(my_acc + G_MU * 4 * NOWN, gm);
          st4(my_acc + G_MUS * 4 * NOWN, gs);
        }
      }
      if (rowbox && colrho) {
        if (GROUPED) {
          gd = ld4(my_acc + G_RHO * 4 * NOWN);
          const F4 gbl = ld4(s_gb - 64);   // g_b of the quad one column to the left (zero outside the box)
#pragma unroll
          for (int kk = 0; kk < 4; kk++) grho.v[kk] += gbl.v[kk];
        }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) gd.v[kk] += grho.v[kk];
        if (!GROUPED || acc_last) st4(acc + G_RHO * pl, gd);
        else st4(my_acc + G_RHO * 4 * NOWN, gd);
      }
      if (fq >= 0) {  // to_bnd(sigma) (libCUFD.cu:403)
        szz = ld4(my_frm);
        sxx = ld4(my_frm + 4 * NFRM);
        sxz = ld4(my_frm + 8 * NFRM);
      }
      st4(fo + F_SZZ * pl, szz);
      st4(fo + F_SXX * pl, sxx);
      st4(fo + F_SXZ * pl, sxz);
    }
    if (LEAN) {   // this warp is done with the velocity tile
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&full[NS]);
      kdone++;
    }
    nb ^= 1;
    if (++ds == NS + 1) ds = 0;
    if (++stage == NS) { stage = 0; phase ^= 1; }
  }
  if (dyn && tid == PRODUCER_TID) {   // the last CTA to get here rewinds the counter for the next reverse launch
    __threadfence();
    if (atomicInc(reinterpret_cast<unsigned int *>(a.unit_counter) + 1, gridDim.x - 1) == gridDim.x - 1) a.unit_counter[0] = 0;
  }
}


// =================================================================================================
// adj_step_kernel: source_grad + adjoint velocity + residual injection + adjoint stress
//   replaces source_grad / el_velocity_adj / res_injection / el_stress_adj
//   (reference: libCUFD.cu:376,405-427, el_velocity_adj.cu:22-108, el_stress_adj.cu:22-104, utilities.cu:569-593)
// Same skeleton: the producer lane streams the adjoint stress triple with halo 8 / 3 (72 x 34) and the adjoint
// velocity pair (64 x 32) through the TMA ring.  Every thread updates the adjoint velocities of its quad (whole
// 64 x 32 region), publishes them -- and, in tiles that touch the absorbing layers, the new phi memory variables of
// its quad -- in shared memory; after the block barrier the owner threads update the adjoint stresses of the same
// quad.  Residuals are injected through a small shared table (receivers add into it before the barrier, the owner
// of the cell picks the sum up and clears it).
// =================================================================================================
constexpr int ANS = ADJ_NS;   // ring stages of the adjoint kernel
constexpr int AS_BYTES = 3 * VCOLS * VPITCH * 4;                 // adjoint stresses, rows z0-8.., columns x0-3..
constexpr int AS_PAD = (AS_BYTES + 127) / 128 * 128;
constexpr int AV_BYTES = 2 * SCOLS * SPITCH * 4;                 // adjoint velocities, rows z0-4.., columns x0-2..
constexpr int ASTAGE_BYTES = AS_PAD + AV_BYTES;
constexpr int APHI_BYTES = 4 * SCOLS * SPITCH * 4;               // new phi of the region (tiles touching the CPML)
constexpr int AINJ_BYTES = SCOLS * SPITCH * 4;                   // residual injection table
constexpr int ANB = ADJ_DB ? 2 : 1;                               // buffers of the phi / injection tiles
#ifndef ADJ_SV1
#define ADJ_SV1 1   // 1: the new-adjoint-velocity tile is single-buffered too (ordered by the same arrive / wait pair as the phi
                    // and injection tiles: its first write of an item comes after the wait, its last read before the arrive);
                    // 16 KB more L1: adj C3 292.5 -> 289.0 us (8 shots), 933.8 -> 923.6 us (25), C2 57.9 -> 57.7 us
#endif
constexpr int ANV = (ADJ_SV1 && ADJ_SPLIT && !ADJ_DB) ? 1 : 2;    // buffers of the adjoint-velocity tile
constexpr size_t ADJ_SMEM =
    (size_t)ANS * ASTAGE_BYTES + ANV * AV_BYTES + ANB * (APHI_BYTES + AINJ_BYTES) + (ANS + 1) * sizeof(TileDesc) + (ANS + 1) * 8 + 128;
static_assert(AV_BYTES % 128 == 0, "TMA destination alignment");

__global__ void __launch_bounds__(NCOMPUTE, CTAS_PER_SM) adj_step_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float *s_v_base = reinterpret_cast<float *>(base + ANS * ASTAGE_BYTES);                         // [2][2][SCOLS][SPITCH]
  float *s_phi_base = reinterpret_cast<float *>(base + ANS * ASTAGE_BYTES + ANV * AV_BYTES);                      // [ANB][4][SCOLS][SPITCH]
  float *s_inj_base = reinterpret_cast<float *>(base + ANS * ASTAGE_BYTES + ANV * AV_BYTES + ANB * APHI_BYTES);   // [ANB][SCOLS][SPITCH]
  unsigned char *tail = base + ANS * ASTAGE_BYTES + ANV * AV_BYTES + ANB * (APHI_BYTES + AINJ_BYTES);
  TileDesc *sdesc = reinterpret_cast<TileDesc *>(tail);                                          // [ANS + 1]
  uint64_t *full = reinterpret_cast<uint64_t *>(tail + (ANS + 1) * sizeof(TileDesc));

  const Grid &g = a.g;
  const int tid = threadIdx.x;
  const int ntiles = g.tiles_z * g.tiles_x;
  const int nitems = a.batch * ntiles;
  const int stride = gridDim.x;   // round-robin item order (see fwd_step_kernel)
  const int ain = a.cur_a ? S_AB : S_AA, aout = a.cur_a ? S_AA : S_AB;
  const int psi_i = a.cur_a ? S_PSI_B : S_PSI_A, psi_o = a.cur_a ? S_PSI_A : S_PSI_B;
  const int phi_i = a.cur_a ? S_PHI_B : S_PHI_A, phi_o = a.cur_a ? S_PHI_A : S_PHI_B;
  const int P = g.P;
  const long long pl = g.plane;
  const int nxp = a.pr.nxp;
  const int zp_hi = g.nz - g.nPml - g.nPad - 1;
  // psi arrays only matter within 2 cells of the layers
  const int zq_lo = g.nPml + 2, zq_hi = g.nz - g.nPad - g.nPml - 3;
  const int xq_lo = g.nPml + 2, xq_hi = g.nx - g.nPml - 3;

  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < ANS; s++) mbar_init(&full[s], 1);
    mbar_init(&full[ANS], NCOMPUTE / 32);   // "done with the phi / injection tiles of the previous item": one arrival per warp
    fence_barrier_init();
  }
  for (int i = tid; i < ANB * AINJ_BYTES / 16; i += NCOMPUTE) reinterpret_cast<float4 *>(s_inj_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  auto produce = [&](int item, int stage, int ds, bool first = false) {
    const int io = a.order ? nitems - 1 - item : item;
    const int tile = io / a.batch, shot = io - tile * a.batch;   // shot fastest
    const int z0 = (tile % g.tiles_z) * TILE_Z + g.z_off, x0 = (tile / g.tiles_z) * TILE_X;
    const int sz = a.st.src_z[shot], sx = a.st.src_x[shot];
    TileDesc d;
    d.soff = (long long)shot * S_COUNT * pl + (long long)x0 * P + z0;
    d.moff = x0 * P + z0;
    d.z0 = z0; d.x0 = x0; d.shot = shot; d.tile = tile; d.sz = sz; d.sx = sx;
    d.r0 = a.st.rec_ptr[shot * (ntiles + 1) + tile];
    d.r1 = a.st.rec_ptr[shot * (ntiles + 1) + tile + 1];
    int fl = 0;
    if ((z0 - 4 < zq_lo) || (z0 + TILE_Z + 3 > zq_hi) || (x0 - 2 < xq_lo) || (x0 + TILE_X + 1 > xq_hi)) fl |= TF_PML;
    if (sz >= z0 && sz < z0 + TILE_Z && sx >= x0 && sx < x0 + TILE_X) fl |= TF_SRC;
    d.flags = fl;
    d.pad[0] = d.pad[1] = d.pad[2] = d.pad[3] = 0;
    sdesc[ds] = d;
    unsigned char *sb = base + stage * ASTAGE_BYTES;
    const int p0 = shot * S_COUNT + ain;
    if (first && !a.indep) pdl_wait();   // everything above reads static tables only
    mbar_arrive_expect_tx(&full[stage], AS_BYTES + AV_BYTES);
    tma_load_3d(sb, &a.tm.s3, z0 - 8, x0 - 3 + XM, p0 + F_SZZ, &full[stage]);
    tma_load_3d(sb + AS_PAD, &a.tm.vn, z0 - 4, x0 - 2 + XM, p0 + F_VZ, &full[stage]);
#if ADJ_PF_MODEL
    if (ADJ_PF_MODEL == 1 || shot == 0) tma_prefetch_3d(&a.tm.m5, z0 - 4, x0 - 2 + XM, M_LDT);
#endif
#if FWI_L2PF
    if (fl & TF_PML) {  // CPML memory of the layers this tile touches -> L2
      const int ps = shot * S_COUNT;
      if ((z0 - 4 < zq_lo) || (z0 + TILE_Z + 3 > zq_hi)) {
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + psi_i + PSI_VX_Z);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + psi_i + PSI_VZ_Z);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + phi_i + PHI_SXZ_Z);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + phi_i + PHI_SZZ_Z);
      }
      if ((x0 - 2 < xq_lo) || (x0 + TILE_X + 1 > xq_hi)) {
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + psi_i + PSI_VX_X);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + psi_i + PSI_VZ_X);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + phi_i + PHI_SXX_X);
        tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + phi_i + PHI_SXZ_X);
      }
    }
#endif
  };
  if (tid == PRODUCER_TID)
    for (int s = 0; s < ANS; s++)
      if (blockIdx.x + s * stride < nitems) produce(blockIdx.x + s * stride, s, s, s == 0);
  __syncthreads();   // the first descriptors are visible
  // Inside the backward loop the previous launch is rev_image of the same time index: it reads what this kernel reads
  // and writes nothing this kernel touches, and it only passed its own griddepcontrol.wait after the adjoint step
  // before it had completed.  So this launch starts working while the reverse step's last (partial) round of items
  // drains, and waits at its END instead -- that keeps "this grid complete => everything before it complete" for
  // the launch that follows.
  if (!a.indep) pdl_wait();

  const float dt = g.dt;
  // adjoint-kernel spelling of the differences: (-c1 (..) + c2 (..)) / h  (el_stress_adj.cu:54-61)
  const float kz1 = -C1 * g.rdz, kz2 = -C2 * g.rdz, kx1 = -C1 * g.rdx, kx2 = -C2 * g.rdx;
  const int q = tid & 15, c = tid >> 4;
  const bool inner = q >= 1 && q <= TILE_Z / 4 && c >= 2 && c < TILE_X + 2;
  const int sj = c * SPITCH + 4 * q;
  const int gx_max = g.nx + XM - 1;
  const int cm2 = (c > 0 ? 2 : 1) * VPITCH, cp2 = (c < SCOLS - 1 ? 2 : 1) * VPITCH;   // keep halo-column reads in the tile
  const float *zprof = a.pr.z;

  int stage = 0, phase = 0, nb = 0, ds = 0, kdone = 0;
  for (int item = blockIdx.x; item < nitems; item += stride) {
    const TileDesc d = sdesc[ds];   // written by the producer >= 1 block barrier ago
    const int gz = d.z0 - 4 + 4 * q, gx = d.x0 - 2 + c;
    const bool inb = (unsigned)gx < (unsigned)g.nx && (unsigned)gz < (unsigned)g.zlive;
    const bool owner = inner && inb;
    float *sq = a.state + g.origin + d.soff + ((long long)(c - 2) * P + 4 * q - 4);   // + slot * pl
    const float *mq = a.m.ldt + ((long long)min(gx, gx_max) * P + gz);
    const bool pml_tile = d.flags & TF_PML;
    // quad with at least one active cell (2 <= z <= nz-nPad-3, 2 <= x <= nx-3): the only ones that touch CPML memory
    const bool actq = gx >= 2 && gx <= g.ax_hi && gz + 3 >= 2 && gz <= g.az_hi;
    const F4 ldt = ld4(mq), l2mdt = ld4(mq + pl), amudt = ld4(mq + 2 * pl);
    const F4 byadt = ld4(mq + 3 * pl), bybdt = ld4(mq + 4 * pl);
    // double increments next to the source in the ADJOINT kernels too (FWI_F64_ADJ): off -- measured to make no
    // difference to grad_stf, which is decided by the forward arithmetic (profiles/r2_parity.md)
    const bool near_src = FWI_F64_ADJ && FWI_F64_UPDATE > 1 && abs(gx - d.sx) <= FWI_F64_UPDATE &&
                          gz + 3 >= d.sz - FWI_F64_UPDATE && gz <= d.sz + FWI_F64_UPDATE;
    mbar_wait(&full[stage], phase);

    const unsigned char *sb = base + stage * ASTAGE_BYTES;
    const float *sa = reinterpret_cast<const float *>(sb);              // [3][VCOLS][VPITCH]: adjoint szz sxx sxz
    const float *sva = reinterpret_cast<const float *>(sb + AS_PAD);    // [2][SCOLS][SPITCH]: adjoint vz vx
    float *s_v = s_v_base + (ANV == 1 ? 0 : nb) * (AV_BYTES / 4);
    float *s_phi = s_phi_base + (ADJ_DB ? nb : 0) * (APHI_BYTES / 4);
    float *s_inj = s_inj_base + (ADJ_DB ? nb : 0) * (AINJ_BYTES / 4);

    // ---- adjoint velocity on 16 quads x 32 columns (el_velocity_adj.cu:56-100) ----
    const float *zz = sa + (c + 1) * VPITCH + 4 * (q + 1);
    const float *xx = zz + VCOLS * VPITCH;
    const float *xz = xx + VCOLS * VPITCH;
    const F4 zzB = ld4(zz), xxB = ld4(xx), xzB = ld4(xz);
    float dszz_dx[4], dsxx_dx[4], dsxz_dz[4], dszz_dz[4], dsxx_dz[4], dsxz_dx[4];
    dx4(ld4(zz - VPITCH), zzB, ld4(zz + VPITCH), ld4(zz + cp2), kx1, kx2, dszz_dx);   // ad_plus_x
    dx4(ld4(xx - VPITCH), xxB, ld4(xx + VPITCH), ld4(xx + cp2), kx1, kx2, dsxx_dx);
    dx4(ld4(xz - cm2), ld4(xz - VPITCH), xzB, ld4(xz + VPITCH), kx1, kx2, dsxz_dx);   // ad_minus_x
    dz_plus4(ld4(zz - 4), zzB, ld4(zz + 4), kz1, kz2, dszz_dz);
    dz_plus4(ld4(xx - 4), xxB, ld4(xx + 4), kz1, kz2, dsxx_dz);
    dz_minus4(ld4(xz - 4), xzB, ld4(xz + 4), kz1, kz2, dsxz_dz);
    F4 vz = ld4(sva + sj), vx = ld4(sva + SCOLS * SPITCH + sj);

    // source_grad (utilities.cu:582-593): adjoint stress at the source BEFORE this step's injection and update
    if ((d.flags & TF_SRC) && owner && gx == d.sx && (unsigned)(d.sz - gz) < 4u) {
      const int ks = d.sz - gz;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; kk++)
        if (kk == ks) { s1 = zzB.v[kk]; s2 = xxB.v[kk]; }
      a.stf_grad[d.shot * g.nSteps + a.it] = (float)(-((double)s1 + 3.0 * (double)s2) * (double)dt);
    }
    // the phi / injection tiles are single-buffered: every warp has finished the stress half of the previous item
    // (arrive at the end of an item, wait here: the skew between warps is absorbed by the work above)
    if (ADJ_SPLIT && !ADJ_DB && kdone > 0) mbar_wait(&full[ANS], (kdone - 1) & 1);
    // residual injection at time index `it` (utilities.cu:569-580): receivers of this tile add into the table
    for (int r = d.r0 + tid; r < d.r1; r += NCOMPUTE) {
      const int loc = a.st.rec_loc[d.shot * a.st.nrp + r];
      const int lz = loc & 0xffff, lx = loc >> 16;
      atomicAdd(&s_inj[(lx + 2) * SPITCH + lz + 4],
                a.res[((long long)d.shot * g.nSteps + a.it) * a.st.nrp + a.st.rec_id[d.shot * a.st.nrp + r]]);
    }

    float rKx = 1.0f, rKxh = 1.0f, ax = 0.0f, axh = 0.0f;
    F4 rKz{{1.f, 1.f, 1.f, 1.f}}, rKzh{{1.f, 1.f, 1.f, 1.f}}, az = zero4(), azh = zero4();
    bool zq_pml = false, xp = false;
    if (!pml_tile) {
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {  // coefficients carry dt and are 0 on inactive cells
        // el_velocity_adj.cu:69-71,90-92: the (lambda + 2.0 mu) term promotes the sum to double
        if (FWI_F64_UPDATE == 1 || (FWI_F64_ADJ && FWI_F64_UPDATE > 1 && near_src)) {
          vx.v[kk] = (float)((double)vx.v[kk] + ((double)ldt.v[kk] * (double)dszz_dx[kk] + (double)l2mdt.v[kk] * (double)dsxx_dx[kk] +
                                                 (double)amudt.v[kk] * (double)dsxz_dz[kk]));
          vz.v[kk] = (float)((double)vz.v[kk] + ((double)l2mdt.v[kk] * (double)dszz_dz[kk] + (double)ldt.v[kk] * (double)dsxx_dz[kk] +
                                                 (double)amudt.v[kk] * (double)dsxz_dx[kk]));
        } else {
          vx.v[kk] += fmaf(ldt.v[kk], dszz_dx[kk], fmaf(l2mdt.v[kk], dsxx_dx[kk], amudt.v[kk] * dsxz_dz[kk]));
          vz.v[kk] += fmaf(l2mdt.v[kk], dszz_dz[kk], fmaf(ldt.v[kk], dsxx_dz[kk], amudt.v[kk] * dsxz_dx[kk]));
        }
      }
    } else {
      F4 f_szz_z = zero4(), f_sxz_x = zero4(), f_sxz_z = zero4(), f_sxx_x = zero4();   // new phi of the quad
      if (actq) {
        float tpx1[4] = {0, 0, 0, 0}, tpx2[4] = {0, 0, 0, 0}, tpz1[4] = {0, 0, 0, 0}, tpz2[4] = {0, 0, 0, 0};
        const float *xpf = a.pr.x + gx + XM;
        rKx = xpf[PR_RK * nxp];
        rKxh = xpf[PR_RKH * nxp];
        ax = xpf[PR_A * nxp];
        axh = xpf[PR_AH * nxp];
        xp = gx < g.nPml || gx > g.nx - g.nPml - 1;
        zq_pml = gz < g.nPml || gz + 3 > zp_hi;
        if (ax != 0.0f) {  // a_x * D+x(psi_xx)
          const float *p = sq + (psi_i + PSI_VX_X) * pl;
          float dd[4];
          dx4(ld4(p - P), ld4(p), ld4(p + P), ld4(p + 2 * P), kx1, kx2, dd);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) tpx1[kk] = ax * dd[kk];
        }
        if (axh != 0.0f) {  // a_x_half * D-x(psi_zx)
          const float *p = sq + (psi_i + PSI_VZ_X) * pl;
          float dd[4];
          dx4(ld4(p - 2 * P), ld4(p - P), ld4(p), ld4(p + P), kx1, kx2, dd);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) tpz2[kk] = axh * dd[kk];
        }
        if (zq_pml) {
          rKz = ld4(zprof + PR_RK * P + gz);
          rKzh = ld4(zprof + PR_RKH * P + gz);
          az = ld4(zprof + PR_A * P + gz);
          azh = ld4(zprof + PR_AH * P + gz);
          const float *p1 = sq + (psi_i + PSI_VX_Z) * pl;  // a_z_half * D-z(psi_xz)
          const float *p2 = sq + (psi_i + PSI_VZ_Z) * pl;  // a_z * D+z(psi_zz)
          float dd[4];
          dz_minus4(ld4(p1 - 4), ld4(p1), ld4(p1 + 4), kz1, kz2, dd);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) tpx2[kk] = azh.v[kk] * dd[kk];
          dz_plus4(ld4(p2 - 4), ld4(p2), ld4(p2 + 4), kz1, kz2, dd);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) tpz1[kk] = az.v[kk] * dd[kk];
        }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int z = gz + kk;
          if (z >= 2 && z <= g.az_hi) {
            if (FWI_F64_UPDATE == 1 || (FWI_F64_ADJ && FWI_F64_UPDATE > 1 && near_src)) {
            vx.v[kk] = (float)((double)vx.v[kk] + ((double)tpx1[kk] + (double)(ldt.v[kk] * dszz_dx[kk] * rKx) +
                                                   (double)l2mdt.v[kk] * (double)(dsxx_dx[kk] * rKx) + (double)tpx2[kk] +
                                                   (double)(amudt.v[kk] * rKzh.v[kk] * dsxz_dz[kk])));
            vz.v[kk] = (float)((double)vz.v[kk] + ((double)tpz1[kk] + (double)l2mdt.v[kk] * (double)(dszz_dz[kk] * rKz.v[kk]) +
                                                   (double)(ldt.v[kk] * dsxx_dz[kk] * rKz.v[kk]) + (double)tpz2[kk] +
                                                   (double)(amudt.v[kk] * rKxh * dsxz_dx[kk])));
            } else {
            vx.v[kk] += tpx1[kk] + ldt.v[kk] * dszz_dx[kk] * rKx + l2mdt.v[kk] * dsxx_dx[kk] * rKx + tpx2[kk] +
                        amudt.v[kk] * rKzh.v[kk] * dsxz_dz[kk];
            vz.v[kk] += tpz1[kk] + l2mdt.v[kk] * dszz_dz[kk] * rKz.v[kk] + ldt.v[kk] * dsxx_dz[kk] * rKz.v[kk] + tpz2[kk] +
                        amudt.v[kk] * rKxh * dsxz_dx[kk];
            }
          }
        }
        // phi memory of the quad, CPML cells only (el_velocity_adj.cu:74-79,95-100); buoyancies are 0 on inactive cells
        if (xp) {
          const float bx = xpf[PR_B * nxp], bxh = xpf[PR_BH * nxp];
          f_sxx_x = ld4s(sq + (phi_i + PHI_SXX_X) * pl);
          f_sxz_x = ld4s(sq + (phi_i + PHI_SXZ_X) * pl);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            f_sxx_x.v[kk] = fmaf(bxh, f_sxx_x.v[kk], bybdt.v[kk] * vx.v[kk]);
            f_sxz_x.v[kk] = fmaf(bx, f_sxz_x.v[kk], byadt.v[kk] * vz.v[kk]);
          }
          if (owner) {
            st4(sq + (phi_o + PHI_SXX_X) * pl, f_sxx_x);
            st4(sq + (phi_o + PHI_SXZ_X) * pl, f_sxz_x);
          }
        }
        if (zq_pml) {
          const F4 bz = ld4(zprof + PR_B * P + gz), bzh = ld4(zprof + PR_BH * P + gz);
          f_sxz_z = ld4s(sq + (phi_i + PHI_SXZ_Z) * pl);
          f_szz_z = ld4s(sq + (phi_i + PHI_SZZ_Z) * pl);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const int z = gz + kk;
            if (z < g.nPml || z > zp_hi) {
              f_sxz_z.v[kk] = fmaf(bz.v[kk], f_sxz_z.v[kk], bybdt.v[kk] * vx.v[kk]);
              f_szz_z.v[kk] = fmaf(bzh.v[kk], f_szz_z.v[kk], byadt.v[kk] * vz.v[kk]);
            }
          }
          if (owner) {
            st4(sq + (phi_o + PHI_SXZ_Z) * pl, f_sxz_z);
            st4(sq + (phi_o + PHI_SZZ_Z) * pl, f_szz_z);
          }
        }
      }
      st4(s_phi + PHI_SZZ_Z * SCOLS * SPITCH + sj, f_szz_z);
      st4(s_phi + PHI_SXZ_X * SCOLS * SPITCH + sj, f_sxz_x);
      st4(s_phi + PHI_SXZ_Z * SCOLS * SPITCH + sj, f_sxz_z);
      st4(s_phi + PHI_SXX_X * SCOLS * SPITCH + sj, f_sxx_x);
    }
    st4(s_v + sj, vz);
    st4(s_v + SCOLS * SPITCH + sj, vx);
    float *ao = sq + aout * pl;
    if (owner) {
      st4(ao + F_VZ * pl, vz);
      st4(ao + F_VX * pl, vx);
    }
    __syncthreads();  // s_v / s_phi / s_inj are complete; nobody reads ring slot `stage` any more
    if (tid == PRODUCER_TID && item + ANS * stride < nitems) produce(item + ANS * stride, stage, ds == 0 ? ANS : ds - 1);

    // ---- adjoint stress of the same quad, owner threads (el_stress_adj.cu:52-95) ----
    if (owner) {
      F4 szz = zzB, sxx = xxB, sxz = xzB;
      if (d.r1 > d.r0) {  // res_injection: szz += res, sxx += 3 res
        const F4 r = ld4(s_inj + sj);
        st4(s_inj + sj, zero4());
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {   // utilities.cu:575-579: RSXXZZ is the double literal 3.0 -> one rounding
          szz.v[kk] += r.v[kk];
          sxx.v[kk] = (float)((double)sxx.v[kk] + 3.0 * (double)r.v[kk]);
        }
      }
      const float *pz = s_v + sj;
      const float *px = pz + SCOLS * SPITCH;
      float dvz_dx[4], dvx_dz[4], dvx_dx[4], dvz_dz[4];
      dx4(ld4(pz - SPITCH), vz, ld4(pz + SPITCH), ld4(pz + 2 * SPITCH), kx1, kx2, dvz_dx);   // ad_plus_x(vz)
      dz_plus4(ld4(px - 4), vx, ld4(px + 4), kz1, kz2, dvx_dz);                              // ad_plus_z(vx)
      dx4(ld4(px - 2 * SPITCH), ld4(px - SPITCH), vx, ld4(px + SPITCH), kx1, kx2, dvx_dx);   // ad_minus_x(vx)
      dz_minus4(ld4(pz - 4), vz, ld4(pz + 4), kz1, kz2, dvz_dz);                             // ad_minus_z(vz)
      if (!pml_tile) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          sxz.v[kk] += fmaf(dvz_dx[kk], byadt.v[kk], dvx_dz[kk] * bybdt.v[kk]);
          sxx.v[kk] = fmaf(bybdt.v[kk], dvx_dx[kk], sxx.v[kk]);
          szz.v[kk] = fmaf(byadt.v[kk], dvz_dz[kk], szz.v[kk]);
        }
      } else {
        float t_xz_x[4] = {0, 0, 0, 0}, t_xz_z[4] = {0, 0, 0, 0}, t_xx[4] = {0, 0, 0, 0}, t_zz[4] = {0, 0, 0, 0};
        if (actq) {
          float dd[4];
          if (ax != 0.0f) {  // a_x * D+x(phi_xz_x)
            const float *p = s_phi + PHI_SXZ_X * SCOLS * SPITCH + sj;
            dx4(ld4(p - SPITCH), ld4(p), ld4(p + SPITCH), ld4(p + 2 * SPITCH), kx1, kx2, dd);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) t_xz_x[kk] = ax * dd[kk];
          }
          if (axh != 0.0f) {  // a_x_half * D-x(phi_xx_x)
            const float *p = s_phi + PHI_SXX_X * SCOLS * SPITCH + sj;
            dx4(ld4(p - 2 * SPITCH), ld4(p - SPITCH), ld4(p), ld4(p + SPITCH), kx1, kx2, dd);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) t_xx[kk] = axh * dd[kk];
          }
          if (zq_pml) {
            const float *p1 = s_phi + PHI_SXZ_Z * SCOLS * SPITCH + sj;   // a_z * D+z(phi_xz_z)
            dz_plus4(ld4(p1 - 4), ld4(p1), ld4(p1 + 4), kz1, kz2, dd);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) t_xz_z[kk] = az.v[kk] * dd[kk];
            const float *p2 = s_phi + PHI_SZZ_Z * SCOLS * SPITCH + sj;   // a_z_half * D-z(phi_zz_z)
            dz_minus4(ld4(p2 - 4), ld4(p2), ld4(p2 + 4), kz1, kz2, dd);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) t_zz[kk] = azh.v[kk] * dd[kk];
          }
        }
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int z = gz + kk;
          if (actq && z >= 2 && z <= g.az_hi) {
            sxz.v[kk] += t_xz_x[kk] + dvz_dx[kk] * rKx * byadt.v[kk] + t_xz_z[kk] + dvx_dz[kk] * rKz.v[kk] * bybdt.v[kk];
            sxx.v[kk] += t_xx[kk] + bybdt.v[kk] * dvx_dx[kk] * rKxh;
            szz.v[kk] += t_zz[kk] + byadt.v[kk] * dvz_dz[kk] * rKzh.v[kk];
          }
        }
        // psi memory within 2 cells of the layers (el_stress_adj.cu:68,71,89-94)
        if (actq && (gx < xq_lo || gx > xq_hi)) {
          const float *xpf = a.pr.x + gx + XM;
          const float bx = xpf[PR_B * nxp], bxh = xpf[PR_BH * nxp];
          F4 p1 = ld4(sq + (psi_i + PSI_VZ_X) * pl), p2 = ld4(sq + (psi_i + PSI_VX_X) * pl);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            p1.v[kk] = fmaf(bxh, p1.v[kk], sxz.v[kk] * amudt.v[kk]);
            p2.v[kk] = fmaf(bx, p2.v[kk], fmaf(ldt.v[kk], szz.v[kk], l2mdt.v[kk] * sxx.v[kk]));
          }
          st4(sq + (psi_o + PSI_VZ_X) * pl, p1);
          st4(sq + (psi_o + PSI_VX_X) * pl, p2);
        }
        if (actq && (gz < zq_lo || gz + 3 > zq_hi)) {
          const F4 bz = ld4(zprof + PR_B * P + gz), bzh = ld4(zprof + PR_BH * P + gz);
          F4 p1 = ld4(sq + (psi_i + PSI_VX_Z) * pl), p2 = ld4(sq + (psi_i + PSI_VZ_Z) * pl);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const int z = gz + kk;
            if (z < zq_lo || z > zq_hi) {
              p1.v[kk] = fmaf(bzh.v[kk], p1.v[kk], sxz.v[kk] * amudt.v[kk]);
              p2.v[kk] = fmaf(bz.v[kk], p2.v[kk], fmaf(l2mdt.v[kk], szz.v[kk], ldt.v[kk] * sxx.v[kk]));
            }
          }
          st4(sq + (psi_o + PSI_VX_Z) * pl, p1);
          st4(sq + (psi_o + PSI_VZ_Z) * pl, p2);
        }
      }
      st4(ao + F_SZZ * pl, szz);
      st4(ao + F_SXX * pl, sxx);
      st4(ao + F_SXZ * pl, sxz);
    }
    // s_phi and the injection table are single-buffered: everyone is done with them before the next item writes
    if (ADJ_SPLIT && !ADJ_DB) {
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&full[ANS]);
      kdone++;
    } else if (!ADJ_DB && (pml_tile || d.r1 > d.r0)) {
      __syncthreads();
    }
    nb ^= 1;
    if (++ds == ANS + 1) ds = 0;
    if (++stage == ANS) { stage = 0; phase ^= 1; }
  }
  if (a.indep) pdl_wait();
}


// =================================================================================================
// bwd_step_kernel: ONE launch per time index of the backward loop.  Per (tile, shot) work item it runs
//   phase A: the adjoint step of time index it+1 (exactly adj_step_kernel's item), then
//   phase R: the reverse-time step it+1 -> it with frame restore and the imaging condition (rev_image_kernel's item),
// on the same thread <-> quad mapping.  The imaging condition of index `it` needs the adjoint state AFTER the adjoint
// step it+1 (libCUFD.cu:374-427: el_velocity(false) / el_stress(false) run before this index's adjoint kernels), i.e.
// exactly what phase A has just produced for the quad the thread owns: the five adjoint quads go from phase A to
// phase R IN REGISTERS.  Against the two separate launches this removes the second read of the adjoint state (20 B per
// box cell from HBM, plus its L2 prefetch boxes), one launch per time index, and -- with the new adjoint velocities of
// the whole 64 x 32 region sitting in phase A's shared tile -- lets the density spray (el_velocity.cu:105-110) be
// gathered in the kernel: four accumulator planes instead of five.
//   Ring: the two phases of an item are two consecutive "sub-items" of one 2-slot TMA ring (slot = the larger of the two
//   box sets, 47.5 KB); the producer lane runs two sub-items ahead.  Tiles that lie outside the reconstruction
//   rectangle have no phase R.  Each phase has ONE block barrier and its own hand-over tiles, so the next write to a
//   tile is always separated from the last read by the other phase's barrier (tiles without a phase R add one).
// =================================================================================================
constexpr int MNS = 2;
constexpr int MSTAGE_BYTES = RSTAGE_BYTES > ASTAGE_BYTES ? RSTAGE_BYTES : ASTAGE_BYTES;
constexpr int MFRM_BYTES = NOWN * 3 * 16;                 // landing slots of the saved stress frames: owner quads only
constexpr int MGB_BYTES = SCOLS * SPITCH * 4;             // rho-b imaging term of the region (x-1 neighbour hand-over)
constexpr size_t MRG_SMEM = (size_t)MNS * MSTAGE_BYTES + AV_BYTES + APHI_BYTES + AINJ_BYTES + SV_BYTES + MFRM_BYTES + MGB_BYTES +
                            (MNS + 1) * sizeof(TileDesc) + MNS * 8 + 128;
static_assert(MSTAGE_BYTES % 128 == 0, "TMA destination alignment");
enum : int { TF_REV = 8 };   // phase-A descriptor: the tile has a phase R

__global__ void __launch_bounds__(NCOMPUTE, CTAS_PER_SM) bwd_step_kernel(const __grid_constant__ BwdArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  unsigned char *hp = base + MNS * MSTAGE_BYTES;
  float *s_va = reinterpret_cast<float *>(hp);                                   // [2][SCOLS][SPITCH] new adjoint velocities (phase A)
  float *s_phi = reinterpret_cast<float *>(hp + AV_BYTES);                       // [4][SCOLS][SPITCH] new phi (phase A)
  float *s_inj = reinterpret_cast<float *>(hp + AV_BYTES + APHI_BYTES);          // [SCOLS][SPITCH] residual injection table (phase A)
  float *s_vr = reinterpret_cast<float *>(hp + AV_BYTES + APHI_BYTES + AINJ_BYTES);              // [2][SCOLS][SPITCH] rewound velocities (phase R)
  float *s_frm = reinterpret_cast<float *>(hp + AV_BYTES + APHI_BYTES + AINJ_BYTES + SV_BYTES);  // [3][NOWN] quads: saved szz sxx sxz (phase R)
  float *s_gb = reinterpret_cast<float *>(hp + AV_BYTES + APHI_BYTES + AINJ_BYTES + SV_BYTES + MFRM_BYTES);   // [SCOLS][SPITCH] (phase R)
  unsigned char *tail = hp + AV_BYTES + APHI_BYTES + AINJ_BYTES + SV_BYTES + MFRM_BYTES + MGB_BYTES;
  TileDesc *sdesc = reinterpret_cast<TileDesc *>(tail);                          // [MNS + 1]
  uint64_t *full = reinterpret_cast<uint64_t *>(tail + (MNS + 1) * sizeof(TileDesc));   // [MNS]

  const Grid &g = a.g;
  const int tid = threadIdx.x;
  const int ntiles = g.tiles_z * g.tiles_x;
  const int nitems = a.batch * ntiles;
  const int stride = gridDim.x;   // round-robin item order (see fwd_step_kernel)
  const int it_a = a.it + 1;      // time index of the adjoint step in phase A; phase R rewinds it_a -> a.it
  const int ain = a.cur_a ? S_AB : S_AA, aout = a.cur_a ? S_AA : S_AB;
  const int psi_i = a.cur_a ? S_PSI_B : S_PSI_A, psi_o = a.cur_a ? S_PSI_A : S_PSI_B;
  const int phi_i = a.cur_a ? S_PHI_B : S_PHI_A, phi_o = a.cur_a ? S_PHI_A : S_PHI_B;
  const int fin = a.cur_f ? S_FB : S_FA, fout = a.cur_f ? S_FA : S_FB;
  const int P = g.P;
  const long long pl = g.plane;
  const int nxp = a.pr.nxp;
  const int zp_hi = g.nz - g.nPml - g.nPad - 1;
  const int zq_lo = g.nPml + 2, zq_hi = g.nz - g.nPad - g.nPml - 3;   // psi arrays matter within 2 cells of the layers (Q5)
  const int xq_lo = g.nPml + 2, xq_hi = g.nx - g.nPml - 3;

  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < MNS; s++) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  for (int i = tid; i < AINJ_BYTES / 16; i += NCOMPUTE) reinterpret_cast<float4 *>(s_inj)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  // ---- producer: the sub-item sequence A(item), [R(item)], A(item + stride), ... two sub-items ahead ----
  int p_item = blockIdx.x, p_kind = 0;   // only the producer lane's copies are used
  auto produce_next = [&](int stage, int ds, bool first) {
    if (p_item >= nitems) return;
    const int io = a.order ? nitems - 1 - p_item : p_item;
    const int tile = io / a.batch, shot = io - tile * a.batch;   // shot fastest
    const int z0 = (tile % g.tiles_z) * TILE_Z + g.z_off, x0 = (tile / g.tiles_z) * TILE_X;
    const int sz = a.st.src_z[shot], sx = a.st.src_x[shot];
    const bool has_rev = !(z0 > g.zhi + 2 || z0 + TILE_Z - 1 < g.zlo - 2 || x0 > g.xhi + 2 || x0 + TILE_X - 1 < g.xlo - 2);
    TileDesc d;
    d.soff = (long long)shot * S_COUNT * pl + (long long)x0 * P + z0;
    d.moff = x0 * P + z0;
    d.z0 = z0; d.x0 = x0; d.shot = shot; d.tile = tile; d.sz = sz; d.sx = sx;
    d.pad[0] = d.pad[1] = d.pad[2] = d.pad[3] = 0;
    unsigned char *sb = base + stage * MSTAGE_BYTES;
    int fl = 0;
    if (sz >= z0 && sz < z0 + TILE_Z && sx >= x0 && sx < x0 + TILE_X) fl |= TF_SRC;
    if (p_kind == 0) {   // ---- phase A boxes: adjoint stress triple (halo 8 / 3) + adjoint velocity pair (halo 4 / 2) ----
      d.r0 = a.st.rec_ptr[shot * (ntiles + 1) + tile];
      d.r1 = a.st.rec_ptr[shot * (ntiles + 1) + tile + 1];
      if ((z0 - 4 < zq_lo) || (z0 + TILE_Z + 3 > zq_hi) || (x0 - 2 < xq_lo) || (x0 + TILE_X + 1 > xq_hi)) fl |= TF_PML;
      if (has_rev) fl |= TF_REV;
      d.flags = fl;
      sdesc[ds] = d;
      const int p0 = shot * S_COUNT + ain;
      if (first) pdl_wait();   // everything above reads static tables only
      mbar_arrive_expect_tx(&full[stage], AS_BYTES + AV_BYTES);
      tma_load_3d(sb, &a.tm.s3, z0 - 8, x0 - 3 + XM, p0 + F_SZZ, &full[stage]);
      tma_load_3d(sb + AS_PAD, &a.tm.vn, z0 - 4, x0 - 2 + XM, p0 + F_VZ, &full[stage]);
#if ADJ_PF_MODEL
      if (ADJ_PF_MODEL == 1 || shot == 0) tma_prefetch_3d(&a.tm.m5, z0 - 4, x0 - 2 + XM, M_LDT);
#endif
#if FWI_L2PF
      if (fl & TF_PML) {  // CPML memory of the layers this tile touches -> L2
        const int ps = shot * S_COUNT;
        if ((z0 - 4 < zq_lo) || (z0 + TILE_Z + 3 > zq_hi)) {
          tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + psi_i + PSI_VX_Z);
          tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + psi_i + PSI_VZ_Z);
          tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + phi_i + PHI_SXZ_Z);
          tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + phi_i + PHI_SZZ_Z);
        }
        if ((x0 - 2 < xq_lo) || (x0 + TILE_X + 1 > xq_hi)) {
          tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + psi_i + PSI_VX_X);
          tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + psi_i + PSI_VZ_X);
          tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + phi_i + PHI_SXX_X);
          tma_prefetch_3d(&a.tm.r1, z0 - 4, x0 - 2 + XM, ps + phi_i + PHI_SXZ_X);
        }
      }
#endif
      if (has_rev) p_kind = 1; else p_item += stride;
    } else {   // ---- phase R boxes: forward stress triple of time it+1 (halo 8 / 4) + forward velocity pair (halo 4 / 2) ----
      d.r0 = d.r1 = 0;
      if (!(z0 - 4 > g.zlo - 1 + g.f_in && z0 + TILE_Z + 3 < g.zhi + 1 - g.f_in && x0 - 2 > g.xlo - 1 + g.f_in &&
          x0 + TILE_X + 1 < g.xhi + 1 - g.f_in))
      fl |= TF_FRAME;
      d.flags = fl;
      sdesc[ds] = d;
      const int p0 = shot * S_COUNT + fin;
      mbar_arrive_expect_tx(&full[stage], RSTAGE_BYTES);
      tma_load_3d(sb, &a.tm.sw, z0 - 8, x0 - 4 + XM, p0 + F_SZZ, &full[stage]);
      tma_load_3d(sb + RW_BYTES, &a.tm.vn, z0 - 4, x0 - 2 + XM, p0 + F_VZ, &full[stage]);
#if FWI_L2PF
      tma_prefetch_3d(&a.tm.g4, z0, x0 + XM, shot * G_COUNT);   // the four accumulator planes of the owner tile -> L2
#endif
      p_kind = 0;
      p_item += stride;
    }
  };
  if (tid == PRODUCER_TID)
    for (int s = 0; s < MNS; s++) produce_next(s, s, s == 0);
  __syncthreads();   // the first descriptors are visible
  pdl_wait();

  const float dt = g.dt;
  // adjoint-kernel spelling of the differences: (-c1 (..) + c2 (..)) / h  (el_stress_adj.cu:54-61)
  const float akz1 = -C1 * g.rdz, akz2 = -C2 * g.rdz, akx1 = -C1 * g.rdx, akx2 = -C2 * g.rdx;
  const float half_rdt = 0.5f / dt;            // 0.5 byc^2 dt      = (0.5 / dt) (byc dt)^2
  const float q_rdt = 250000.0f / dt;          // 1e6 mu_bar^2 dt/4 = (250000 / dt) (mu_bar dt)^2
  const float dt6 = dt * 1e6f;
  const int q = tid & 15, c = tid >> 4;
  const bool inner = q >= 1 && q <= TILE_Z / 4 && c >= 2 && c < TILE_X + 2;
  const int sj = c * SPITCH + 4 * q;
  const int gx_max = g.nx + XM - 1;
  const int cm2 = (c > 0 ? 2 : 1) * VPITCH, cp2 = (c < SCOLS - 1 ? 2 : 1) * VPITCH;   // keep halo-column reads in the tile
  const float *zprof = a.pr.z;
  float *my_frm = s_frm + 4 * ((q - 1) + (TILE_Z / 4) * (c - 2));   // owner quads only (nobody else uses theirs)

  int stage = 0, phase = 0, ds = 0;
  auto advance = [&]() {
    if (++ds == MNS + 1) ds = 0;
    if (++stage == MNS) { stage = 0; phase ^= 1; }
  };
  for (int item = blockIdx.x; item < nitems; item += stride) {
    // =============================== phase A: adjoint step of time index it_a ===============================
    const TileDesc d = sdesc[ds];   // written by the producer >= 1 block barrier ago
    const int gz = d.z0 - 4 + 4 * q, gx = d.x0 - 2 + c;
    const bool inb = (unsigned)gx < (unsigned)g.nx && (unsigned)gz < (unsigned)g.zlive;
    const bool owner = inner && inb;
    float *sq = a.state + g.origin + d.soff + ((long long)(c - 2) * P + 4 * q - 4);   // + slot * pl
    const float *mq = a.m.ldt + ((long long)min(gx, gx_max) * P + gz);
    const bool pml_tile = d.flags & TF_PML;
    const bool has_rev = d.flags & TF_REV;
    // quad with at least one active cell (2 <= z <= nz-nPad-3, 2 <= x <= nx-3): the only ones that touch CPML memory
    const bool actq = gx >= 2 && gx <= g.ax_hi && gz + 3 >= 2 && gz <= g.az_hi;
    const F4 ldt = ld4(mq), l2mdt = ld4(mq + pl), amudt = ld4(mq + 2 * pl);
    const F4 byadt = ld4(mq + 3 * pl), bybdt = ld4(mq + 4 * pl);
    // double increments next to the source in the ADJOINT kernels too (FWI_F64_ADJ): off -- measured to make no
    // difference to grad_stf, which is decided by the forward arithmetic (profiles/r2_parity.md)
    const bool near_src = FWI_F64_ADJ && FWI_F64_UPDATE > 1 && abs(gx - d.sx) <= FWI_F64_UPDATE &&
                          gz + 3 >= d.sz - FWI_F64_UPDATE && gz <= d.sz + FWI_F64_UPDATE;
    mbar_wait(&full[stage], phase);

    F4 vz, vx;          // adjoint velocities of the quad: pre-update, then new
    F4 szz, sxx, sxz;   // adjoint stresses of the quad (owner threads): pre-update, then new
    {
      const unsigned char *sb = base + stage * MSTAGE_BYTES;
      const float *sa = reinterpret_cast<const float *>(sb);              // [3][VCOLS][VPITCH]: adjoint szz sxx sxz
      const float *sva = reinterpret_cast<const float *>(sb + AS_PAD);    // [2][SCOLS][SPITCH]: adjoint vz vx

      // ---- adjoint velocity on 16 quads x 32 columns (el_velocity_adj.cu:56-100) ----
      const float *zz = sa + (c + 1) * VPITCH + 4 * (q + 1);
      const float *xx = zz + VCOLS * VPITCH;
      const float *xz = xx + VCOLS * VPITCH;
      const F4 zzB = ld4(zz), xxB = ld4(xx), xzB = ld4(xz);
      float dszz_dx[4], dsxx_dx[4], dsxz_dz[4], dszz_dz[4], dsxx_dz[4], dsxz_dx[4];
      dx4(ld4(zz - VPITCH), zzB, ld4(zz + VPITCH), ld4(zz + cp2), akx1, akx2, dszz_dx);   // ad_plus_x
      dx4(ld4(xx - VPITCH), xxB, ld4(xx + VPITCH), ld4(xx + cp2), akx1, akx2, dsxx_dx);
      dx4(ld4(xz - cm2), ld4(xz - VPITCH), xzB, ld4(xz + VPITCH), akx1, akx2, dsxz_dx);   // ad_minus_x
      dz_plus4(ld4(zz - 4), zzB, ld4(zz + 4), akz1, akz2, dszz_dz);
      dz_plus4(ld4(xx - 4), xxB, ld4(xx + 4), akz1, akz2, dsxx_dz);
      dz_minus4(ld4(xz - 4), xzB, ld4(xz + 4), akz1, akz2, dsxz_dz);
      vz = ld4(sva + sj);
      vx = ld4(sva + SCOLS * SPITCH + sj);
      szz = zzB; sxx = xxB; sxz = xzB;

      // source_grad (utilities.cu:582-593): adjoint stress at the source BEFORE this step's injection and update
      if ((d.flags & TF_SRC) && owner && gx == d.sx && (unsigned)(d.sz - gz) < 4u) {
        const int ks = d.sz - gz;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
          if (kk == ks) { s1 = zzB.v[kk]; s2 = xxB.v[kk]; }
        a.stf_grad[d.shot * g.nSteps + it_a] = (float)(-((double)s1 + 3.0 * (double)s2) * (double)dt);
      }
      // residual injection at time index it_a (utilities.cu:569-580): receivers of this tile add into the table
      for (int r = d.r0 + tid; r < d.r1; r += NCOMPUTE) {
        const int loc = a.st.rec_loc[d.shot * a.st.nrp + r];
        const int lz = loc & 0xffff, lx = loc >> 16;
        atomicAdd(&s_inj[(lx + 2) * SPITCH + lz + 4],
                  a.res[((long long)d.shot * g.nSteps + it_a) * a.st.nrp + a.st.rec_id[d.shot * a.st.nrp + r]]);
      }

      float rKx = 1.0f, rKxh = 1.0f, ax = 0.0f, axh = 0.0f;
      F4 rKz{{1.f, 1.f, 1.f, 1.f}}, rKzh{{1.f, 1.f, 1.f, 1.f}}, az = zero4(), azh = zero4();
      bool zq_pml = false, xp = false;
      if (!pml_tile) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {  // coefficients carry dt and are 0 on inactive cells
          // el_velocity_adj.cu:69-71,90-92: the (lambda + 2.0 mu) term promotes the sum to double
          if (FWI_F64_UPDATE == 1 || (FWI_F64_ADJ && FWI_F64_UPDATE > 1 && near_src)) {
            vx.v[kk] = (float)((double)vx.v[kk] + ((double)ldt.v[kk] * (double)dszz_dx[kk] + (double)l2mdt.v[kk] * (double)dsxx_dx[kk] +
                                                   (double)amudt.v[kk] * (double)dsxz_dz[kk]));
            vz.v[kk] = (float)((double)vz.v[kk] + ((double)l2mdt.v[kk] * (double)dszz_dz[kk] + (double)ldt.v[kk] * (double)dsxx_dz[kk] +
                                                   (double)amudt.v[kk] * (double)dsxz_dx[kk]));
          } else {
            vx.v[kk] += fmaf(ldt.v[kk], dszz_dx[kk], fmaf(l2mdt.v[kk], dsxx_dx[kk], amudt.v[kk] * dsxz_dz[kk]));
            vz.v[kk] += fmaf(l2mdt.v[kk], dszz_dz[kk], fmaf(ldt.v[kk], dsxx_dz[kk], amudt.v[kk] * dsxz_dx[kk]));
          }
        }
      } else {
        F4 f_szz_z = zero4(), f_sxz_x = zero4(), f_sxz_z = zero4(), f_sxx_x = zero4();   // new phi of the quad
        if (actq) {
          float tpx1[4] = {0, 0, 0, 0}, tpx2[4] = {0, 0, 0, 0}, tpz1[4] = {0, 0, 0, 0}, tpz2[4] = {0, 0, 0, 0};
          const float *xpf = a.pr.x + gx + XM;
          rKx = xpf[PR_RK * nxp];
          rKxh = xpf[PR_RKH * nxp];
          ax = xpf[PR_A * nxp];
          axh = xpf[PR_AH * nxp];
          xp = gx < g.nPml || gx > g.nx - g.nPml - 1;
          zq_pml = gz < g.nPml || gz + 3 > zp_hi;
          if (ax != 0.0f) {  // a_x * D+x(psi_xx)
            const float *p = sq + (psi_i + PSI_VX_X) * pl;
            float dd[4];
            dx4(ld4(p - P), ld4(p), ld4(p + P), ld4(p + 2 * P), akx1, akx2, dd);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) tpx1[kk] = ax * dd[kk];
          }
          if (axh != 0.0f) {  // a_x_half * D-x(psi_zx)
            const float *p = sq + (psi_i + PSI_VZ_X) * pl;
            float dd[4];
            dx4(ld4(p - 2 * P), ld4(p - P), ld4(p), ld4(p + P), akx1, akx2, dd);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) tpz2[kk] = axh * dd[kk];
          }
          if (zq_pml) {
            rKz = ld4(zprof + PR_RK * P + gz);
            rKzh = ld4(zprof + PR_RKH * P + gz);
            az = ld4(zprof + PR_A * P + gz);
            azh = ld4(zprof + PR_AH * P + gz);
            const float *p1 = sq + (psi_i + PSI_VX_Z) * pl;  // a_z_half * D-z(psi_xz)
            const float *p2 = sq + (psi_i + PSI_VZ_Z) * pl;  // a_z * D+z(psi_zz)
            float dd[4];
            dz_minus4(ld4(p1 - 4), ld4(p1), ld4(p1 + 4), akz1, akz2, dd);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) tpx2[kk] = azh.v[kk] * dd[kk];
            dz_plus4(ld4(p2 - 4), ld4(p2), ld4(p2 + 4), akz1, akz2, dd);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) tpz1[kk] = az.v[kk] * dd[kk];
          }
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const int z = gz + kk;
            if (z >= 2 && z <= g.az_hi) {
              if (FWI_F64_UPDATE == 1 || (FWI_F64_ADJ && FWI_F64_UPDATE > 1 && near_src)) {
                vx.v[kk] = (float)((double)vx.v[kk] + ((double)tpx1[kk] + (double)(ldt.v[kk] * dszz_dx[kk] * rKx) +
                                                       (double)l2mdt.v[kk] * (double)(dsxx_dx[kk] * rKx) + (double)tpx2[kk] +
                                                       (double)(amudt.v[kk] * rKzh.v[kk] * dsxz_dz[kk])));
                vz.v[kk] = (float)((double)vz.v[kk] + ((double)tpz1[kk] + (double)l2mdt.v[kk] * (double)(dszz_dz[kk] * rKz.v[kk]) +
                                                       (double)(ldt.v[kk] * dsxx_dz[kk] * rKz.v[kk]) + (double)tpz2[kk] +
                                                       (double)(amudt.v[kk] * rKxh * dsxz_dx[kk])));
              } else {
                vx.v[kk] += tpx1[kk] + ldt.v[kk] * dszz_dx[kk] * rKx + l2mdt.v[kk] * dsxx_dx[kk] * rKx + tpx2[kk] +
                            amudt.v[kk] * rKzh.v[kk] * dsxz_dz[kk];
                vz.v[kk] += tpz1[kk] + l2mdt.v[kk] * dszz_dz[kk] * rKz.v[kk] + ldt.v[kk] * dsxx_dz[kk] * rKz.v[kk] + tpz2[kk] +
                            amudt.v[kk] * rKxh * dsxz_dx[kk];
              }
            }
          }
          // phi memory of the quad, CPML cells only (el_velocity_adj.cu:74-79,95-100); buoyancies are 0 on inactive cells
          if (xp) {
            const float bx = xpf[PR_B * nxp], bxh = xpf[PR_BH * nxp];
            f_sxx_x = ld4s(sq + (phi_i + PHI_SXX_X) * pl);
            f_sxz_x = ld4s(sq + (phi_i + PHI_SXZ_X) * pl);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              f_sxx_x.v[kk] = fmaf(bxh, f_sxx_x.v[kk], bybdt.v[kk] * vx.v[kk]);
              f_sxz_x.v[kk] = fmaf(bx, f_sxz_x.v[kk], byadt.v[kk] * vz.v[kk]);
            }
            if (owner) {
              st4(sq + (phi_o + PHI_SXX_X) * pl, f_sxx_x);
              st4(sq + (phi_o + PHI_SXZ_X) * pl, f_sxz_x);
            }
          }
          if (zq_pml) {
            const F4 bz = ld4(zprof + PR_B * P + gz), bzh = ld4(zprof + PR_BH * P + gz);
            f_sxz_z = ld4s(sq + (phi_i + PHI_SXZ_Z) * pl);
            f_szz_z = ld4s(sq + (phi_i + PHI_SZZ_Z) * pl);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              const int z = gz + kk;
              if (z < g.nPml || z > zp_hi) {
                f_sxz_z.v[kk] = fmaf(bz.v[kk], f_sxz_z.v[kk], bybdt.v[kk] * vx.v[kk]);
                f_szz_z.v[kk] = fmaf(bzh.v[kk], f_szz_z.v[kk], byadt.v[kk] * vz.v[kk]);
              }
            }
            if (owner) {
              st4(sq + (phi_o + PHI_SXZ_Z) * pl, f_sxz_z);
              st4(sq + (phi_o + PHI_SZZ_Z) * pl, f_szz_z);
            }
          }
        }
        st4(s_phi + PHI_SZZ_Z * SCOLS * SPITCH + sj, f_szz_z);
        st4(s_phi + PHI_SXZ_X * SCOLS * SPITCH + sj, f_sxz_x);
        st4(s_phi + PHI_SXZ_Z * SCOLS * SPITCH + sj, f_sxz_z);
        st4(s_phi + PHI_SXX_X * SCOLS * SPITCH + sj, f_sxx_x);
      }
      st4(s_va + sj, vz);
      st4(s_va + SCOLS * SPITCH + sj, vx);
      float *ao = sq + aout * pl;
      if (owner) {
        st4(ao + F_VZ * pl, vz);
        st4(ao + F_VX * pl, vx);
      }
      __syncthreads();  // s_va / s_phi / s_inj are complete; nobody reads ring slot `stage` any more
      if (tid == PRODUCER_TID) produce_next(stage, ds == 0 ? MNS : ds - 1, false);

      // ---- adjoint stress of the same quad, owner threads (el_stress_adj.cu:52-95) ----
      if (owner) {
        if (d.r1 > d.r0) {  // res_injection: szz += res, sxx += 3 res
          const F4 r = ld4(s_inj + sj);
          st4(s_inj + sj, zero4());
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {   // utilities.cu:575-579: RSXXZZ is the double literal 3.0 -> one rounding
            szz.v[kk] += r.v[kk];
            sxx.v[kk] = (float)((double)sxx.v[kk] + 3.0 * (double)r.v[kk]);
          }
        }
        const float *pz = s_va + sj;
        const float *px = pz + SCOLS * SPITCH;
        float dvz_dx[4], dvx_dz[4], dvx_dx[4], dvz_dz[4];
        dx4(ld4(pz - SPITCH), vz, ld4(pz + SPITCH), ld4(pz + 2 * SPITCH), akx1, akx2, dvz_dx);   // ad_plus_x(vz)
        dz_plus4(ld4(px - 4), vx, ld4(px + 4), akz1, akz2, dvx_dz);                              // ad_plus_z(vx)
        dx4(ld4(px - 2 * SPITCH), ld4(px - SPITCH), vx, ld4(px + SPITCH), akx1, akx2, dvx_dx);   // ad_minus_x(vx)
        dz_minus4(ld4(pz - 4), vz, ld4(pz + 4), akz1, akz2, dvz_dz);                             // ad_minus_z(vz)
        if (!pml_tile) {
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            sxz.v[kk] += fmaf(dvz_dx[kk], byadt.v[kk], dvx_dz[kk] * bybdt.v[kk]);
            sxx.v[kk] = fmaf(bybdt.v[kk], dvx_dx[kk], sxx.v[kk]);
            szz.v[kk] = fmaf(byadt.v[kk], dvz_dz[kk], szz.v[kk]);
          }
        } else {
          float t_xz_x[4] = {0, 0, 0, 0}, t_xz_z[4] = {0, 0, 0, 0}, t_xx[4] = {0, 0, 0, 0}, t_zz[4] = {0, 0, 0, 0};
          if (actq) {
            float dd[4];
            if (ax != 0.0f) {  // a_x * D+x(phi_xz_x)
              const float *p = s_phi + PHI_SXZ_X * SCOLS * SPITCH + sj;
              dx4(ld4(p - SPITCH), ld4(p), ld4(p + SPITCH), ld4(p + 2 * SPITCH), akx1, akx2, dd);
#pragma unroll
              for (int kk = 0; kk < 4; kk++) t_xz_x[kk] = ax * dd[kk];
            }
            if (axh != 0.0f) {  // a_x_half * D-x(phi_xx_x)
              const float *p = s_phi + PHI_SXX_X * SCOLS * SPITCH + sj;
              dx4(ld4(p - 2 * SPITCH), ld4(p - SPITCH), ld4(p), ld4(p + SPITCH), akx1, akx2, dd);
#pragma unroll
              for (int kk = 0; kk < 4; kk++) t_xx[kk] = axh * dd[kk];
            }
            if (zq_pml) {
              const float *p1 = s_phi + PHI_SXZ_Z * SCOLS * SPITCH + sj;   // a_z * D+z(phi_xz_z)
              dz_plus4(ld4(p1 - 4), ld4(p1), ld4(p1 + 4), akz1, akz2, dd);
#pragma unroll
              for (int kk = 0; kk < 4; kk++) t_xz_z[kk] = az.v[kk] * dd[kk];
              const float *p2 = s_phi + PHI_SZZ_Z * SCOLS * SPITCH + sj;   // a_z_half * D-z(phi_zz_z)
              dz_minus4(ld4(p2 - 4), ld4(p2), ld4(p2 + 4), akz1, akz2, dd);
#pragma unroll
              for (int kk = 0; kk < 4; kk++) t_zz[kk] = azh.v[kk] * dd[kk];
            }
          }
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const int z = gz + kk;
            if (actq && z >= 2 && z <= g.az_hi) {
              sxz.v[kk] += t_xz_x[kk] + dvz_dx[kk] * rKx * byadt.v[kk] + t_xz_z[kk] + dvx_dz[kk] * rKz.v[kk] * bybdt.v[kk];
              sxx.v[kk] += t_xx[kk] + bybdt.v[kk] * dvx_dx[kk] * rKxh;
              szz.v[kk] += t_zz[kk] + byadt.v[kk] * dvz_dz[kk] * rKzh.v[kk];
            }
          }
          // psi memory within 2 cells of the layers (el_stress_adj.cu:68,71,89-94)
          if (actq && (gx < xq_lo || gx > xq_hi)) {
            const float *xpf = a.pr.x + gx + XM;
            const float bx = xpf[PR_B * nxp], bxh = xpf[PR_BH * nxp];
            F4 p1 = ld4(sq + (psi_i + PSI_VZ_X) * pl), p2 = ld4(sq + (psi_i + PSI_VX_X) * pl);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              p1.v[kk] = fmaf(bxh, p1.v[kk], sxz.v[kk] * amudt.v[kk]);
              p2.v[kk] = fmaf(bx, p2.v[kk], fmaf(ldt.v[kk], szz.v[kk], l2mdt.v[kk] * sxx.v[kk]));
            }
            st4(sq + (psi_o + PSI_VZ_X) * pl, p1);
            st4(sq + (psi_o + PSI_VX_X) * pl, p2);
          }
          if (actq && (gz < zq_lo || gz + 3 > zq_hi)) {
            const F4 bz = ld4(zprof + PR_B * P + gz), bzh = ld4(zprof + PR_BH * P + gz);
            F4 p1 = ld4(sq + (psi_i + PSI_VX_Z) * pl), p2 = ld4(sq + (psi_i + PSI_VZ_Z) * pl);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              const int z = gz + kk;
              if (z < zq_lo || z > zq_hi) {
                p1.v[kk] = fmaf(bzh.v[kk], p1.v[kk], sxz.v[kk] * amudt.v[kk]);
                p2.v[kk] = fmaf(bz.v[kk], p2.v[kk], fmaf(l2mdt.v[kk], szz.v[kk], ldt.v[kk] * sxx.v[kk]));
              }
            }
            st4(sq + (psi_o + PSI_VX_Z) * pl, p1);
            st4(sq + (psi_o + PSI_VZ_Z) * pl, p2);
          }
        }
        st4(ao + F_SZZ * pl, szz);
        st4(ao + F_SXX * pl, sxx);
        st4(ao + F_SXZ * pl, sxz);
      }
    }
    advance();
    if (!has_rev) {
      __syncthreads();   // no phase R (and its barrier) between this item's reads of the phase-A tiles and the next item's writes
      continue;
    }

    // ========== phase R: forward state it+1 -> it inside the inner box, frame restore, imaging with the NEW adjoint quads ==========
    {
      const TileDesc dr = sdesc[ds];   // same tile / shot; carries the frame flag
      float *acc = a.gacc + g.origin + (long long)d.shot * (G_COUNT - S_COUNT) * pl + d.soff +
                   ((long long)(c - 2) * P + 4 * q - 4);   // shot * G_COUNT * pl + cell
      const bool colbox = gx >= g.xlo && gx <= g.xhi;
      bool bx[4];   // cell inside the inner box (reconstruction / imaging region)
#pragma unroll
      for (int kk = 0; kk < 4; kk++) bx[kk] = colbox && (unsigned)(gz + kk - g.zlo) <= (unsigned)(g.zhi - g.zlo);
      const bool in_rect = gx >= g.xlo - 2 && gx <= g.xhi + 2 && gz + 3 >= g.zlo - 2 && gz <= g.zhi + 2;
      // saved frame values of the quad (to_bnd, libCUFD.cu:388,403), copied asynchronously (LDGSTS, no registers): the
      // velocities straight into the quad's place in the rewound-velocity tile, the stresses into the owner's landing zone
      int fq = -1;
      if ((dr.flags & TF_FRAME) && in_rect) fq = frame_quad(g, gz, gx);
      if (fq >= 0) {
        const float *frm = a.frames + ((long long)d.shot * g.nSteps + a.it) * 5 * g.f_len + 4 * fq;
        cp_async16(s_vr + sj, frm + F_VZ * g.f_len);
        cp_async16(s_vr + SCOLS * SPITCH + sj, frm + F_VX * g.f_len);
        if (inner) {
#pragma unroll
          for (int f = 0; f < 3; f++) cp_async16(my_frm + f * 4 * NOWN, frm + (F_SZZ + f) * g.f_len);
        }
      }
      mbar_wait(&full[stage], phase);

      const unsigned char *sb = base + stage * MSTAGE_BYTES;
      const float *sw = reinterpret_cast<const float *>(sb);              // [3][WCOLS][VPITCH]: szz sxx sxz of time it+1
      const float *sv = reinterpret_cast<const float *>(sb + RW_BYTES);   // [2][SCOLS][SPITCH]: vz vx of time it+1
      const float kz1 = -akz1, kz2 = -akz2, kx1 = -akx1, kx2 = -akx2;

      // ---- v^{it} = v^{it+1} - velocity(sigma^{it+1}) on 16 quads x 32 columns; rho imaging terms (el_velocity.cu:84-110) ----
      const float *zz = sw + (c + 2) * VPITCH + 4 * (q + 1);
      const float *xx = zz + WCOLS * VPITCH;
      const float *xz = xx + WCOLS * VPITCH;
      float ea[4], eb[4];
      const F4 szzB = ld4(zz), sxxB = ld4(xx), sxzB = ld4(xz);
      {
        float d1[4], d2[4];
        dz_plus4(ld4(zz - 4), szzB, ld4(zz + 4), kz1, kz2, d1);                       // dszz_dz
        dx4(ld4(xz - 2 * VPITCH), ld4(xz - VPITCH), sxzB, ld4(xz + VPITCH), kx1, kx2, d2);   // dsxz_dx
#pragma unroll
        for (int kk = 0; kk < 4; kk++) ea[kk] = d1[kk] + d2[kk];
        dz_minus4(ld4(xz - 4), sxzB, ld4(xz + 4), kz1, kz2, d1);                      // dsxz_dz
        dx4(ld4(xx - VPITCH), sxxB, ld4(xx + VPITCH), ld4(xx + 2 * VPITCH), kx1, kx2, d2);   // dsxx_dx
#pragma unroll
        for (int kk = 0; kk < 4; kk++) eb[kk] = d1[kk] + d2[kk];
      }
      F4 fvz = ld4(sv + sj), fvx = ld4(sv + SCOLS * SPITCH + sj);
      F4 ga = zero4(), gb = zero4();
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        if (bx[kk]) {
          fvz.v[kk] = fmaf(-ea[kk], byadt.v[kk], fvz.v[kk]);
          fvx.v[kk] = fmaf(-eb[kk], bybdt.v[kk], fvx.v[kk]);
          // g = -v_adj (d sigma) dt * (-byc^2 / 2)     (el_velocity.cu:101-104); vz / vx = the new adjoint velocities
          ga.v[kk] = (vz.v[kk] * ea[kk]) * (half_rdt * byadt.v[kk] * byadt.v[kk]);
          gb.v[kk] = (vx.v[kk] * eb[kk]) * (half_rdt * bybdt.v[kk] * bybdt.v[kk]);
        }
      }
      // the spray of el_velocity.cu:105-110 as a gather: cell (z, x) receives g_a(z, x) + g_b(z, x) + g_a(z-1, x) +
      // g_b(z, x-1).  g_a of the cell above the quad comes from the thread above (same half-warp), g_b of the column
      // to the left through shared memory after the barrier.
      const float ga_up = __shfl_up_sync(0xffffffffu, ga.v[3], 1, 16);
      st4(s_gb + sj, gb);
      if (fq >= 0) {  // exact values of time `it` on the ring: already in the shared tile
        cp_async_wait_all();
        fvz = ld4(s_vr + sj);
        fvx = ld4(s_vr + SCOLS * SPITCH + sj);
      } else {
        st4(s_vr + sj, fvz);
        st4(s_vr + SCOLS * SPITCH + sj, fvx);
      }
      float *fo = sq + fout * pl;
      const bool wr = owner && in_rect;
      if (wr) {
        st4(fo + F_VZ * pl, fvz);
        st4(fo + F_VX * pl, fvx);
      }
      // accumulators of the stress half, requested before the barrier
      F4 gl, gm, gs, gd;
      const bool colrho = gx >= g.xlo && gx <= g.xhi + 1;   // the x+1 spray also lands in column xhi + 1 (Q2)
      const bool rowbox = gz + 3 >= g.zlo && gz <= g.zhi;
      if (wr && rowbox) {
        if (colbox) { gl = ld4s(acc + G_LAM * pl); gm = ld4s(acc + G_MU * pl); gs = ld4s(acc + G_MUS * pl); }
        if (colrho) gd = ld4s(acc + G_RHO * pl);
      }
      __syncthreads();  // s_vr / s_gb are complete; nobody reads ring slot `stage` any more
      if (tid == PRODUCER_TID) produce_next(stage, ds == 0 ? MNS : ds - 1, false);

      // ---- sigma^{it} = sigma^{it+1} - source - stress(v^{it}) on the owner quads; lambda / mu imaging (el_stress.cu:90-124) ----
      if (wr) {
        F4 fzz = szzB, fxx = sxxB, fxz = sxzB;
        if (rowbox && colrho) {
          const F4 gbl = ld4(s_gb + sj - SPITCH);
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const float up = kk == 0 ? ga_up : ga.v[kk - 1];
            // rows outside the box receive nothing: g_a / g_b are zero there, and the z+1 spray stops at zhi (el_velocity.cu:107)
            const bool rowin = (unsigned)(gz + kk - g.zlo) <= (unsigned)(g.zhi - g.zlo);
            gd.v[kk] += rowin ? (ga.v[kk] + gb.v[kk]) + (up + gbl.v[kk]) : 0.0f;
          }
          st4(acc + G_RHO * pl, gd);
        }
        if (rowbox && colbox) {
          const float *pz = s_vr + sj;
          const float *px = pz + SCOLS * SPITCH;
          float dvz_dz[4], dvx_dz[4], dvx_dx[4], dvz_dx[4];
          dz_minus4(ld4(pz - 4), fvz, ld4(pz + 4), kz1, kz2, dvz_dz);
          dz_plus4(ld4(px - 4), fvx, ld4(px + 4), kz1, kz2, dvx_dz);
          dx4(ld4(px - 2 * SPITCH), ld4(px - SPITCH), fvx, ld4(px + SPITCH), kx1, kx2, dvx_dx);
          dx4(ld4(pz - SPITCH), fvz, ld4(pz + SPITCH), ld4(pz + 2 * SPITCH), kx1, kx2, dvz_dx);
          if ((dr.flags & TF_SRC) && gx == d.sx && (unsigned)(d.sz - gz) < 4u) {  // add_source(isFor=false): utilities.cu:538-551
            const float amp = a.st.stf[d.shot * g.nSteps + a.it];
            const float azz = SRC_SCALE * amp * dt;
            const double axx = 3.0 * (double)SRC_SCALE * (double)amp * (double)dt;
            const int ks = d.sz - gz;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              fzz.v[kk] -= (kk == ks) ? azz : 0.0f;
              fxx.v[kk] = (kk == ks) ? (float)((double)fxx.v[kk] - axx) : fxx.v[kk];
            }
          }
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            if (bx[kk]) {
#if FWI_F64_UPDATE == 1
              fzz.v[kk] = (float)((double)fzz.v[kk] - ((double)l2mdt.v[kk] * (double)dvz_dz[kk] + (double)ldt.v[kk] * (double)dvx_dx[kk]));
              fxx.v[kk] = (float)((double)fxx.v[kk] - ((double)ldt.v[kk] * (double)dvz_dz[kk] + (double)l2mdt.v[kk] * (double)dvx_dx[kk]));
#else
              fzz.v[kk] = fmaf(-l2mdt.v[kk], dvz_dz[kk], fmaf(-ldt.v[kk], dvx_dx[kk], fzz.v[kk]));
              fxx.v[kk] = fmaf(-l2mdt.v[kk], dvx_dx[kk], fmaf(-ldt.v[kk], dvz_dz[kk], fxx.v[kk]));
#endif
              const float e = dvx_dz[kk] + dvz_dx[kk];
              fxz.v[kk] = fmaf(-amudt.v[kk], e, fxz.v[kk]);
              // el_stress.cu:109-116; szz / sxx / sxz = the new adjoint stresses of the quad
              gl.v[kk] += -(szz.v[kk] + sxx.v[kk]) * (dvz_dz[kk] + dvx_dx[kk]) * dt6;
              gm.v[kk] += (-2.0f * szz.v[kk] * dvz_dz[kk] - 2.0f * sxx.v[kk] * dvx_dx[kk]) * dt6;
              //  s = -sxz_adj (exz + ezx) dt mu_bar / sum(1/mu) 1e6, mu_bar / sum(1/mu) == mu_bar^2 / 4; zero where mu_bar == 0
              gs.v[kk] += -sxz.v[kk] * e * (q_rdt * amudt.v[kk] * amudt.v[kk]);
            }
          }
          st4(acc + G_LAM * pl, gl);
          st4(acc + G_MU * pl, gm);
          st4(acc + G_MUS * pl, gs);
        }
        if (fq >= 0) {  // to_bnd(sigma) (libCUFD.cu:403)
          fzz = ld4(my_frm);
          fxx = ld4(my_frm + 4 * NOWN);
          fxz = ld4(my_frm + 8 * NOWN);
        }
        st4(fo + F_SZZ * pl, fzz);
        st4(fo + F_SXX * pl, fxx);
        st4(fo + F_SXZ * pl, fxz);
      }
    }
    advance();
  }
}

}  // namespace

// -1: pick the reverse kernel build by working-set size; 0 / 1: force the double-buffered / LEAN build
std::atomic<int> g_rev_lean_force{-1};
std::atomic<int> g_acc_group_force{0};   // 0: automatic (reverse_acc_group)
bool reverse_is_lean(const Grid &g, int batch);
void set_rev_lean(int v) { g_rev_lean_force.store(v, std::memory_order_relaxed); }
void set_acc_group(int v) { g_acc_group_force.store(v, std::memory_order_relaxed); }

size_t reverse_smem_bytes() { return REV_SMEM; }

size_t adjoint_smem_bytes() { return ADJ_SMEM; }

size_t merged_smem_bytes() { return MRG_SMEM; }

// backward loop of one time index in ONE launch: adjoint step it+1, reverse step it+1 -> it + imaging (see bwd_step_kernel)
void launch_backward_merged(const BwdArgs &a_in, cudaStream_t s) {
  BwdArgs a = a_in;
  a.order = FWI_ZIGZAG ? (a.it & 1) : 0;
  const int nitems = a.batch * a.g.tiles_z * a.g.tiles_x;
  const int blocks = nitems < sm_count() * CTAS_PER_SM ? nitems : sm_count() * CTAS_PER_SM;
  launch_step(bwd_step_kernel, blocks, NCOMPUTE, MRG_SMEM, s, a);
}

void configure_backward_kernels() {
  cudaFuncSetAttribute(bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MRG_SMEM);
  cudaFuncSetAttribute(rev_image_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rev_smem<false>());
  cudaFuncSetAttribute(rev_image_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rev_smem<true>());
  cudaFuncSetAttribute(rev_image_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rev_smem<false>() + racc_bytes<false>());
  cudaFuncSetAttribute(rev_image_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rev_smem<true>() + racc_bytes<true>());
  cudaFuncSetAttribute(adj_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ADJ_SMEM);
}

// Item order (FWI_ZIGZAG): the adjoint step runs ascending, the reverse step -- which reads the adjoint state the
// adjoint step has just written, and whose output the next adjoint step's neighbourhood in L2 follows -- descending.
void launch_adjoint_step(const BwdArgs &a_in, cudaStream_t s) {
  BwdArgs a = a_in;
  a.order = 0;
  const int nitems = a.batch * a.g.tiles_z * a.g.tiles_x;
  const int blocks = nitems < sm_count() * CTAS_PER_SM ? nitems : sm_count() * CTAS_PER_SM;
  launch_step(adj_step_kernel, blocks, NCOMPUTE, ADJ_SMEM, s, a);
}

void launch_reverse_imaging(const BwdArgs &a_in, cudaStream_t s) {
  BwdArgs a = a_in;
  a.order = FWI_ZIGZAG ? 1 : 0;
  const Grid &g = a.g;
  const int tz0 = max(g.zlo - 2, 0) / TILE_Z, tz1 = min(g.zhi + 2, g.nz - 1) / TILE_Z;
  const int tx0 = max(g.xlo - 2, 0) / TILE_X, tx1 = min(g.xhi + 2, g.nx - 1) / TILE_X;
  const int ntz = tz1 - tz0 + 1, ntx = tx1 - tx0 + 1;
  if (a.acc_group < 1) a.acc_group = 1;
  if (a.acc_group > a.batch) a.acc_group = a.batch;
  const int nunits = ((a.batch + a.acc_group - 1) / a.acc_group) * ntz * ntx;
  const int blocks = nunits < sm_count() * CTAS_PER_SM ? nunits : sm_count() * CTAS_PER_SM;
  const bool lean = reverse_is_lean(g, a.batch);
  if (a.acc_group > 1) {
    if (lean) launch_step(rev_image_kernel<true, true>, blocks, NCOMPUTE, rev_smem<true>() + racc_bytes<true>(), s, a, tz0, tx0, ntz, ntz * ntx);
    else launch_step(rev_image_kernel<false, true>, blocks, NCOMPUTE, rev_smem<false>() + racc_bytes<false>(), s, a, tz0, tx0, ntz, ntz * ntx);
  } else {
    if (lean) launch_step(rev_image_kernel<true, false>, blocks, NCOMPUTE, rev_smem<true>(), s, a, tz0, tx0, ntz, ntz * ntx);
    else launch_step(rev_image_kernel<false, false>, blocks, NCOMPUTE, rev_smem<false>(), s, a, tz0, tx0, ntz, ntz * ntx);
  }
}

// working set of one launch = forward + adjoint fields and accumulators of every box cell of the batch; beyond a few
// L2 capacities the kernel is bound by its streams and wants the larger L1 (see rev_smem)
bool reverse_is_lean(const Grid &g, int batch) {
  const double working_set = 100.0 * batch * (double)(g.zhi - g.zlo + 1) * (g.xhi - g.xlo + 1);
  const int force = g_rev_lean_force.load(std::memory_order_relaxed);   // A/B switch (fwi_b200_set_option)
  return force >= 0 ? force == 1 : (REV_LEAN_AUTO && working_set > 512e6);
}

// Shots per accumulator slot of the reverse kernel (BwdArgs::acc_group).  Grouping pays where the kernel is bound by
// its HBM streams (the LEAN regime); where the fields stay in L2 (C2) the whole gradient does not change and a slot per
// shot is kept.  Per launch, C3 grid / C5 grid (8 shots), us:
//                                   C3 8 shots   C3 25 shots   C5 8 shots
//   one slot per shot                  450          1476          4584
//   groups, units dealt round-robin    357 (8)      1158 (9), 1222 (25)     3823 (4), 4064 (8)
//   groups, units claimed dynamically  351 (8)      1077 (9), 1049 (25)     3730 (4), 3476 (8)
// Dealt statically, the CTAs drift apart over the rounds of a launch, each at its own shot of its own tile, and the
// accesses scatter over HBM pages -- the longer the launch, the more (C5: 139 rounds); claimed from a counter, the
// units in flight are always ~148 consecutive tiles.  Dynamic: the whole batch is one group (at most 32 shots);
// static: groups of at most 12.  Either way more groups when a launch would have fewer than 6 rounds of units.
int reverse_acc_group(const Grid &g, int batch, bool dyn) {
  const int force = g_acc_group_force.load(std::memory_order_relaxed);
  if (force >= 1) return force < batch ? force : batch;
  if (!reverse_is_lean(g, batch)) return 1;
  const int tz0 = max(g.zlo - 2, 0) / TILE_Z, tz1 = min(g.zhi + 2, g.nz - 1) / TILE_Z;
  const int tx0 = max(g.xlo - 2, 0) / TILE_X, tx1 = min(g.xhi + 2, g.nx - 1) / TILE_X;
  const long long ntiles = (long long)(tz1 - tz0 + 1) * (tx1 - tx0 + 1);
  const long long want = 6LL * sm_count() * CTAS_PER_SM;
  const int gmax = dyn ? 32 : 12;
  long long ngrp = (batch + gmax - 1) / gmax;
  while (ntiles * ngrp < want && ngrp < batch) ngrp++;
  return (int)((batch + ngrp - 1) / ngrp);
}

}  // namespace fwi
