// Backward (reverse-time) half of the gradient for sm_100a, same persistent TMA-fed structure as fwi_forward.cu.
//
// rev_image_kernel: wavefield reconstruction it+1 -> it inside the inner box + frame restore + imaging condition
//   replaces el_velocity(isFor=false) / to_bnd x2 / add_source(isFor=false) / el_stress(isFor=false) / to_bnd x3
//   (reference: deps/CustomOps/FWI/Src/libCUFD.cu:380-403, el_velocity.cu:84-117, el_stress.cu:90-131,
//    utilities.cu:394-424,538-551)
//   The imaging condition only ACCUMULATES per-cell source terms (5 planes per concurrent shot: lambda, mu-direct,
//   mu-spray amplitude S, rho-a, rho-b); the reference's 4-point atomic "spray" (el_stress.cu:113-124,
//   el_velocity.cu:101-110) is linear in those and is applied once, as a deterministic gather, by finalize_kernel.
//
// One CTA of 16 warps per SM loops over (shot, tile) items of the tile range that covers the inner box + frame ring.
// The producer lane streams, three items ahead, the stress triple of time it+1 with halo 8 / 4 (72 x 36) and the
// velocity pair with halo 4 / 2 (64 x 32) through a 3-stage TMA ring.  Every thread owns one float4 quad of the
// 64 x 32 region: it rewinds the velocities of its quad (all threads; the result goes to a double-buffered shared
// tile), then -- owner threads -- rewinds the stresses of the same quad from the neighbours' rewound velocities.
#include "fwi_device.cuh"
#include "fwi_host.hpp"

namespace fwi {
using namespace dev;

namespace {

constexpr int NS = 3;
constexpr int RW_BYTES = 3 * WCOLS * VPITCH * 4;   // stress triple, rows z0-8.., columns x0-4..
constexpr int RV_BYTES = 2 * SCOLS * SPITCH * 4;   // velocity pair, rows z0-4.., columns x0-2..
constexpr int RSTAGE_BYTES = RW_BYTES + RV_BYTES;
constexpr int SV_BYTES = 2 * SCOLS * SPITCH * 4;   // rewound velocities
constexpr size_t REV_SMEM = (size_t)NS * RSTAGE_BYTES + 2 * SV_BYTES + (NS + 1) * sizeof(TileDesc) + NS * 8 + 128;
static_assert(RW_BYTES % 128 == 0 && RV_BYTES % 128 == 0, "TMA destination alignment");

__global__ void __launch_bounds__(NCOMPUTE, 1)
rev_image_kernel(const __grid_constant__ BwdArgs a, int tz_first, int tx_first, int ntz, int ntiles) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  float *s_v_base = reinterpret_cast<float *>(base + NS * RSTAGE_BYTES);                            // [2][2][SCOLS][SPITCH]
  TileDesc *sdesc = reinterpret_cast<TileDesc *>(base + NS * RSTAGE_BYTES + 2 * SV_BYTES);          // [NS + 1]
  uint64_t *full = reinterpret_cast<uint64_t *>(base + NS * RSTAGE_BYTES + 2 * SV_BYTES + (NS + 1) * sizeof(TileDesc));

  const Grid &g = a.g;
  const int tid = threadIdx.x;
  const int nitems = a.batch * ntiles;
  const int stride = gridDim.x;
  const int fin = a.cur_f ? S_FB : S_FA, fout = a.cur_f ? S_FA : S_FB;
  const int ain = a.cur_a ? S_AB : S_AA;
  const int P = g.P;
  const long long pl = g.plane;

  if (tid == 0) {
    for (int s = 0; s < NS; s++) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto produce = [&](int item, int stage, int ds) {
    const int shot = item / ntiles, t = item - shot * ntiles;
    const int z0 = (tz_first + t % ntz) * TILE_Z, x0 = (tx_first + t / ntz) * TILE_X;
    const int sz = a.st.src_z[shot], sx = a.st.src_x[shot];
    TileDesc d;
    d.soff = (long long)shot * S_COUNT * pl + (long long)x0 * P + z0;
    d.moff = x0 * P + z0;
    d.z0 = z0; d.x0 = x0; d.shot = shot; d.tile = t; d.sz = sz; d.sx = sx; d.r0 = d.r1 = 0;
    int fl = 0;
    // every tile whose owner cells or their +-4 halo can touch the ring (the old launch's frame_tile test)
    if (!(z0 - 4 > g.zlo + 2 && z0 + TILE_Z + 3 < g.zhi - 2 && x0 - 2 > g.xlo + 2 && x0 + TILE_X + 1 < g.xhi - 2)) fl |= TF_FRAME;
    if (sz >= z0 && sz < z0 + TILE_Z && sx >= x0 && sx < x0 + TILE_X) fl |= TF_SRC;
    d.flags = fl;
    d.pad[0] = d.pad[1] = d.pad[2] = d.pad[3] = 0;
    sdesc[ds] = d;
    unsigned char *sb = base + stage * RSTAGE_BYTES;
    const int p0 = shot * S_COUNT + fin;
    mbar_arrive_expect_tx(&full[stage], RSTAGE_BYTES);
    tma_load_3d(sb, &a.tm.sw, z0 - 8, x0 - 4 + XM, p0 + F_SZZ, &full[stage]);
    tma_load_3d(sb + RW_BYTES, &a.tm.vn, z0 - 4, x0 - 2 + XM, p0 + F_VZ, &full[stage]);
  };
  if (tid == PRODUCER_TID)
    for (int s = 0; s < NS; s++)
      if (blockIdx.x + s * stride < nitems) produce(blockIdx.x + s * stride, s, s);

  const float dt = g.dt;
  const float kz1 = C1 * g.rdz, kz2 = C2 * g.rdz, kx1 = C1 * g.rdx, kx2 = C2 * g.rdx;
  const float half_rdt = 0.5f / dt;            // 0.5 byc^2 dt      = (0.5 / dt) (byc dt)^2
  const float q_rdt = 250000.0f / dt;          // 1e6 mu_bar^2 dt/4 = (250000 / dt) (mu_bar dt)^2
  const float dt6 = dt * 1e6f;
  const int q = tid & 15, c = tid >> 4;
  const bool inner = q >= 1 && q <= TILE_Z / 4 && c >= 2 && c < TILE_X + 2;
  const int sj = c * SPITCH + 4 * q;
  const int gx_max = g.nx + XM - 1;

  int stage = 0, phase = 0, nb = 0, ds = 0;
  for (int item = blockIdx.x; item < nitems; item += stride) {
    mbar_wait(&full[stage], phase);
    const TileDesc d = sdesc[ds];
    const int gz = d.z0 - 4 + 4 * q, gx = d.x0 - 2 + c;
    const bool inb = (unsigned)gx < (unsigned)g.nx && (unsigned)gz < (unsigned)g.nz;
    const bool owner = inner && inb;
    const long long toff = d.soff + ((long long)(c - 2) * P + 4 * q - 4);
    float *sq = a.state + g.origin + toff;                                       // + slot * pl
    float *acc = a.gacc + g.origin + (long long)d.shot * (G_COUNT - S_COUNT) * pl + toff;   // shot * G_COUNT * pl + cell
    const float *mq = a.m.ldt + ((long long)min(gx, gx_max) * P + gz);
    const bool colbox = gx >= g.xlo && gx <= g.xhi;
    bool bx[4];   // cell inside the inner box (reconstruction / imaging region)
#pragma unroll
    for (int kk = 0; kk < 4; kk++) bx[kk] = colbox && (unsigned)(gz + kk - g.zlo) <= (unsigned)(g.zhi - g.zlo);
    const bool in_rect = gx >= g.xlo - 2 && gx <= g.xhi + 2 && gz + 3 >= g.zlo - 2 && gz <= g.zhi + 2;
    const bool frame_tile = d.flags & TF_FRAME;
    const float *frm = a.frames + ((long long)d.shot * g.nSteps + a.it) * 5 * g.f_len;
    const FrameCol fc(g, gx);

    // global operands of the velocity half: buoyancies, adjoint velocities and the rho accumulators of the quad
    const F4 byadt = ld4(mq + 3 * pl), bybdt = ld4(mq + 4 * pl);
    F4 vza = zero4(), vxa = zero4(), accA = zero4(), accB = zero4();
    if (owner) {
      vza = ld4(sq + (ain + F_VZ) * pl);
      vxa = ld4(sq + (ain + F_VX) * pl);
      accA = ld4(acc + G_RHO_A * pl);
      accB = ld4(acc + G_RHO_B * pl);
    }

    const unsigned char *sb = base + stage * RSTAGE_BYTES;
    const float *sw = reinterpret_cast<const float *>(sb);              // [3][WCOLS][VPITCH]: szz sxx sxz of time it+1
    const float *sv = reinterpret_cast<const float *>(sb + RW_BYTES);   // [2][SCOLS][SPITCH]: vz vx of time it+1
    float *s_v = s_v_base + nb * (SV_BYTES / 4);

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
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      if (bx[kk]) {
        vz.v[kk] = fmaf(-ea[kk], byadt.v[kk], vz.v[kk]);
        vx.v[kk] = fmaf(-eb[kk], bybdt.v[kk], vx.v[kk]);
        // g = -v_adj (d sigma) dt * (-byc^2 / 2)     (el_velocity.cu:101-104)
        accA.v[kk] += (vza.v[kk] * ea[kk]) * (half_rdt * byadt.v[kk] * byadt.v[kk]);
        accB.v[kk] += (vxa.v[kk] * eb[kk]) * (half_rdt * bybdt.v[kk] * bybdt.v[kk]);
      }
    }
    if (owner) {
      st4(acc + G_RHO_A * pl, accA);
      st4(acc + G_RHO_B * pl, accB);
    }
    if (frame_tile && in_rect) {  // to_bnd(v): exact values of time `it` on the 5-cell ring (libCUFD.cu:388)
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const int fidx = fc.idx(gz + kk);
        if (fidx >= 0) {
          vz.v[kk] = frm[F_VZ * g.f_len + fidx];
          vx.v[kk] = frm[F_VX * g.f_len + fidx];
        }
      }
    }
    st4(s_v + sj, vz);
    st4(s_v + SCOLS * SPITCH + sj, vx);
    float *fo = sq + fout * pl;
    const bool wr = owner && in_rect;
    if (wr) {
      st4(fo + F_VZ * pl, vz);
      st4(fo + F_VX * pl, vx);
    }
    // global operands of the stress half, requested before the barrier
    F4 ldt, l2mdt, amudt, za, xa, xza, gl, gm, gs;
    if (wr) {
      ldt = ld4(mq); l2mdt = ld4(mq + pl); amudt = ld4(mq + 2 * pl);
      za = ld4(sq + (ain + F_SZZ) * pl); xa = ld4(sq + (ain + F_SXX) * pl); xza = ld4(sq + (ain + F_SXZ) * pl);
      gl = ld4(acc + G_LAM * pl); gm = ld4(acc + G_MU * pl); gs = ld4(acc + G_MUS * pl);
    }
    __syncthreads();  // s_v is complete; nobody reads ring slot `stage` any more
    if (tid == PRODUCER_TID && item + NS * stride < nitems) produce(item + NS * stride, stage, ds == 0 ? NS : ds - 1);

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
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          if (bx[kk]) {
            szz.v[kk] = fmaf(-l2mdt.v[kk], dvz_dz[kk], fmaf(-ldt.v[kk], dvx_dx[kk], szz.v[kk]));
            sxx.v[kk] = fmaf(-l2mdt.v[kk], dvx_dx[kk], fmaf(-ldt.v[kk], dvz_dz[kk], sxx.v[kk]));
            const float e = dvx_dz[kk] + dvz_dx[kk];
            sxz.v[kk] = fmaf(-amudt.v[kk], e, sxz.v[kk]);
            // el_stress.cu:109-116
            gl.v[kk] += -(za.v[kk] + xa.v[kk]) * (dvz_dz[kk] + dvx_dx[kk]) * dt6;
            gm.v[kk] += (-2.0f * za.v[kk] * dvz_dz[kk] - 2.0f * xa.v[kk] * dvx_dx[kk]) * dt6;
            //  s = -sxz_adj (exz + ezx) dt mu_bar / sum(1/mu) 1e6, mu_bar / sum(1/mu) == mu_bar^2 / 4; zero where mu_bar == 0
            gs.v[kk] += -xza.v[kk] * e * (q_rdt * amudt.v[kk] * amudt.v[kk]);
          }
        }
        st4(acc + G_LAM * pl, gl);
        st4(acc + G_MU * pl, gm);
        st4(acc + G_MUS * pl, gs);
      }
      if (frame_tile) {  // to_bnd(sigma) (libCUFD.cu:403)
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int fidx = fc.idx(gz + kk);
          if (fidx >= 0) {
            szz.v[kk] = frm[F_SZZ * g.f_len + fidx];
            sxx.v[kk] = frm[F_SXX * g.f_len + fidx];
            sxz.v[kk] = frm[F_SXZ * g.f_len + fidx];
          }
        }
      }
      st4(fo + F_SZZ * pl, szz);
      st4(fo + F_SXX * pl, sxx);
      st4(fo + F_SXZ * pl, sxz);
    }
    nb ^= 1;
    if (++ds == NS + 1) ds = 0;
    if (++stage == NS) { stage = 0; phase ^= 1; }
  }
}

}  // namespace

size_t reverse_smem_bytes() { return REV_SMEM; }

void configure_backward_kernels() {
  cudaFuncSetAttribute(rev_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REV_SMEM);
}

void launch_reverse_imaging(const BwdArgs &a, cudaStream_t s) {
  const Grid &g = a.g;
  const int tz0 = max(g.zlo - 2, 0) / TILE_Z, tz1 = min(g.zhi + 2, g.nz - 1) / TILE_Z;
  const int tx0 = max(g.xlo - 2, 0) / TILE_X, tx1 = min(g.xhi + 2, g.nx - 1) / TILE_X;
  const int ntz = tz1 - tz0 + 1, ntx = tx1 - tx0 + 1;
  const int nitems = a.batch * ntz * ntx;
  const int blocks = nitems < sm_count() ? nitems : sm_count();
  rev_image_kernel<<<blocks, NCOMPUTE, REV_SMEM, s>>>(a, tz0, tx0, ntz, ntz * ntx);
}

}  // namespace fwi
