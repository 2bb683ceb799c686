// Device-side building blocks shared by the sm_100a step kernels: float4 "quads", the 4th-order staggered
// differences on quads, and the TMA (cp.async.bulk.tensor) + mbarrier primitives the persistent kernels use to
// stream halo tiles from HBM into shared memory while the previous tile is being computed.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <utility>

#include <string>

#include "fwi_host.hpp"
#include "fwi_kernels.cuh"

#ifndef FWI_PDL
#define FWI_PDL 1   // programmatic dependent launch: a step kernel's prologue overlaps the tail of the previous launch
#endif
#ifndef FWI_STREAM_LD
#define FWI_STREAM_LD 0   // measured: L1::no_allocate loads are slower here (rev C3 447 -> 590 us)
#endif

#ifndef FWI_F64_UPDATE
#define FWI_F64_UPDATE 8   // 0: float increments everywhere; 1: stress / adjoint-velocity increments summed in double everywhere (one rounding
                           // per update, like the reference's (lambda + 2.0 mu) expressions); R > 1: only in the forward kernel and only for
                           // quads within R cells of the shot's source.  (Measured and dropped: the reference's arithmetic to the letter for
                           // those quads -- IEEE division by dz, unscaled coefficients: C2 grad_stf 1.0e-3 -> 4.9e-4, forward kernel +34 %.)
#endif

#ifndef FWI_F64_ADJ
#define FWI_F64_ADJ 0
#endif

namespace fwi {
namespace dev {

constexpr float C1 = 1.125f;                     // 9/8   (el_stress.cu:45)
constexpr float C2 = (float)(1.0 / 24.0);        // 1/24  (el_stress.cu:46)
constexpr float SRC_SCALE = 2250000.0f;          // pow(1500,2)  utilities.cu:528

struct F4 {
  float v[4];
};
__device__ __forceinline__ F4 ld4(const float *p) {
  const float4 t = *reinterpret_cast<const float4 *>(p);
  return F4{{t.x, t.y, t.z, t.w}};
}
// streaming 16-byte load: the line is not kept in L1 (per-shot operands that are read once per launch), which leaves the
// L1 data array to the loads that do have reuse (coefficients, CPML memory read with x-neighbours)
__device__ __forceinline__ F4 ld4s(const float *p) {
#if FWI_STREAM_LD
  F4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
  return r;
#else
  return ld4(p);
#endif
}
__device__ __forceinline__ void st4(float *p, const F4 &a) {
  *reinterpret_cast<float4 *>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
__device__ __forceinline__ F4 zero4() { return F4{{0.f, 0.f, 0.f, 0.f}}; }

// 7 consecutive samples w[0..6] = f[z-2 .. z+4]  ->  D-z at the 4 cells z..z+3   (el_stress.cu:54)
//   (c1 (f[z]-f[z-1]) - c2 (f[z+1]-f[z-2])) / h, with c1/h and c2/h folded into k1, k2
__device__ __forceinline__ void dz_minus4(const F4 &A, const F4 &B, const F4 &C, float k1, float k2, float *out) {
  const float w[7] = {A.v[2], A.v[3], B.v[0], B.v[1], B.v[2], B.v[3], C.v[0]};
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = k1 * (w[k + 2] - w[k + 1]) - k2 * (w[k + 3] - w[k]);
}
// samples u[0..6] = f[z-1 .. z+5]  ->  D+z at the 4 cells   (el_stress.cu:71)
__device__ __forceinline__ void dz_plus4(const F4 &A, const F4 &B, const F4 &C, float k1, float k2, float *out) {
  const float u[7] = {A.v[3], B.v[0], B.v[1], B.v[2], B.v[3], C.v[0], C.v[1]};
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = k1 * (u[k + 2] - u[k + 1]) - k2 * (u[k + 3] - u[k]);
}
// columns x-2, x-1, x, x+1 -> D-x ;  columns x-1, x, x+1, x+2 -> D+x   (same expression shape)
__device__ __forceinline__ void dx4(const F4 &m2, const F4 &m1, const F4 &c0, const F4 &p1, float k1, float k2, float *out) {
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = k1 * (c0.v[k] - m1.v[k]) - k2 * (p1.v[k] - m2.v[k]);
}

// ---- shared-memory addresses, mbarrier, TMA ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        "  .reg .pred p;\n"
        "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "  selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// one 3-D box (z, x, plane) HBM -> shared memory; completion is signalled on `bar` in bytes.
// Out-of-range coordinates (negative z in the first tile row, x beyond the margins) are zero-filled by the TMA unit.
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// one 3-D box HBM -> L2 only (no shared-memory destination, no completion to wait for): used two items ahead for the
// per-shot operands that the compute threads fetch with direct 16-byte loads (CPML memory, adjoint fields, imaging
// accumulators), so that those loads find their sectors in L2 instead of paying the HBM latency inside the item.
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// Programmatic dependent launch (the step kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization):
// a CTA of the next launch becomes resident as soon as an SM is free and runs its prologue -- barrier init, tile
// descriptors, static coefficient loads -- while the previous launch drains; nothing written by the previous launch
// may be touched before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() {
#if FWI_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if FWI_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// ---- tile geometry shared by the persistent step kernels ----------------------------------------
constexpr int SQ = (TILE_Z + 8) / 4;     // 16 quads per column of the "stress region": rows z0-4 .. z0+TILE_Z+3
constexpr int SCOLS = TILE_X + 4;        // 32 columns of the stress region: x0-2 .. x0+TILE_X+1
constexpr int SPITCH = TILE_Z + 8;       // 64
constexpr int VCOLS = TILE_X + 6;        // 34 columns with halo 3: x0-3 .. x0+TILE_X+2
constexpr int WCOLS = TILE_X + 8;        // 36 columns with halo 4: x0-4 .. x0+TILE_X+3
constexpr int VPITCH = TILE_Z + 16;      // 72: rows z0-8 .. z0+TILE_Z+7
constexpr int NCOMPUTE = SQ * SCOLS;     // 512 threads: one quad of the stress region each
constexpr int PRODUCER_TID = NCOMPUTE - 32;  // lane 0 of the warp holding columns 30, 31 (never owners)
static_assert(SQ == 16 && NCOMPUTE == NT_STEP, "one quad per thread, 16 quads per half-warp");

enum : int { TF_PML = 1, TF_FRAME = 2, TF_SRC = 4 };

struct __align__(16) TileDesc {  // built by the producer lane, read (broadcast) by every thread
  long long soff;   // element offset of (z0, x0) in this shot's slot-0 plane: shot * S_COUNT * plane + x0 * P + z0
  int moff;         // x0 * P + z0 (model planes)
  int flags;
  int z0, x0, shot, tile;
  int sz, sx;       // source cell
  int r0, r1;       // receiver range (CSR over tiles)
  int pad[4];
};
static_assert(sizeof(TileDesc) == 64, "descriptor size");

// Boundary frames (the reference's Bnd store, Boundary.cu:17-41, utilities.cu:361-424): per step and field the
// float4 quads that intersect the ring around the inner box -- the two cells outside it on every side plus f_in cells
// inside (f_in = 3: the reference's 5-deep ring; f_in = 0: only what the stencils of the box cells reach) -- laid out as
// [left 2 + f_in columns | right 2 + f_in columns | per middle column: f_ntq top quads, f_nbq bottom quads].  Returns the
// quad slot of the quad starting at row gz (a multiple of 4) in column gx, or -1.  Restoring a whole quad also restores
// a few cells next to the ring with their exact forward values, which is harmless.
__device__ __forceinline__ int frame_quad(const Grid &g, int gz, int gx) {
  if (gx < g.xlo - 2 || gx > g.xhi + 2 || gz < g.f_zq0 || gz > g.zhi + 2) return -1;
  const int nL = 2 + g.f_in;
  if (gx <= g.xlo - 1 + g.f_in) return (gx - (g.xlo - 2)) * g.f_nqB + ((gz - g.f_zq0) >> 2);
  if (gx >= g.xhi + 1 - g.f_in) return (nL + gx - (g.xhi + 1 - g.f_in)) * g.f_nqB + ((gz - g.f_zq0) >> 2);
  const int t = gz >> 2, mid = 2 * nL * g.f_nqB + (gx - (g.xlo + g.f_in)) * (g.f_ntq + g.f_nbq);
  if ((unsigned)(t - g.f_tq0) < (unsigned)g.f_ntq) return mid + (t - g.f_tq0);
  if ((unsigned)(t - g.f_bq0) < (unsigned)g.f_nbq) return mid + g.f_ntq + (t - g.f_bq0);
  return -1;
}
// 16-byte asynchronous copy global -> shared (LDGSTS): no register staging; complete after cp_async_wait_all()
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ bool tile_touches_frame(const Grid &g, int z0, int x0) {
  return !(z0 > g.zhi + 2 || z0 + TILE_Z - 1 < g.zlo - 2 || x0 > g.xhi + 2 || x0 + TILE_X - 1 < g.xlo - 2) &&
         !(z0 > g.zlo - 1 + g.f_in && z0 + TILE_Z - 1 < g.zhi + 1 - g.f_in && x0 > g.xlo - 1 + g.f_in &&
           x0 + TILE_X - 1 < g.xhi + 1 - g.f_in);
}

}  // namespace dev

int sm_count();  // fwi_forward.cu

// <<<blocks, threads, smem, stream>>> with the programmatic-stream-serialization attribute (see pdl_wait())
template <typename... KArgs, typename... Args>
inline void launch_step(void (*kernel)(KArgs...), int blocks, int threads, size_t smem, cudaStream_t s, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = FWI_PDL;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (e != cudaSuccess)   // surfaces at the launch that failed, not at the end of the run
    throw Error(FWI_B200_ERR_CUDA, std::string("CUDA: kernel launch failed: ") + cudaGetErrorString(e));
}

}  // namespace fwi
