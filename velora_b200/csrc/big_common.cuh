// Shared constants and helpers of the 512-neuron class kernels (liquid_big.cu: forward, liquid_big_bwd.cu: backward).
#pragma once
#include <cuda_bf16.h>
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace vb {
namespace big {
using namespace tcx;
template <int V> using IC = std::integral_constant<int, V>;

constexpr int I = 8, NI = 308, NC = 204;
constexpr int A_CHUNKS = 39;                 // a: 308 features + 4 zeros
constexpr int H_CHUNK0 = 39;                 // h starts at feature 312
constexpr int H_CHUNKS = 26;                 // 204 h + the constant-1 feature + 3 zeros
constexpr int KC = 66;                       // 528 features, 33 K-steps
constexpr int NPASS = 4, PN = 52, PROWS = 4 * PN;     // 52 neurons = 208 accumulator columns per pass
constexpr int STAGE_CHUNKS = 6, NSTAGE_PER_PASS = KC / STAGE_CHUNKS;      // 11 stages of K = 48
constexpr int STAGE_BYTES = STAGE_CHUNKS * PROWS * 16;                    // 19968
constexpr int RING = 3;
constexpr int A_STAGES = 6;                  // stages 0..5 (chunks 0..35) contract over the a-part of z~ only
// Element-wise threads per sample: warps w, w + 4, w + 8, ... address the same TMEM lanes and split a row's features.  Measured in round 2
// on the training forward at 32768 x 64 (profiles/r02/README.md): three threads 4.41 ms, four 4.29 ms, six 4.52 ms (24 warps at 72
// registers execute 27 % more instructions - per-thread overheads - at 50 % instead of 45 % issue utilisation).
#ifndef VB_BIG_EWT
#define VB_BIG_EWT 4
#endif
constexpr int EWT = VB_BIG_EWT;
constexpr int EW = TM * EWT;
constexpr int NTHREADS = EW + 64;            // + producer warp + MMA warp
// contiguous split of `total` items over `parts` owners
__host__ __device__ constexpr int part_cnt(int total, int parts, int p) { return total / parts + (p < total % parts ? 1 : 0); }
__host__ __device__ constexpr int part_off(int total, int parts, int p) { return p * (total / parts) + (p < total % parts ? p : total % parts); }
constexpr int AMAX = (A_CHUNKS + EWT - 1) / EWT;      // a-chunks held in registers per thread

constexpr int S_Z = 0;
constexpr int S_STAGE = S_Z + KC * PLANE;
constexpr int S_WIN = S_STAGE + RING * STAGE_BYTES;            // fp32 [I][320]
constexpr int S_BIN = S_WIN + I * 320 * 4;                     // fp32 [320]
constexpr int S_WOUT = S_BIN + 320 * 4;                        // fp32 [208]
constexpr int S_YS = S_WOUT + 208 * 4;                         // fp32 [128]
constexpr int S_BAR = S_YS + (EWT - 1) * 128 * 4;
constexpr int SMEM_BYTES = S_BAR + 128;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// h_traj in the kernel's own tile-interleaved order: [T][tile][NC / 4][128 rows][4 floats]: a warp reads / writes 512 contiguous bytes
__host__ __device__ __forceinline__ size_t h_index(int t, int64_t n_tiles, int64_t tile, int row, int j) {
    return ((((size_t)t * n_tiles + tile) * (NC / 4) + j / 4) * TM + row) * 4 + (j & 3);
}
__device__ __forceinline__ void ew_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EW) : "memory"); }



// ---- activations saved by the training forward for the backward (vb_liquid_bf16_fwd with a saved block), R = T * n_tiles * 128 rows
//      in the kernel's own row order r = (t * n_tiles + tile) * 128 + row:
//   G  [t][tile][quad (51)][row][16 bf16]  per command neuron (A, Bc, Cc, mish'(c)): ds = dh' * (A, Bc, Cc), dc = dm * mish'(c)
//   M  [t][tile][quad (51)][row][4 bf16]   m = mish(c) (for dW_out)
//   Z  [r][528 bf16] row-major             z~ = [a (308) | 0 (4) | h_{t-1} (204) | 1 | 0 (11)]: the wgrad GEMM's second operand
//   U  [r][312 bf16] row-major             mish'(W_in x + b_in) (zero padded): du = dz_a * U is then one streaming pass
constexpr int NQUAD = NC / 4;                         // 51
constexpr int ZW = KC * 8;                            // 528
constexpr int DSW = 4 * NC;                           // 816 = ds features per row, feature 4 j + hs
constexpr int AW = A_CHUNKS * 8;                      // 312 = padded a-part
constexpr size_t G_ROW_BYTES = (size_t)NQUAD * 32, M_ROW_BYTES = (size_t)NQUAD * 8, Z_ROW_BYTES = (size_t)ZW * 2, U_ROW_BYTES = (size_t)AW * 2;
__host__ __device__ inline int64_t saved_rows(int64_t B, int T) { return (int64_t)T * ((B + TM - 1) / TM) * TM; }
__host__ __device__ inline size_t saved_g_off(int64_t) { return 0; }
__host__ __device__ inline size_t saved_m_off(int64_t R) { return (size_t)R * G_ROW_BYTES; }
__host__ __device__ inline size_t saved_z_off(int64_t R) { return (size_t)R * (G_ROW_BYTES + M_ROW_BYTES); }
__host__ __device__ inline size_t saved_u_off(int64_t R) { return (size_t)R * (G_ROW_BYTES + M_ROW_BYTES + Z_ROW_BYTES); }
__host__ __device__ inline size_t saved_total_bytes(int64_t R) { return (size_t)R * (G_ROW_BYTES + M_ROW_BYTES + Z_ROW_BYTES + U_ROW_BYTES); }

}  // namespace big
}  // namespace vb
