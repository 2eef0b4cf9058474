// tcgen05 / TMEM / mbarrier helpers and the bf16x3 operand split shared by the tensor-core kernels (liquid_tc.cu, ncp_tc.cu).
//
// Operand planes: every MMA operand lives in shared memory in the no-swizzle canonical layout
//   plane[8-feature chunk][row][16 B = 8 bf16]
// so that a thread that owns one row stores its chunk with one conflict-free STS.128, and the same bytes can be read K-major
// (contraction over features) or MN-major (contraction over rows).  fp32 values are split into three bf16 parts
// (x = p1 + p2 + p3 to 24 bits); products are accumulated smallest first because the TMEM accumulator truncates between MMAs.
#pragma once
#include "common.cuh"

namespace vb {
namespace tcx {

constexpr int TM = 128;                  // rows per tile = threads per CTA = TMEM lanes
constexpr int PLANE = TM * 16;           // bytes per plane of TM rows

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16 |
           (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32 | (uint64_t)1 << 46;
}
// c = f32, a = b = bf16
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// eight consecutive columns as four packed pairs (columns 2k, 2k + 1)
__device__ __forceinline__ void tmem_ld8_p(uint32_t taddr, f2 (&v)[4]) {
    asm volatile("{\n\t.reg .b32 r<8>;\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x8.b32 {r0,r1,r2,r3,r4,r5,r6,r7}, [%4];\n\t"
                 "tcgen05.wait::ld.sync.aligned;\n\t"
                 "mov.b64 %0, {r0, r1};\n\tmov.b64 %1, {r2, r3};\n\tmov.b64 %2, {r4, r5};\n\tmov.b64 %3, {r6, r7};\n\t}"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "r"(taddr) : "memory");
}
// 24 consecutive columns as 12 packed pairs: three x8 loads in flight, one wait
__device__ __forceinline__ void tmem_ld24_p(uint32_t taddr, f2 (&v)[12]) {
    asm volatile("{\n\t.reg .b32 r<24>;\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x8.b32 {r0,r1,r2,r3,r4,r5,r6,r7}, [%12];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x8.b32 {r8,r9,r10,r11,r12,r13,r14,r15}, [%13];\n\t"
                 "tcgen05.ld.sync.aligned.32x32b.x8.b32 {r16,r17,r18,r19,r20,r21,r22,r23}, [%14];\n\t"
                 "tcgen05.wait::ld.sync.aligned;\n\t"
                 "mov.b64 %0, {r0, r1};\n\tmov.b64 %1, {r2, r3};\n\tmov.b64 %2, {r4, r5};\n\tmov.b64 %3, {r6, r7};\n\t"
                 "mov.b64 %4, {r8, r9};\n\tmov.b64 %5, {r10, r11};\n\tmov.b64 %6, {r12, r13};\n\tmov.b64 %7, {r14, r15};\n\t"
                 "mov.b64 %8, {r16, r17};\n\tmov.b64 %9, {r18, r19};\n\tmov.b64 %10, {r20, r21};\n\tmov.b64 %11, {r22, r23};\n\t}"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]), "=l"(v[4]), "=l"(v[5]), "=l"(v[6]), "=l"(v[7]), "=l"(v[8]), "=l"(v[9]),
                   "=l"(v[10]), "=l"(v[11])
                 : "r"(taddr), "r"(taddr + 8), "r"(taddr + 16) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// x0, x1 -> three packed bf16x2 words (low half = x0): x = p1 + p2 + p3 to 24 bits
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(x1), "f"(x0));
    float r0 = x0 - __uint_as_float(p1 << 16), r1 = x1 - __uint_as_float(p1 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(r1), "f"(r0));
    float q0 = r0 - __uint_as_float(p2 << 16), q1 = r1 - __uint_as_float(p2 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(q1), "f"(q0));
}

// split3 of a packed pair: the two residual subtractions are one FADD2 each
__device__ __forceinline__ void split3_p(f2 x, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(hi(x)), "f"(lo(x)));
    f2 r = sub2(x, pk(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u)));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(hi(r)), "f"(lo(r)));
    f2 q = sub2(r, pk(__uint_as_float(p2 << 16), __uint_as_float(p2 & 0xffff0000u)));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(hi(q)), "f"(lo(q)));
}
// four pairs (eight consecutive features) -> one 16-byte chunk in each of the three part planes
template <int CPP>
__device__ __forceinline__ void store_chunk3_p(unsigned char* planes, int chunk, int row, f2 f0, f2 f1, f2 f2_, f2 f3) {
    uint4 a, b, c;
    split3_p(f0, a.x, b.x, c.x);
    split3_p(f1, a.y, b.y, c.y);
    split3_p(f2_, a.z, b.z, c.z);
    split3_p(f3, a.w, b.w, c.w);
    uint4* p = reinterpret_cast<uint4*>(planes);
    p[(0 * CPP + chunk) * TM + row] = a;
    p[(1 * CPP + chunk) * TM + row] = b;
    p[(2 * CPP + chunk) * TM + row] = c;
}
// two pairs (four consecutive features) -> the lower (half = 0) or upper (half = 1) 8 bytes of one chunk in each part plane
template <int CPP>
__device__ __forceinline__ void store_half_chunk3_p(unsigned char* planes, int chunk, int row, int half, f2 f0, f2 f1) {
    uint2 a, b, c;
    split3_p(f0, a.x, b.x, c.x);
    split3_p(f1, a.y, b.y, c.y);
    uint2* p = reinterpret_cast<uint2*>(planes);
    p[((0 * CPP + chunk) * TM + row) * 2 + half] = a;
    p[((1 * CPP + chunk) * TM + row) * 2 + half] = b;
    p[((2 * CPP + chunk) * TM + row) * 2 + half] = c;
}

// eight consecutive features of this thread's sample -> one 16-byte chunk in each of the three part planes
// (CPP = chunks per part: 3 for the z~ planes, 4 for the ds planes)
template <int CPP>
__device__ __forceinline__ void store_chunk3(unsigned char* planes, int chunk, int row, const float (&f)[8]) {
    uint4 a, b, c;
    split3(f[0], f[1], a.x, b.x, c.x);
    split3(f[2], f[3], a.y, b.y, c.y);
    split3(f[4], f[5], a.z, b.z, c.z);
    split3(f[6], f[7], a.w, b.w, c.w);
    uint4* p = reinterpret_cast<uint4*>(planes);
    p[(0 * CPP + chunk) * TM + row] = a;
    p[(1 * CPP + chunk) * TM + row] = b;
    p[(2 * CPP + chunk) * TM + row] = c;
}

}  // namespace tcx
}  // namespace vb
