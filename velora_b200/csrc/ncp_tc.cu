// NCPNetwork (the SAC critics, velora/models/lnn/ncp.py:299-328) forward on the tensor cores, fp32-equivalent (bf16x3).
//
//   y = W3 mish(W2 mish(W1 x + b1) + b2) + b3          critic shapes: (5 | 4) -> 77 -> 51 -> (1 | 2)
//
// One CTA = 128 threads = 128 samples = the 128 TMEM lanes.  Layer 1 (K = 5) and layer 3 (N = 1 or 2) are a few FMAs per
// thread; layer 2 (77 x 51, 90 % of the MACs) is ONE accumulator: the three bf16 parts of a~1 = [mish(u1) | 1] are issued one
// after the other (K = 80 each) against the three weight parts stacked along N (3 x 64 = 192 rows), so column group pb holds
// a~1 W2_pb^T (+ bias through the constant-1 feature) and 15 MMAs of N = 192 replace 30 of N = 64.  The epilogue adds the
// three groups smallest first.  Used for large batches; the generic tile kernels (generic.cu) keep the small ones and the
// backward.
#include "common.cuh"
#include "tc_common.cuh"

namespace vb {
namespace ncp_tc {
using namespace tcx;

template <int I_, int NI_, int NC_, int O_>
struct Cfg {
    static constexpr int I = I_, NI = NI_, NC = NC_, O = O_;
    static constexpr int NIp = ceil4(NI), NCp = ceil4(NC), Op = ceil4(O);
    static constexpr int KC = (NI + 1 + 7) / 8;         // 8-feature chunks of a~1 = [a1 | 1 | 0..]
    static constexpr int KF = KC * 8;
    static constexpr int NP = (NC + 15) / 16 * 16;      // layer-2 outputs per part, padded for the MMA
    static constexpr int EC = (NC + 7) / 8;             // 8-column chunks the epilogue reads
    static_assert(KC % 2 == 0, "a~1 must be a whole number of K = 16 steps");
    static_assert(3 * NP <= 256, "stacked layer-2 width must fit one MMA");
    // packed block (NcpLayout, F section)
    static constexpr int W1_T = 0, B1 = W1_T + I * NIp, W2_T = B1 + NIp, B2 = W2_T + NI * NCp, W3 = B2 + NCp, B3 = W3 + O * NCp;
    // shared memory
    static constexpr int S_A = 0;                               // 3 parts x KC planes
    static constexpr int S_WB = S_A + 3 * KC * PLANE;           // [KC][3 NP rows][16 B]
    static constexpr int S_W1 = S_WB + KC * 3 * NP * 16;        // fp32 [I + 1][KF]: W1^T rows, then b1
    static constexpr int S_W3 = S_W1 + (I + 1) * KF * 4;        // fp32 [O][NP], then b3 [4]
    static constexpr int S_BAR = S_W3 + (O * NP + 4) * 4;
    static constexpr int BYTES = S_BAR + 64;
    static constexpr int TMEM_COLS = 256;
};

template <class C>
__global__ void __launch_bounds__(TM) ncp_fwd_tc_kernel(const float* __restrict__ packed, const float* __restrict__ x,
                                                        float* __restrict__ y, int64_t B) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* aP = smem + C::S_A;
    unsigned char* wB = smem + C::S_WB;
    float* w1s = reinterpret_cast<float*>(smem + C::S_W1);
    float* w3s = reinterpret_cast<float*>(smem + C::S_W3);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + C::S_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::S_BAR + 16);
    const int tid = threadIdx.x, warp = tid >> 5;

    // layer-2 operand: B[pb * NP + n][k] = part pb of (k < NI ? W2[n][k] : k == NI ? b2[n] : 0), zero rows for n >= NC
    for (int item = tid; item < C::NP * C::KC; item += TM) {
        const int n = item % C::NP, kc = item / C::NP;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = 8 * kc + e;
            v[e] = n >= C::NC ? 0.f : (k < C::NI ? __ldg(packed + C::W2_T + k * C::NCp + n) : (k == C::NI ? __ldg(packed + C::B2 + n) : 0.f));
        }
        uint4 part[3];
        split3(v[0], v[1], part[0].x, part[1].x, part[2].x);
        split3(v[2], v[3], part[0].y, part[1].y, part[2].y);
        split3(v[4], v[5], part[0].z, part[1].z, part[2].z);
        split3(v[6], v[7], part[0].w, part[1].w, part[2].w);
        uint4* q = reinterpret_cast<uint4*>(wB);
#pragma unroll
        for (int pb = 0; pb < 3; ++pb) q[kc * (3 * C::NP) + pb * C::NP + n] = part[pb];
    }
    for (int e = tid; e < (C::I + 1) * C::KF; e += TM) {
        const int i = e / C::KF, j = e % C::KF;
        w1s[e] = j < C::NI ? __ldg(packed + (i < C::I ? C::W1_T + i * C::NIp : C::B1) + j) : 0.f;
    }
    for (int e = tid; e < C::O * C::NP + 4; e += TM) {
        const int o = e / C::NP, j = e % C::NP;
        w3s[e] = e < C::O * C::NP ? (j < C::NC ? __ldg(packed + C::W3 + o * C::NCp + j) : 0.f)
                                  : (e - C::O * C::NP < C::O ? __ldg(packed + C::B3 + e - C::O * C::NP) : 0.f);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_addr = smem_u32(aP), w_addr = smem_u32(wB);
    uint32_t ph = 0;

    const int64_t n_tiles = (B + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile * TM + tid;
        const bool valid = b < B;
        float xr[C::I];
#pragma unroll
        for (int i = 0; i < C::I; ++i) xr[i] = valid ? __ldg(x + b * C::I + i) : 0.f;
        // ---------------- layer 1 + Mish on CUDA cores, eight neurons (one plane chunk) at a time
#pragma unroll 2
        for (int c = 0; c < C::KC; ++c) {
            float u[8];
            {
                const float4 b0 = *reinterpret_cast<const float4*>(w1s + C::I * C::KF + 8 * c);
                const float4 b1 = *reinterpret_cast<const float4*>(w1s + C::I * C::KF + 8 * c + 4);
                u[0] = b0.x; u[1] = b0.y; u[2] = b0.z; u[3] = b0.w; u[4] = b1.x; u[5] = b1.y; u[6] = b1.z; u[7] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < C::I; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(w1s + i * C::KF + 8 * c);
                const float4 w1 = *reinterpret_cast<const float4*>(w1s + i * C::KF + 8 * c + 4);
                u[0] = fmaf(xr[i], w0.x, u[0]); u[1] = fmaf(xr[i], w0.y, u[1]); u[2] = fmaf(xr[i], w0.z, u[2]); u[3] = fmaf(xr[i], w0.w, u[3]);
                u[4] = fmaf(xr[i], w1.x, u[4]); u[5] = fmaf(xr[i], w1.y, u[5]); u[6] = fmaf(xr[i], w1.z, u[6]); u[7] = fmaf(xr[i], w1.w, u[7]);
            }
            float a[8];
            {
                float u4[4] = {u[0], u[1], u[2], u[3]}, a4[4];
                mish4(u4, a4);
                a[0] = a4[0]; a[1] = a4[1]; a[2] = a4[2]; a[3] = a4[3];
                float v4[4] = {u[4], u[5], u[6], u[7]};
                mish4(v4, a4);
                a[4] = a4[0]; a[5] = a4[1]; a[6] = a4[2]; a[7] = a4[3];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {      // the constant-1 (bias) feature and the zero padding behind it
                const int k = 8 * c + e;
                if (k == C::NI) a[e] = 1.0f;
                else if (k > C::NI) a[e] = 0.f;
            }
            store_chunk3<C::KC>(aP, c, tid, a);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = make_idesc(128, 3 * C::NP, 0, 0);
#pragma unroll
            for (int pa = 2; pa >= 0; --pa)          // smallest part first: the accumulator truncates
#pragma unroll
                for (int ks = 0; ks < C::KC / 2; ++ks)
                    mma_bf16(tmem, make_desc(a_addr + (pa * C::KC + 2 * ks) * PLANE, PLANE, 128),
                             make_desc(w_addr + 2 * ks * (3 * C::NP * 16), 3 * C::NP * 16, 128), idesc, !(pa == 2 && ks == 0));
            mma_commit(bar);
        }
        mbar_wait(bar, ph);
        ph ^= 1;
        tc_fence_after();
        // ---------------- Mish + layer 3
        float yo[C::O];
#pragma unroll
        for (int o = 0; o < C::O; ++o) yo[o] = w3s[C::O * C::NP + o];
#pragma unroll
        for (int c = 0; c < C::EC; ++c) {
            float s8[8];
            {
                float hi[8], mid[8], lo[8];
                tmem_ld8(taddr + 8 * c, hi);
                tmem_ld8(taddr + C::NP + 8 * c, mid);
                tmem_ld8(taddr + 2 * C::NP + 8 * c, lo);
#pragma unroll
                for (int e = 0; e < 8; ++e) s8[e] = (lo[e] + mid[e]) + hi[e];
            }
            float a2[8];
            {
                float u4[4] = {s8[0], s8[1], s8[2], s8[3]}, a4[4];
                mish4(u4, a4);
                a2[0] = a4[0]; a2[1] = a4[1]; a2[2] = a4[2]; a2[3] = a4[3];
                float v4[4] = {s8[4], s8[5], s8[6], s8[7]};
                mish4(v4, a4);
                a2[4] = a4[0]; a2[5] = a4[1]; a2[6] = a4[2]; a2[7] = a4[3];
            }
#pragma unroll
            for (int o = 0; o < C::O; ++o) {
                const float4 w0 = *reinterpret_cast<const float4*>(w3s + o * C::NP + 8 * c);
                const float4 w1 = *reinterpret_cast<const float4*>(w3s + o * C::NP + 8 * c + 4);
                yo[o] = fmaf(a2[0], w0.x, fmaf(a2[1], w0.y, fmaf(a2[2], w0.z, fmaf(a2[3], w0.w, yo[o]))));
                yo[o] = fmaf(a2[4], w1.x, fmaf(a2[5], w1.y, fmaf(a2[6], w1.z, fmaf(a2[7], w1.w, yo[o]))));
            }
        }
        if (valid) {
#pragma unroll
            for (int o = 0; o < C::O; ++o) y[b * C::O + o] = yo[o];
        }
        tc_fence_before();      // this tile's TMEM reads are ordered before the next tile's MMAs (issued after the next barrier)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

#define VB_NCP_TC_SHAPES(X) X(5, 77, 51, 1) X(4, 77, 51, 2)

bool has_path(int I, int Ni, int Nc, int O) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return true;
    VB_NCP_TC_SHAPES(X)
#undef X
    return false;
}

template <class C>
static int launch_fwd(const float* packed, const float* x, float* y, int64_t B, cudaStream_t stream) {
    LaunchCfg cfg;
    int rc = cached_launch_cfg(ncp_fwd_tc_kernel<C>, TM, C::BYTES, &cfg);
    if (rc) return rc;
    int smem_sm = 0;
    int dev = 0;
    VB_CUDA(cudaGetDevice(&dev));
    VB_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    int per_sm = smem_sm / (C::BYTES + cfg.static_smem + 1024);      // the occupancy API answers 1 for kernels that allocate TMEM
    if (per_sm > 512 / C::TMEM_COLS) per_sm = 512 / C::TMEM_COLS;
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (B + TM - 1) / TM;
    const int64_t cap = (int64_t)device_sms() * per_sm;
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    ncp_fwd_tc_kernel<C><<<grid < 1 ? 1 : grid, TM, C::BYTES, stream>>>(packed, x, y, B);
    VB_LAUNCH_CHECK();
    return 0;
}

int fwd(const float* packed, const float* x, float* y, int64_t B, int I, int Ni, int Nc, int O, cudaStream_t stream) {
#define X(a, b, c, d) if (I == a && Ni == b && Nc == c && O == d) return launch_fwd<Cfg<a, b, c, d>>(packed, x, y, B, stream);
    VB_NCP_TC_SHAPES(X)
#undef X
    set_error("no tensor-core NCP kernel for (%d,%d,%d,%d)", I, Ni, Nc, O);
    return VB_ERR_BAD_DIMS;
}

}  // namespace ncp_tc
}  // namespace vb
