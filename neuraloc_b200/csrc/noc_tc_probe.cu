// noc_tc_probe.cu — stand-alone validation of the tcgen05 building blocks the tensor-core rollout path uses:
// TMEM allocation, no-swizzle ("interleaved") shared-memory matrix descriptors in both majors, the kind::f16
// instruction descriptor, tcgen05.commit -> mbarrier, tcgen05.ld / tcgen05.st.  One CTA computes
// D[128 x N] = A[128 x K] * B, bf16 operands, fp32 accumulate in TMEM, where B is given either K-major ([N][K]) or
// MN-major (the same buffer read as its transpose).  Exposed as noc_tc_probe() for tests/test_gpu_tc_probe.py.
#include "noc_launch.cuh"
#include "noc_tc.cuh"

namespace noc {

// D = A * B^T with B [N][K] (b_mn_major = 0), or D = A * B with B [K][N] stored through the SAME il_off(k, n, N)
// layout and read MN-major (b_mn_major = 1).  A [128][K] row-major bf16 in global, D [128][N] fp32.
__global__ void __launch_bounds__(128) tc_probe_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int N, int K,
                                                        int b_mn_major, int roundtrip_tmem) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                         // 128 * K * 2 bytes
    unsigned char* sB = smem + 128 * K * 2;           // N * K * 2 bytes
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 128 * K; i += 128) {
        int r = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16*>(sA + il_off(r, k, K)) = A[i];
    }
    if (!b_mn_major) {
        for (int i = tid; i < N * K; i += 128) { int n = i / K, k = i % K; *reinterpret_cast<__nv_bfloat16*>(sB + il_off(n, k, K)) = B[i]; }
    } else {   // B given as [K][N]: stored as a [K rows][N] operand; the MMA reads it with N contiguous (MN-major)
        for (int i = tid; i < N * K; i += 128) { int k = i / N, n = i % N; *reinterpret_cast<__nv_bfloat16*>(sB + il_off(k, n, N)) = B[i]; }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) mbar_init(smem_u32(&mbar), 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // operand writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = tmem_base_s;

    if (tid == 0) {
        const unsigned idesc = umma_idesc_bf16(128, N, b_mn_major);
        for (int kb = 0; kb < K / 16; ++kb) {
            // A, K-major: LBO = distance between the two K-chunks (128 B), SBO = distance between 8-row groups
            unsigned long long ad = umma_desc(smem_u32(sA) + kb * 256, 128, (K >> 3) * 128);
            unsigned long long bd;
            if (!b_mn_major) bd = umma_desc(smem_u32(sB) + kb * 256, 128, (K >> 3) * 128);
            // MN-major: the operand is [K rows][N]; one MMA covers rows 16 kb .. 16 kb + 15 (two 8-row groups, LBO apart);
            // SBO = distance between 16-byte chunks along N (128 B)
            else bd = umma_desc(smem_u32(sB) + kb * 2 * (N >> 3) * 128, (N >> 3) * 128, 128);
            umma_bf16(tbase, ad, bd, idesc, kb > 0);
        }
        umma_commit(smem_u32(&mbar));
    }
    mbar_wait(smem_u32(&mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        const unsigned taddr = tbase + ((unsigned)(warp * 32) << 16) + (unsigned)c0;
        tmem_ld32(taddr, v);
        if (roundtrip_tmem) {                          // park the chunk in other TMEM columns and read it back
            tmem_st32(taddr + 128, v);
            float w[32];
            tmem_ld32(taddr + 128, w);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = w[i];
        }
        for (int i = 0; i < 32; ++i)
            if (c0 + i < N) D[(size_t)tid * N + c0 + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(256) : "memory");
}

}  // namespace noc

// Test hook (not in include/noc_b200.h: it is not part of the product ABI).  A, B: bf16 device buffers; D: fp32 [128][N].
extern "C" int noc_tc_probe(const void* A, const void* B, void* D, int32_t N, int32_t K, int32_t b_mn_major, int32_t roundtrip_tmem, void* stream) {
    using namespace noc;
    if (N % 16 || N < 16 || N > 128 || K % 16 || K < 16 || K > 256) return fail(NOC_ERR_ARG, "probe: bad N/K");
    size_t smem = (size_t)128 * K * 2 + (size_t)N * K * 2;
    NOC_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (float*)D, N, K, b_mn_major, roundtrip_tmem);
    count_launch();
    NOC_CUDA(cudaGetLastError());
    return NOC_OK;
}
