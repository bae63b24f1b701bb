// noc_tc_probe.cu — stand-alone validation of the tcgen05 building blocks the tensor-core rollout path uses:
// TMEM allocation, no-swizzle ("interleaved") shared-memory matrix descriptors in both majors, the kind::f16
// instruction descriptor, tcgen05.commit -> mbarrier, tcgen05.ld / tcgen05.st.  One CTA computes
// D[128 x N] = A[128 x K] * B, bf16 operands, fp32 accumulate in TMEM, where B is given either K-major ([N][K]) or
// MN-major (the same buffer read as its transpose).  Exposed as noc_tc_probe() for tests/test_gpu_tc_probe.py.
#include <cuda_bf16.h>

#include "noc_launch.cuh"

namespace noc {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr & 0x3FFFF) >> 4);            // start address, 16-byte units, bits [0,14)
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;   // leading byte offset, bits [16,30)
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;   // stride byte offset, bits [32,46)
    d |= 1ull << 46;                                               // descriptor version 1 (Blackwell)
    return d;                                                      // base offset 0, layout type 0 (no swizzle)
}

// kind::f16 instruction descriptor: bf16 x bf16 -> f32, A K-major (cute/arch/mma_sm100_desc.hpp: InstrDescriptor)
__device__ __forceinline__ unsigned umma_idesc_bf16(int M, int N, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)b_mn_major << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, int accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, int parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n"
                 :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(unsigned taddr, const float (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                    "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                    "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
                    "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
                    "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
                    "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
                    "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// byte offset of element (row, k) of a [rows x K] bf16 operand in the interleaved layout: 8-row x 16-byte core
// matrices; the K-chunks (8 elements) of one 8-row group are adjacent (128 B apart), row groups K/8 * 128 B apart
__device__ __forceinline__ int il_off(int row, int k, int K) {
    return ((row >> 3) * (K >> 3) + (k >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

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
