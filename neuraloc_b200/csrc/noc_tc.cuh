// noc_tc.cuh — thin inline-PTX wrappers over the Blackwell tensor-core path (tcgen05 / TMEM) used by the probe
// (noc_tc_probe.cu) and the tensor-core rollout kernels (noc_tc_rollout.cuh, noc_ts_rollout.cuh).  Descriptor encodings follow
// cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor); canonical no-swizzle layouts follow the comments of
// cute/atom/mma_traits_sm100.hpp (make_umma_desc):  K-major ((8,n),2):((1,SBO),LBO),  MN-major ((1,n),(8,k)):((X,SBO),(1,LBO))
// in 16-byte units.  Both were validated on a B200 against torch matmul (tests/test_gpu_tc_probe.py).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>

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
// one lane of a converged warp (the pattern the compiler recognises as "single thread": no per-lane replay of the MMAs)
__device__ __forceinline__ unsigned elect_one_sync() {
    unsigned pred = 0, laneid = 0;
    asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}\n"
                 : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred;
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// issue only: the registers are valid after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_issue(unsigned taddr, unsigned (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    tmem_ld32_issue(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// v = [ta] + [tb] (two accumulators summed in round-to-nearest fp32), one wait for both loads
__device__ __forceinline__ void tmem_ld32_sum(unsigned ta, unsigned tb, float (&v)[32]) {
    unsigned r[32], q[32];
    tmem_ld32_issue(ta, r);
    tmem_ld32_issue(tb, q);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(q[i]);
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


__device__ __forceinline__ void tmem_ld16_issue(unsigned taddr, unsigned (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    tmem_ld16_issue(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16_bits(unsigned taddr, const unsigned (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32_bits(unsigned taddr, const unsigned (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
                    "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
                    "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// one warp; writes the base address to smem.  `last` = no further allocation by this CTA (gives up the permit)
__device__ __forceinline__ void tmem_alloc(unsigned smem_dst, int ncols, bool last = true) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    if (last) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, int ncols) {       // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes (operands written by threads) -> visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) primitives used by the streamed swarm kernel (noc_ts_rollout.cuh).  Two CTAs of a cluster
// issue ONE MMA of M = 128 (64 rows from each CTA) or M = 256; B's N rows are split between the two CTAs' shared
// memories; each CTA's TMEM holds its own rows of D (M = 128: lanes 0-63 hold columns [0, N/2), lanes 64-127 hold
// [N/2, N) -- the "2x2" atom of cute/atom/mma_traits_sm100.hpp: tmem_frg_2sm).  Barrier protocol after
// cutlass/pipeline/sm100_pipeline.hpp and cutlass/arch/barrier.h (umma_arrive_multicast_2x1SM, umma_arrive_2x1SM_sm0).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_count_x() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ unsigned mapa_shared(unsigned saddr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void fence_mbar_init_cluster() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Arrival on a barrier of another CTA of the cluster (cutlass/arch/barrier.h: ClusterBarrier::arrive(cta_id)).  Default
// semantics (release, CTA scope): everything handed over through these barriers is read by the tensor core / TMA (async proxy,
// after fence.proxy.async), never by another CTA's generic loads, so no cluster-scope fence -- which would cost a MEMBAR + ERRBAR
// per arrival and an L1 invalidation (CCTL.IVALL) per wait: 32 % of all stall samples in the first profile of the swarm kernel.
__device__ __forceinline__ void mbar_arrive_cluster(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must not hang the GPU (a hung box is a lost lease) -- after ~4 s of spinning the kernel
// reports which barrier it was waiting on and traps.  The clock is read once per 64 failed polls.
static __device__ __noinline__ void mbar_timeout(unsigned mbar, int parity, int tag) {
    printf("[noc] mbarrier wait timed out: block %d thread %d tag %d addr %u parity %d\n", (int)blockIdx.x, (int)threadIdx.x, tag, mbar, parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned mbar, int parity, int tag) {
    unsigned done = 0;
    long long t0 = 0;
    for (unsigned it = 0; ; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (done) break;
        if ((it & 63) == 63) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000ll) mbar_timeout(mbar, parity, tag);
        }
    }
}
// kind::f16 instruction descriptor, fp16 x fp16 -> f32, both operands K-major
__device__ __forceinline__ unsigned umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma2_f16(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, int accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all MMAs issued so far by this thread -> one arrival on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit_mc(unsigned mbar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" :: "r"(mbar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(unsigned smem_dst, int ncols) {       // the same warp index in both CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(unsigned taddr, int ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// TMA: 2-D tiled load into THIS CTA's shared memory, completion bytes reported to the barrier at `mbar_cluster_addr`
// (the pair leader's "full" barrier).  cute/arch/copy_sm100_tma.hpp: SM100_TMA_2SM_LOAD_2D.
__device__ __forceinline__ void tma2_load_2d(unsigned dst_smem, const void* tmap, unsigned mbar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(dst_smem), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
// bulk copy of `bytes` (multiple of 16) from THIS CTA's shared memory into another CTA's of the cluster; the bytes are reported
// to the mbarrier at `mbar_cluster_addr` (in the destination CTA).  SASS: UBLKCP.
__device__ __forceinline__ void dsmem_bulk_copy(unsigned dst_cluster_addr, unsigned src_cta_addr, unsigned bytes, unsigned mbar_cluster_addr) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<unsigned long long>(tmap)) : "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(unsigned taddr, unsigned (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
// fp32 pair -> packed fp16x2 "hi" and the packed fp16x2 residual "lo": v = hi + lo to ~2^-23 |v| (fp16 has 11 significant
// bits; the residual of a round-to-nearest fp16 is exactly representable in fp32 and again rounded to 11 bits).
// Four instructions per pair: F2FP, two mixed-precision FHFMA (residual = hi * -1 + v, reading the fp16 halves in place:
// PTX fma.rn.f32.f16, sm_100+), F2FP.
__device__ __forceinline__ void split2_f16(float a, float b, unsigned& hi, unsigned& lo) {
    unsigned h;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));          // low half = a, high half = b
    float ra, rb;
    asm("{\n\t.reg .b16 l, u, m1;\n\tmov.b32 {l, u}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
        "fma.rn.f32.f16 %0, l, m1, %3;\n\tfma.rn.f32.f16 %1, u, m1, %4;\n\t}\n" : "=f"(ra), "=f"(rb) : "r"(h), "f"(a), "f"(b));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
    hi = h;
}

// ------------------------------------------------------------------------------------------------------------------
// Packed-fp32 epilogue math (FFMA2 / FADD2 / FMUL2, sm_100+): the activation epilogues of the tensor kernels are bound by
// instruction issue, and two neighbouring hidden units of one sample sit in neighbouring registers after tcgen05.ld.
// Same approximate MUFU transcendentals as act_tanh<float> / tanh_only<float> (noc_types.cuh); per element 7.5, 4 and 1
// instructions for the three epilogues instead of 10, 9 and 2.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tc_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tc_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tc_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kTwoLog2e = 2.885390081777927f;          // 2 log2(e): exp(2x) = 2^(kTwoLog2e x)
// act = |pre| + ln(1 + exp(-2|pre|)) (the antiderivative of tanh, Phi.py:8-12), th = tanh(pre), for pre = acc * r
__device__ __forceinline__ void act_tanh2(float2 acc, float2 r, float2& act, float2& th) {
    const float2 pre = __fmul2_rn(acc, r);
    float2 e;
    e.x = tc_ex2(fabsf(pre.x) * -kTwoLog2e);
    e.y = tc_ex2(fabsf(pre.y) * -kTwoLog2e);
    const float2 den = __fadd2_rn(e, make_float2(1.f, 1.f));
    act.x = fmaf(tc_lg2(den.x), 0.6931471805599453f, fabsf(pre.x));
    act.y = fmaf(tc_lg2(den.y), 0.6931471805599453f, fabsf(pre.y));
    float2 q;
    q.x = tc_rcp(den.x); q.y = tc_rcp(den.y);
    const float2 t = __ffma2_rn(q, make_float2(2.f, 2.f), make_float2(-1.f, -1.f));     // (1 - e) / (1 + e) = 2 / (1 + e) - 1
    th.x = copysignf(t.x, pre.x); th.y = copysignf(t.y, pre.y);
}
// tanh(acc * r + b) * w with r and b pre-multiplied by 2 log2(e):  tanh(x) = 1 - 2 / (1 + exp(2x)), no sign handling
// (exp(2x) = inf gives 1, 0 gives -1)
__device__ __forceinline__ float2 tanh2_w(float2 acc, float2 rc, float2 bc, float2 w) {
    const float2 p = __ffma2_rn(acc, rc, bc);
    float2 e, q;
    e.x = tc_ex2(p.x); e.y = tc_ex2(p.y);
    const float2 den = __fadd2_rn(e, make_float2(1.f, 1.f));
    q.x = tc_rcp(den.x); q.y = tc_rcp(den.y);
    return __fmul2_rn(__ffma2_rn(q, make_float2(-2.f, -2.f), make_float2(1.f, 1.f)), w);
}

}  // namespace noc
