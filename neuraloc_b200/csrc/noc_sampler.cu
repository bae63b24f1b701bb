// noc_sampler.cu — on-device sampler of the initial distribution rho_0 = N(xInit, var0^2 I) the reference draws on the host
// (src/initProb.py:27-28,107-120,132-140,252-262: `xInit + cvt(var0 * torch.randn(n, d))`; the quadcopter perturbs its first
// three columns only).  BASELINE.json configs[3] is 2^24 samples x 150 columns = 10 GB of input: generated here it never
// exists on the host.  Counter-based Philox4x32-10 (Salmon et al., SC'11 — the generator behind cuRAND and torch's CUDA
// randn) + Box-Muller: element e of the [n, d] matrix takes output word e % 4 of counter block e / 4 under key = seed, so the
// result depends on (seed, n, d) only, not on the launch geometry; shards draw disjoint counter ranges through `row0`.
#include <cuda_runtime.h>

#include <cstdint>

#include "noc_launch.cuh"

namespace noc {

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// four standard normals from one counter block: Box-Muller on (u0,u1) and (u2,u3), u = (word + 0.5) 2^-32 in (0,1)
template <typename real>
__device__ inline void normals4(const uint32_t (&w)[4], real (&z)[4]) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if constexpr (sizeof(real) == 4) {
            const float u1 = ((float)w[2 * p] + 0.5f) * 2.3283064365386963e-10f, u2 = ((float)w[2 * p + 1] + 0.5f) * 2.3283064365386963e-10f;
            const float r = sqrtf(-2.0f * logf(fminf(u1, 0.99999994f)));
            float sn, cs;
            sincospif(2.0f * u2, &sn, &cs);
            z[2 * p] = r * cs; z[2 * p + 1] = r * sn;
        } else {
            const double u1 = ((double)w[2 * p] + 0.5) * 2.3283064365386963e-10, u2 = ((double)w[2 * p + 1] + 0.5) * 2.3283064365386963e-10;
            const double r = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
            z[2 * p] = r * cs; z[2 * p + 1] = r * sn;
        }
    }
}

template <typename real>
static __global__ void __launch_bounds__(256) sample_rho0_kernel(const real* __restrict__ center, int d, int noise_cols, real sd, uint32_t k0,
                                                                 uint32_t k1, long long row0, long long n, real* __restrict__ x) {
    const long long total = n * d, e0 = row0 * d;                 // global element index of this shard's first element
    const long long g_first = e0 >> 2, g_last = (e0 + total - 1) >> 2;
    for (long long g = g_first + (long long)blockIdx.x * blockDim.x + threadIdx.x; g <= g_last; g += (long long)gridDim.x * blockDim.x) {
        uint32_t w[4];
        philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u, k0, k1, w);
        real z[4];
        normals4<real>(w, z);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long e = 4 * g + j - e0;
            if (e >= 0 && e < total) {
                const int c = (int)((e + e0) % d);
                x[e] = center[c] + (c < noise_cols ? sd * z[j] : real(0));
            }
        }
    }
}

static __global__ void philox_raw_kernel(uint32_t k0, uint32_t k1, long long g0, long long ngroups, uint32_t* out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ngroups; i += (long long)gridDim.x * blockDim.x) {
        uint32_t w[4];
        const long long g = g0 + i;
        philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u, k0, k1, w);
        for (int j = 0; j < 4; ++j) out[4 * i + j] = w[j];
    }
}

}  // namespace noc

using namespace noc;

extern "C" int noc_sample_rho0(const void* center, int32_t d, int32_t noise_cols, double var0, uint64_t seed, int64_t row0, int64_t n,
                               int32_t dtype, void* x, void* stream) {
    if (!center || !x || d < 1 || n < 1 || row0 < 0) return fail(NOC_ERR_ARG, "noc_sample_rho0: bad arguments");
    if (noise_cols < 0 || noise_cols > d) noise_cols = d;
    cudaStream_t st = (cudaStream_t)stream;
    const long long groups = (n * d + 3) / 4 + 1;
    const int grid = (int)std::min<long long>((groups + 255) / 256, 148 * 16);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    if (dtype == NOC_F32)
        sample_rho0_kernel<float><<<grid, 256, 0, st>>>((const float*)center, d, noise_cols, (float)var0, k0, k1, row0, n, (float*)x);
    else if (dtype == NOC_F64)
        sample_rho0_kernel<double><<<grid, 256, 0, st>>>((const double*)center, d, noise_cols, var0, k0, k1, row0, n, (double*)x);
    else
        return fail(NOC_ERR_ARG, "dtype must be NOC_F32 or NOC_F64");
    count_launch();
    NOC_CUDA(cudaGetLastError());
    return NOC_OK;
}

extern "C" int noc_philox_raw(uint64_t seed, int64_t group0, int64_t ngroups, void* out_u32, void* stream) {
    if (!out_u32 || ngroups < 1) return fail(NOC_ERR_ARG, "noc_philox_raw: bad arguments");
    philox_raw_kernel<<<(int)std::min<long long>((ngroups + 255) / 256, 1024), 256, 0, (cudaStream_t)stream>>>(
        (uint32_t)seed, (uint32_t)(seed >> 32), group0, ngroups, (uint32_t*)out_u32);
    count_launch();
    NOC_CUDA(cudaGetLastError());
    return NOC_OK;
}
