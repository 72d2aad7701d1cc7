// tc_selftest.cu -- self-test of the tcgen05 building blocks (tc.cuh): D[128 x N] = A[128 x K] * B[N x K]^T on one CTA.
// mode 0: single-pass tf32 (inputs truncated to 10-bit mantissa by the tensor core); mode 1: 3xTF32 split
// (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo), which is what the policy GEMMs use to stay inside the 1e-5 parity bar.
#include <string>

#include "../../include/mhppo.h"
#include "tc.cuh"

namespace mhppo {
int api_fail(int code, const std::string &msg);
void api_count_launch();

template <int N, int K>
__global__ void __launch_bounds__(128) k_tc_selftest(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int mode) {
    extern __shared__ __align__(1024) float smem[];
    float *a_hi = smem, *a_lo = a_hi + 128 * K, *b_hi = a_lo + 128 * K, *b_lo = b_hi + N * K;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        const float v = A[i], hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        a_hi[tc::tile_index(r, k, K)] = hi; a_lo[tc::tile_index(r, k, K)] = v - hi;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        const float v = B[i], hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        b_hi[tc::tile_index(r, k, K)] = hi; b_lo[tc::tile_index(r, k, K)] = v - hi;
    }
    if (tid == 0) tc::mbar_init(&bar, 1);
    if (warp == 0) tc::tmem_alloc(&tmem_base, N < 32 ? 32 : N);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, N);
        bool acc = false;
        const int npass = mode ? 3 : 1;
        for (int pass = 0; pass < npass; ++pass) {
            const float *pa = (pass == 1) ? a_lo : a_hi, *pb = (pass == 2) ? b_lo : b_hi;
            for (int k0 = 0; k0 < K; k0 += 8) {
                const uint64_t ad = tc::make_desc(tc::smem_u32(pa) + (k0 / 4) * 128, K);
                const uint64_t bd = tc::make_desc(tc::smem_u32(pb) + (k0 / 4) * 128, K);
                tc::mma_tf32(tm, ad, bd, idesc, acc);
                acc = true;
            }
        }
        tc::mma_commit(&bar);
    }
    const bool ok = tc::mbar_wait(&bar, 0);
    tc::fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c0 + j < N) D[(size_t)tid * N + c0 + j] = ok ? v[j] : __int_as_float(0x7fc00000);
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, N < 32 ? 32 : N);
}
}  // namespace mhppo

using namespace mhppo;

extern "C" int mhppo_tc_selftest(const float *A_dev, const float *B_dev, float *D_dev, int32_t N, int32_t K, int32_t mode, void *stream) {
    if (!A_dev || !B_dev || !D_dev) return api_fail(MHPPO_EINVAL, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t sm = sizeof(float) * 2 * ((size_t)128 * K + (size_t)N * K) + 1024;
    cudaError_t e = cudaSuccess;
    if (N == 64 && K == 32) {
        e = cudaFuncSetAttribute(k_tc_selftest<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) k_tc_selftest<64, 32><<<1, 128, sm, s>>>(A_dev, B_dev, D_dev, mode);
    } else if (N == 32 && K == 64) {
        e = cudaFuncSetAttribute(k_tc_selftest<32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) k_tc_selftest<32, 64><<<1, 128, sm, s>>>(A_dev, B_dev, D_dev, mode);
    } else if (N == 64 && K == 16) {
        e = cudaFuncSetAttribute(k_tc_selftest<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) k_tc_selftest<64, 16><<<1, 128, sm, s>>>(A_dev, B_dev, D_dev, mode);
    } else if (N == 16 && K == 32) {
        e = cudaFuncSetAttribute(k_tc_selftest<16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) k_tc_selftest<16, 32><<<1, 128, sm, s>>>(A_dev, B_dev, D_dev, mode);
    } else return api_fail(MHPPO_EUNSUPPORTED, "selftest shapes: (N,K) in {(64,32),(32,64),(64,16),(16,32)}");
    api_count_launch();
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? 0 : api_fail(MHPPO_ECUDA, cudaGetErrorString(e));
}
