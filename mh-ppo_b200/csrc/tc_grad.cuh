// tc_grad.cuh -- the fused forward + loss + backward kernel of the PPO update with the forward pass and
// backward-data on the 5th-generation tensor cores (13-input nets: cross / wait actors and critics).
//
// Same contract as k_ppo_grad<16, HEAD> (ppo_update.cuh), which stays as the exact-fp32 cross-check and serves the wider
// choice nets.  One persistent 256-thread CTA per SM walks 128-sample tiles; in warps 0-3 thread = sample = TMEM lane.
//   forward       D[128 x J] = A[128 x K] * W[J x K]^T         (tc_mlp.cuh: 3xTF32, hi/lo operand tiles, fp32 in TMEM)
//   backward-data dIn[128 x K] = delta[128 x J] * W[J x K]     = the same MMA shapes with B = the FLAT transposed weights
//                 Wt[k][j] read as an [N = K rows][reduction = J] K-major tile (BwdTiles)
//   weight grad   dWt[k][j] += sum_s in[s][k] * delta[s][j]    FFMA register tiles (WgradAcc) on fp32 rows in shared memory:
//                 the reduction runs over samples, which a tf32 MMA could only read from 128B-swizzled MN-major tiles
//                 (tc.cuh); it is instead OVERLAPPED with the tensor core: the backward-data MMAs of layer l are issued,
//                 the CUDA cores accumulate the weight gradient of layer l while they run, then the epilogue turns the
//                 accumulator into delta_{l-1} (ReLU mask kept in registers from the forward pass).
// Every epilogue writes its row twice: fp32 into the row buffer (weight gradient) and hi/lo into the next A operand tile.
#pragma once
#include "ppo_update.cuh"
#include "tc_kernels.cuh"

namespace mhppo {

// B operands of backward-data: tile element (row = k, reduction index = j) = W[j][k] = Wt[k][j] (flat layout, mlp.cuh)
struct BwdTiles {
    static constexpr int W4 = H3 * tcm::OUTP, W3 = H2 * H3, W2 = H1 * H2;      // [32 x 16], [64 x 32], [32 x 64]
    static constexpr int FLOATS = 2 * (W4 + W3 + W2);
    float *w4h, *w4l, *w3h, *w3l, *w2h, *w2l;
    __device__ void carve(float *p) { w4h = p; p += W4; w4l = p; p += W4; w3h = p; p += W3; w3l = p; p += W3; w2h = p; p += W2; w2l = p; }
    template <int KP>
    __device__ void stage(const float *__restrict__ g) {
        auto fill = [&](float *h, float *l, const float *wt, int N, int J, int Jsrc) {   // wt[k * Jsrc + j], k < N, j < Jsrc, zero-padded to J
            for (int i = threadIdx.x; i < N * J; i += blockDim.x) {
                const int k = i / J, j = i % J;
                const float v = (j < Jsrc) ? wt[k * Jsrc + j] : 0.f, hv = tcm::hi_part(v);
                h[tc::tile_index(k, j, J)] = hv; l[tc::tile_index(k, j, J)] = v - hv;
            }
        };
        fill(w4h, w4l, g + off_w4(KP), H3, tcm::OUTP, OP);
        fill(w3h, w3l, g + off_w3(KP), H2, H3, H3);
        fill(w2h, w2l, g + off_w2(KP), H1, H2, H2);
    }
};

constexpr int kTcGradRow = 16 + H1 + H2 + H3 + OP;       // 148 floats: 16-byte aligned rows, stride = 20 (mod 32) words -> conflict-free 128-bit accesses
constexpr int kTcGradNetFloats = (tcm::NetTiles<16>::FLOATS + 255) & ~255;
constexpr size_t kTcGradSmemFloats = (size_t)kTcGradNetFloats + BwdTiles::FLOATS + 2 * 128 * H2 + (size_t)128 * kTcGradRow;

constexpr int kTcGradBlock = 256;      // warps 0-3: thread = sample (MMA issue, epilogues); all 8 warps: weight-gradient tiles
template <int HEAD>
__global__ void __launch_bounds__(kTcGradBlock, 1) k_ppo_grad_tc(SampleSet ss, const float *__restrict__ net, LossArgs la,
                                                              float *__restrict__ gpartial, double *__restrict__ lpartial,
                                                              int *__restrict__ fail_flag) {
    constexpr int KP = 16, ROW = kTcGradRow;
    typedef WgradAcc<KP, kTcGradBlock / 64> WG;
    extern __shared__ __align__(1024) float smem[];
    __shared__ TcShared sh;
    __shared__ double red[kTcGradBlock / 32];
    tcm::NetTiles<KP> w;
    BwdTiles bw;
    w.carve(smem);
    bw.carve(smem + kTcGradNetFloats);
    float *ah = smem + kTcGradNetFloats + BwdTiles::FLOATS, *al = ah + 128 * H2, *rows = al + 128 * H2;
    w.stage(net);
    bw.stage<KP>(net);
    const uint32_t tmem = tc_prologue(&sh);
    const int tid = threadIdx.x;
    const int srow = tid & 127;                           // the sample (tile row, TMEM lane) this thread works on in the epilogues
    const uint32_t trow = tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
    uint32_t phase = 0;
    bool ok = true;
    auto sync_for_mma = [&]() { tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after(); };
    auto wait_mma = [&]() { ok &= tc::mbar_wait(&sh.bar, phase); phase ^= 1; tc::fence_after(); };

    float *row = rows + (size_t)srow * ROW;
    WG wg;
    wg.init(tid);
    constexpr int HALF = 128 / (kTcGradBlock / 64);     // rows per weight-gradient group
    const int s0 = wg.half * HALF;
    const bool sample_thread = tid < 128;
    LossAcc acc;

    // Epilogues: the two threads t and t + 128 share sample t (a warp reads the TMEM lanes of quadrant warp % 4, so warps
    // w and w + 4 see the same rows); each takes half of the layer's columns: load, bias / ReLU (or mask), fp32 row, hi/lo tile.
    const int hs = tid >> 7;                              // which half of the columns
    for (int64_t base = (int64_t)blockIdx.x * 128; base < ss.Q; base += (int64_t)gridDim.x * 128) {
        int64_t s = 0;
        bool sel = false;
        uint32_t m1 = 0, m2 = 0, m3 = 0;                  // ReLU masks of this thread's columns of the three hidden layers
        if (sample_thread) {
            sel = map_sample(ss, base + tid, s);
            float x[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) x[k] = (sel && k < ss.D) ? ss.x[(int64_t)k * ss.S + s] : 0.f;
#pragma unroll
            for (int k = 0; k < KP; k += 4) st4(row + WG::X + k, make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]));
            tcm::put_row<KP>(ah, al, srow, x);
        }
        sync_for_mma();
        // ---- forward
        if (tid == 0) tcm::issue_layer<KP, H1>(tmem, ah, al, w.w1h, w.w1l, &sh.bar);
        wait_mma();
        {
            float q[16];
            const int c0 = hs * 16;
            tc::tmem_ld16(trow + c0, q);
#pragma unroll
            for (int j = 0; j < 16; ++j) { q[j] = fmaxf(q[j] + w.b1[c0 + j], 0.f); m1 |= (q[j] > 0.f ? 1u : 0u) << j; }
#pragma unroll
            for (int j = 0; j < 16; j += 4) st4(row + WG::A1 + c0 + j, make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]));
            tcm::put_cols<16>(ah, al, srow, H1, c0, q);
        }
        sync_for_mma();
        if (tid == 0) tcm::issue_layer<H1, H2>(tmem, ah, al, w.w2h, w.w2l, &sh.bar);
        wait_mma();
        {
            float q[32];
            const int c0 = hs * 32;
            tc::tmem_ld32(trow + c0, q);
#pragma unroll
            for (int j = 0; j < 32; ++j) { q[j] = fmaxf(q[j] + w.b2[c0 + j], 0.f); m2 |= (q[j] > 0.f ? 1u : 0u) << j; }
#pragma unroll
            for (int j = 0; j < 32; j += 4) st4(row + WG::A2 + c0 + j, make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]));
            tcm::put_cols<32>(ah, al, srow, H2, c0, q);
        }
        sync_for_mma();
        if (tid == 0) tcm::issue_layer<H2, H3>(tmem, ah, al, w.w3h, w.w3l, &sh.bar);
        wait_mma();
        {
            float q[16];
            const int c0 = hs * 16;
            tc::tmem_ld16(trow + c0, q);
#pragma unroll
            for (int j = 0; j < 16; ++j) { q[j] = fmaxf(q[j] + w.b3[c0 + j], 0.f); m3 |= (q[j] > 0.f ? 1u : 0u) << j; }
#pragma unroll
            for (int j = 0; j < 16; j += 4) st4(row + WG::A3 + c0 + j, make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]));
            tcm::put_cols<16>(ah, al, srow, H3, c0, q);
        }
        sync_for_mma();
        if (tid == 0) tcm::issue_layer<H3, tcm::OUTP>(tmem, ah, al, w.w4h, w.w4l, &sh.bar);
        wait_mma();
        if (sample_thread) {
            float q[16];
            tc::tmem_ld16(trow, q);
            // ---- loss epilogue -> dz (an unselected row keeps dz = 0: every gradient term of it vanishes)
            float dz[tcm::OUTP];
#pragma unroll
            for (int j = 0; j < tcm::OUTP; ++j) dz[j] = 0.f;
            float d4[OP] = {0.f, 0.f, 0.f, 0.f};
            if (sel) ppo_loss<HEAD>(la, s, make_float4(q[0] + w.b4[0], q[1] + w.b4[1], q[2] + w.b4[2], q[3] + w.b4[3]), d4, acc);
#pragma unroll
            for (int j = 0; j < OP; ++j) dz[j] = d4[j];
            st4(row + WG::D4, make_float4(d4[0], d4[1], d4[2], d4[3]));
            tcm::put_row<tcm::OUTP>(ah, al, srow, dz);
        }
        sync_for_mma();
        // ---- layer 4: backward-data on the tensor core while the CUDA cores (all 8 warps) take dW4
        if (tid == 0) tcm::issue_layer<tcm::OUTP, H3>(tmem, ah, al, bw.w4h, bw.w4l, &sh.bar);
        wg.layer4(rows, ROW, s0, HALF);
        __syncthreads();                                  // every a3 has been read before the deltas overwrite it
        wait_mma();
        {
            float q[16];
            const int c0 = hs * 16;
            tc::tmem_ld16(trow + c0, q);
#pragma unroll
            for (int j = 0; j < 16; ++j) q[j] = ((m3 >> j) & 1u) ? q[j] : 0.f;             // delta3 = relu'(a3) * (dz W4)
#pragma unroll
            for (int j = 0; j < 16; j += 4) st4(row + WG::A3 + c0 + j, make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]));
            tcm::put_cols<16>(ah, al, srow, H3, c0, q);
        }
        sync_for_mma();
        // ---- layer 3
        if (tid == 0) tcm::issue_layer<H3, H2>(tmem, ah, al, bw.w3h, bw.w3l, &sh.bar);
        wg.layer3(rows, ROW, s0, HALF);
        __syncthreads();
        wait_mma();
        {
            float q[32];
            const int c0 = hs * 32;
            tc::tmem_ld32(trow + c0, q);
#pragma unroll
            for (int j = 0; j < 32; ++j) q[j] = ((m2 >> j) & 1u) ? q[j] : 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) st4(row + WG::A2 + c0 + j, make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]));
            tcm::put_cols<32>(ah, al, srow, H2, c0, q);
        }
        sync_for_mma();
        // ---- layer 2
        if (tid == 0) tcm::issue_layer<H2, H1>(tmem, ah, al, bw.w2h, bw.w2l, &sh.bar);
        wg.layer2(rows, ROW, s0, HALF);
        __syncthreads();
        wait_mma();
        {
            float q[16];
            const int c0 = hs * 16;
            tc::tmem_ld16(trow + c0, q);
#pragma unroll
            for (int j = 0; j < 16; ++j) q[j] = ((m1 >> j) & 1u) ? q[j] : 0.f;
#pragma unroll
            for (int j = 0; j < 16; j += 4) st4(row + WG::A1 + c0 + j, make_float4(q[j], q[j + 1], q[j + 2], q[j + 3]));
        }
        tc::fence_before();
        __syncthreads();                                  // deltas visible; TMEM reads done before the next tile's MMAs
        tc::fence_after();
        // ---- layer 1 and the biases
        wg.layer1_and_biases(rows, ROW, s0, HALF);
        __syncthreads();
    }
    if (!ok) atomicExch(fail_flag, 1);
    wg.template finish<HEAD>(rows, gpartial, lpartial, la, acc, red);
    tc_epilogue(tmem);
}

}  // namespace mhppo
