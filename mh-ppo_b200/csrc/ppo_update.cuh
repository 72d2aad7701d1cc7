// ppo_update.cuh -- update-side kernels of the PPO loop (Algo_PPO.train_model_c / train_model_d,
// PY:778-851): critic forward + advantage statistics, fused actor/critic forward + clipped-surrogate
// / MSE loss + backward with per-CTA gradient accumulation, deterministic gradient reduction, Adam.
//
// One epoch of train_model_c over the samples selected by `route` (cross or wait buffer):
//   1. k_value_stats : V = critic(s), A = RTG - V, per-CTA partial (sum A, sum A^2, n) in fp64
//   2. (host: reduce partials; multi-GPU: all-reduce the 3 scalars -> global mean / unbiased std)
//   3. k_ppo_grad    : forward both nets, loss epilogue, backward, dW partial per CTA
//   4. k_reduce      : fixed-order sum of the CTA partials -> flat gradient (deterministic)
//   5. (multi-GPU: one NCCL all-reduce of the flat gradient buffer)       6. k_adam
//
// Execution model: thread = sample for forward and backward-data (activations of the CTA's 128
// samples sit in shared memory as [sample][feature], deltas overwrite activations in place), then the
// same 128 threads switch roles and each owns a strip of every weight matrix for the weight
// gradient dWt[k][j] += sum_s in[s][k] * delta[s][j], accumulated in registers across all tiles
// of a persistent CTA.
#pragma once
#include "mlp.cuh"

namespace mhppo {

struct SampleSet {
    const float *x;        // features, feature-major [D][S]
    int D;                 // real feature count (<= KP)
    int64_t S;             // sample slots in the buffers (row pitch of x)
    const int32_t *idx;    // compacted list of the selected (car, env) columns m in [0, CN), or null = every column
    int64_t K;             // number of selected columns (CN when idx is null)
    int64_t CN;            // columns per time step; sample slot s = t*CN + m
    int64_t Q;             // work items = T*K; item q -> (t = q / K, k = q % K)
};
// work item -> sample slot; false past the end.  The update walks only the samples routed to the net being
// trained (cross or wait buffer, PY:489-502), so every lane of a tile is busy.
__device__ __forceinline__ bool map_sample(const SampleSet &ss, int64_t q, int64_t &s) {
    if (q >= ss.Q) return false;
    const int64_t t = q / ss.K, k = q - t * ss.K;
    s = t * ss.CN + (ss.idx ? (int64_t)ss.idx[k] : k);
    return true;
}

template <int KP>
__device__ __forceinline__ void load_row(const SampleSet &ss, int64_t s, float *x) {
#pragma unroll
    for (int k = 0; k < KP; ++k) x[k] = (k < ss.D) ? ss.x[(int64_t)k * ss.S + s] : 0.f;
}

// sum over a team of `nthreads` consecutive threads (a whole CTA, or one 128-thread team of it synchronising on its own
// named barrier `bar`); the result is valid in the team's first thread
__device__ __forceinline__ void team_sync(int bar, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory"); }
__device__ __forceinline__ double team_sum(double v, double *red, int ltid, int nthreads, int bar) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((ltid & 31) == 0) red[ltid >> 5] = v;
    team_sync(bar, nthreads);
    double r = 0.0;
    if (ltid == 0) for (int w = 0; w < nthreads / 32; ++w) r += red[w];
    team_sync(bar, nthreads);
    return r;
}

// ---- 1. critic forward + advantage statistics --------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(kFwdBlock, 2) k_value_stats(SampleSet ss, const float *__restrict__ critic, const float *__restrict__ rtg,
                                                           float *__restrict__ V, double *__restrict__ partial /* [grid][3] */) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;
    float *rows = smem + ((net_params(KP) + 3) & ~3);
    __shared__ double red[kFwdBlock / 32];
    stage_net<KP>(sw, critic);
    __syncthreads();
    float *x = rows + (size_t)threadIdx.x * kRowFwd;      // one in-place row per sample
    double sA = 0.0, sAA = 0.0, cnt = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * kFwdBlock; base < ss.Q; base += (int64_t)gridDim.x * kFwdBlock) {
        int64_t s = 0;
        if (map_sample(ss, base + threadIdx.x, s)) {
            load_row<KP>(ss, s, x);
            const float v = mlp_fwd_inplace<KP>(sw, x).x;
            V[s] = v;
            const double A = (double)(rtg[s] - v);          // advantage_batch = rtgs - V, fp32 like torch (PY:786)
            sA += A; sAA += A * A; cnt += 1.0;
        }
    }
    const double t0 = team_sum(sA, red, threadIdx.x, kFwdBlock, 0), t1 = team_sum(sAA, red, threadIdx.x, kFwdBlock, 0),
                 t2 = team_sum(cnt, red, threadIdx.x, kFwdBlock, 0);
    if (threadIdx.x == 0) { partial[blockIdx.x * 3 + 0] = t0; partial[blockIdx.x * 3 + 1] = t1; partial[blockIdx.x * 3 + 2] = t2; }
}

// ---- 3. fused forward + loss + backward ----------------------------------------------------------------
struct LossArgs {
    const float *act, *logp_old, *rtg, *V;     // per-sample
    float adv_mean, adv_inv_std;               // (A - mean) * inv_std, inv_std = 1 / (std_unbiased + 1e-10)  (PY:787)
    float inv_n;                               // 1 / (global sample count)
    float f0, f1;                              // choice only: fraction of samples whose action is 0 / 1 (PY:834-842 broadcast)
};

// backward-data for this thread's sample: din[k] = relu'(ain[k]) * sum_j dout[j] * W[j][k], written over ain
template <int K, int J>
__device__ __forceinline__ void dense_bwd_data(float *__restrict__ ain_din, const float *__restrict__ dout, const float *__restrict__ W /* [J][K] */) {
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
#pragma unroll 2
    for (int j = 0; j < J; ++j) {
        const float dj = dout[j];
        const float4 *w = reinterpret_cast<const float4 *>(W + j * K);
#pragma unroll
        for (int k4 = 0; k4 < K / 4; ++k4) {
            const float4 ww = w[k4];
            acc[4 * k4 + 0] = fmaf(dj, ww.x, acc[4 * k4 + 0]); acc[4 * k4 + 1] = fmaf(dj, ww.y, acc[4 * k4 + 1]);
            acc[4 * k4 + 2] = fmaf(dj, ww.z, acc[4 * k4 + 2]); acc[4 * k4 + 3] = fmaf(dj, ww.w, acc[4 * k4 + 3]);
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) ain_din[k] = (ain_din[k] > 0.f) ? acc[k] : 0.f;
}

// weight-gradient strip owned by this thread: NJ consecutive outputs j0.. of input row kk, summed over the tile
template <int NJ>
__device__ __forceinline__ void wgrad_strip(float (&acc)[NJ], const float *__restrict__ in, int in_stride, int kk,
                                            const float *__restrict__ dl, int dl_stride, int j0, bool active) {
    if (!active) return;
#pragma unroll 4
    for (int s = 0; s < kMlpBlock; ++s) {
        const float a = in[s * in_stride + kk];
        const float *dr = dl + s * dl_stride + j0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[j] = fmaf(a, dr[j], acc[j]);
    }
}

// HEAD: 0 = critic (MSE), 1 = Gaussian actor (clipped surrogate), 2 = categorical actor with the (M,M) broadcast
// A CTA holds kTeams independent 128-thread teams that share the staged weights; each team walks its own tiles and
// synchronises on its own named barrier, so two tiles are in flight per SM (shared memory allows one 128-row tile set
// per team, registers allow 256 threads).
constexpr int kTeams = 1;
template <int KP, int HEAD>
__global__ void __launch_bounds__(kMlpBlock * kTeams) k_ppo_grad(SampleSet ss, const float *__restrict__ net, LossArgs la,
                                                        float *__restrict__ gpartial /* [grid][net_params] */,
                                                        double *__restrict__ lpartial /* [grid] */) {
    extern __shared__ __align__(16) float smem[];
    typedef Strides<KP> St;
    constexpr int NPAR = net_params(KP), NP4 = (NPAR + 3) & ~3;
    float *sw = smem;                                  // flat net (transposed weights)
    float *W2 = sw + NP4;                              // [H2][H1] original orientation, for backward-data
    float *W3 = W2 + H2 * H1;                          // [H3][H2]
    float *W4 = W3 + H3 * H2;                          // [OP][H3]
    float *rows = W4 + OP * H3;
    constexpr int ROW = (St::X + St::A1 + St::A2 + St::A3 + OP) | 1;   // odd row stride: lane = sample is conflict-free
    __shared__ double red_all[kTeams][kMlpBlock / 32];
    stage_net<KP>(sw, net);
    __syncthreads();
    for (int i = threadIdx.x; i < H1 * H2; i += blockDim.x) { const int k = i / H2, j = i % H2; W2[j * H1 + k] = sw[off_w2(KP) + i]; }
    for (int i = threadIdx.x; i < H2 * H3; i += blockDim.x) { const int k = i / H3, j = i % H3; W3[j * H2 + k] = sw[off_w3(KP) + i]; }
    for (int i = threadIdx.x; i < H3 * OP; i += blockDim.x) { const int k = i / OP, j = i % OP; W4[j * H3 + k] = sw[off_w4(KP) + i]; }
    __syncthreads();

    const int team = threadIdx.x / kMlpBlock, tid = threadIdx.x % kMlpBlock, bar = team + 1;
    const int unit = blockIdx.x * kTeams + team, nunits = gridDim.x * kTeams;
    rows += (size_t)team * kMlpBlock * ROW;
    double *red = red_all[team];
#define TEAM_SYNC() team_sync(bar, kMlpBlock)
    float *x = rows + (size_t)tid * ROW, *a1 = x + St::X, *a2 = a1 + St::A1, *a3 = a2 + St::A2, *d4 = a3 + St::A3;
    // weight-gradient strips of this thread (persist over all tiles of this CTA)
    constexpr int NJ1 = (KP <= 16) ? 4 : ((KP <= 32) ? 8 : 16);
    constexpr int KK1 = (KP <= 16) ? 16 : ((KP <= 32) ? 32 : 64);       // threads along k for layer 1
    float g1[NJ1], g2[16], g3[16], g4 = 0.f, gb = 0.f, gbias = 0.f;
    // bias gradients b3 | b2 | b1: 32 + 64 + 32 = 128 columns, one per thread
    const int bias_off = (tid < 32) ? (St::X + St::A1 + St::A2 + tid) : ((tid < 96) ? (St::X + St::A1 + (tid - 32)) : (St::X + (tid - 96)));
#pragma unroll
    for (int j = 0; j < NJ1; ++j) g1[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { g2[j] = 0.f; g3[j] = 0.f; }
    double loss = 0.0;

    for (int64_t base = (int64_t)unit * kMlpBlock; base < ss.Q; base += (int64_t)nunits * kMlpBlock) {
        int64_t s = 0;
        const bool sel = map_sample(ss, base + tid, s);
        float dz[OP] = {0.f, 0.f, 0.f, 0.f};
        if (sel) {
            load_row<KP>(ss, s, x);
            const float4 o = mlp_fwd_rows<KP>(sw, x, a1, a2, a3);
            if (HEAD == 0) {                                           // critic_loss = MSE(V, rtg), PY:808-809
                const float e = o.x - la.rtg[s];
                loss += (double)e * (double)e * (double)la.inv_n;
                dz[0] = 2.0f * e * la.inv_n;
            } else {
                const float An = ((la.rtg[s] - la.V[s]) - la.adv_mean) * la.adv_inv_std;     // PY:786-787
                if (HEAD == 1) {
                    const float th = tanhf(o.x), mu = th * 3.0f + (-1.0f);
                    const float a = la.act[s];
                    const float lp = -((a - mu) * (a - mu)) - 0.57236494292470008f;          // MVN(mu, .5 I).log_prob, PY:795-800
                    const float ratio = expf(lp - la.logp_old[s]);                           // PY:803
                    const float s1 = ratio * An, s2 = fminf(fmaxf(ratio, 0.8f), 1.2f) * An;   // PY:804-805
                    loss += (double)(-fminf(s1, s2)) * (double)la.inv_n;                     // PY:806
                    if (s1 <= s2) dz[0] = -la.inv_n * An * ratio * (2.0f * (a - mu)) * (3.0f * (1.0f - th * th));
                } else {
                    // Categorical(probs [M,2]).log_prob(actions [M,1]) broadcasts to (M,M): lp[i,j] = log p_j(a_i)
                    // (PY:834-842).  With two actions the mean over i collapses to the action frequencies f0, f1.
                    const float m = fmaxf(o.x, o.y), e0 = expf(o.x - m), e1 = expf(o.y - m);
                    const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
                    const float inv_old = expf(-la.logp_old[s]);
                    float dp0 = 0.f, dp1 = 0.f;
                    {
                        const float r = p0 * inv_old, s1 = r * An, s2 = fminf(fmaxf(r, 0.8f), 1.2f) * An;
                        loss += (double)(-fminf(s1, s2)) * (double)(la.f0 * la.inv_n);
                        if (s1 <= s2) dp0 = -la.inv_n * la.f0 * An * inv_old;
                    }
                    {
                        const float r = p1 * inv_old, s1 = r * An, s2 = fminf(fmaxf(r, 0.8f), 1.2f) * An;
                        loss += (double)(-fminf(s1, s2)) * (double)(la.f1 * la.inv_n);
                        if (s1 <= s2) dp1 = -la.inv_n * la.f1 * An * inv_old;
                    }
                    // softmax backward: dz_a = p_a * (dp_a - sum_b p_b dp_b)
                    const float dot = p0 * dp0 + p1 * dp1;
                    dz[0] = p0 * (dp0 - dot); dz[1] = p1 * (dp1 - dot);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < KP; ++k) x[k] = 0.f;
            for (int k = 0; k < H1; ++k) a1[k] = 0.f;
            for (int k = 0; k < H2; ++k) a2[k] = 0.f;
            for (int k = 0; k < H3; ++k) a3[k] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < OP; ++j) d4[j] = dz[j];
        TEAM_SYNC();
        // layer-4 weight / bias gradient: thread (k = tid % 32, j = tid / 32), 32*4 = 128 entries
        {
            const int kk = tid & 31, j = tid >> 5;
            float acc = 0.f, accb = 0.f;
            for (int s2 = 0; s2 < kMlpBlock; ++s2) {
                const float dd = rows[s2 * ROW + (St::X + St::A1 + St::A2 + St::A3) + j];
                acc = fmaf(rows[s2 * ROW + (St::X + St::A1 + St::A2) + kk], dd, acc);
                accb += dd;
            }
            g4 += acc;
            if (kk == 0) gb += accb;          // b4[j]
        }
        TEAM_SYNC();
        dense_bwd_data<H3, OP>(a3, d4, W4);                            // a3 now holds delta3
        TEAM_SYNC();
        wgrad_strip<16>(g3, rows + St::X + St::A1, ROW, tid & 63, rows + St::X + St::A1 + St::A2, ROW, (tid >> 6) * 16, true);   // dW3t[k<64][j<32]
        TEAM_SYNC();
        dense_bwd_data<H2, H3>(a2, a3, W3);                            // a2 now holds delta2
        TEAM_SYNC();
        wgrad_strip<16>(g2, rows + St::X, ROW, tid & 31, rows + St::X + St::A1, ROW, (tid >> 5) * 16, true);                      // dW2t[k<32][j<64]
        TEAM_SYNC();
        dense_bwd_data<H1, H2>(a1, a2, W2);                            // a1 now holds delta1
        TEAM_SYNC();
        wgrad_strip<NJ1>(g1, rows, ROW, tid % KK1, rows + St::X, ROW, (tid / KK1) * NJ1, (tid % KK1) < KP);                        // dW1t[k<KP][j<32]
        // bias gradients b3, b2, b1: the deltas now sit in a3, a2, a1 of every row
        {
            float acc = 0.f;
#pragma unroll 4
            for (int s2 = 0; s2 < kMlpBlock; ++s2) acc += rows[s2 * ROW + bias_off];
            gbias += acc;
        }
        TEAM_SYNC();
    }
    float *gp = gpartial + (size_t)unit * NPAR;
    {
        const int kk = tid % KK1, j0 = (tid / KK1) * NJ1;
        if (kk < KP)
#pragma unroll
            for (int j = 0; j < NJ1; ++j) gp[kk * H1 + j0 + j] = g1[j];
    }
    {
        const int kk = tid & 31, j0 = (tid >> 5) * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) gp[off_w2(KP) + kk * H2 + j0 + j] = g2[j];
    }
    {
        const int kk = tid & 63, j0 = (tid >> 6) * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) gp[off_w3(KP) + kk * H3 + j0 + j] = g3[j];
    }
    {
        const int kk = tid & 31, j = tid >> 5;
        gp[off_w4(KP) + kk * OP + j] = g4;
        if (kk == 0) gp[off_b4(KP) + j] = gb;
    }
    gp[(tid < 32) ? (off_b3(KP) + tid) : ((tid < 96) ? (off_b2(KP) + tid - 32) : (off_b1(KP) + tid - 96))] = gbias;
    const double lt = team_sum(loss, red, tid, kMlpBlock, bar);
    if (tid == 0) lpartial[unit] = lt;
#undef TEAM_SYNC
}

// ---- 4. deterministic reduction of the per-CTA partials ----------------------------------------------
__global__ void __launch_bounds__(256) k_reduce_partials(const float *__restrict__ gpartial, int nblocks, int npar, float *__restrict__ grad) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= npar) return;
    float acc = 0.f;
    for (int b = 0; b < nblocks; ++b) acc += gpartial[(size_t)b * npar + i];
    grad[i] = acc;
}
__global__ void k_reduce_scalars(const double *__restrict__ partial, int nblocks, int width, double *__restrict__ out) {
    const int j = threadIdx.x;
    if (j >= width) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += partial[(size_t)b * width + j];
    out[j] = acc;
}

// ---- 6. Adam (torch.optim.Adam defaults, PY:719-724): p -= step_size * m / (sqrt(v) * inv_sqrt_bc2 + eps) ----
__global__ void __launch_bounds__(256) k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                              float *__restrict__ v, int n, float beta1, float beta2, float step_size,
                                              float inv_sqrt_bc2, float eps, float grad_scale) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * grad_scale;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;                 // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
}

}  // namespace mhppo
