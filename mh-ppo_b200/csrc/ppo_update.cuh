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
// Execution model: thread = sample(s) for forward and backward-data (activations of the CTA's tile sit
// in shared memory as [sample][feature], deltas overwrite activations in place), then the same 128
// threads switch roles and each owns a register tile of every weight matrix for the weight
// gradient dWt[k][j] += sum_s in[s][k] * delta[s][j], accumulated across all tiles of a
// persistent CTA (one per SM).
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
    float *V_out;                              // critic only, optional: V[s] of this forward pass and the advantage statistics
    double *spartial;                          //   [grid][3] partial (sum A, sum A^2, n) with A = rtg - V (replaces k_value_stats)
    HeadCfg head;                              // Gaussian head (HEAD == 1)
    const double *stats_dev;                   // actor heads, optional: (sum A, sum A^2, n) of the whole (all-rank) batch in device
                                               //   memory; adv_mean / adv_inv_std are then derived in the kernel (no host round trip)
};

// advantage normalisation of PY:787 from the partial sums: mean, 1 / (unbiased std + 1e-10) (Algo_PPO.combine_stats)
__device__ __forceinline__ void resolve_adv_stats(LossArgs &la) {
    if (!la.stats_dev) return;
    const double sA = la.stats_dev[0], sAA = la.stats_dev[1], n = la.stats_dev[2];
    if (n < 1.0) { la.adv_mean = 0.f; la.adv_inv_std = 0.f; return; }
    const double mean = sA / n;
    const double var = (n > 1.0) ? fmax(sAA - n * mean * mean, 0.0) / (n - 1.0) : nan("");
    la.adv_mean = (float)mean; la.adv_inv_std = (float)(1.0 / (sqrt(var) + 1e-10));
}

// ---- register-tiled dense layers for the fused kernel -------------------------------------------------
// The first version read one weight per FFMA from shared memory (a 128-bit broadcast load feeds 4 FFMAs of
// one sample) and the weight-gradient strips read one delta per FFMA: 3 shared-memory wavefront cycles per
// FFMA issue cycle, FMA pipe 27 % busy (profiles/round1_ppo_kernels_summary.txt).  Here every thread owns R
// samples (rows tid, tid + 128, ...) so a weight load feeds R x 4 FFMAs, every row access is a 128-bit
// load/store (rows are 16-byte aligned, row stride = 4 mod 32 words: 8 lanes x 16 B per wavefront, conflict-free),
// and the weight gradient is a 4 x 8 register tile per thread.
template <int KP>
struct GradCfg {
    static constexpr int R = (KP <= 16) ? 2 : 1;                    // samples per thread (shared memory bounds it)
    static constexpr int TILE = kMlpBlock * R;                      // samples per CTA tile
    static constexpr int X = 0, A1 = KP, A2 = KP + H1, A3 = KP + H1 + H2, D4 = KP + H1 + H2 + H3, RAW = D4 + OP;
    static constexpr int ROW = ((RAW - 4 + 31) / 32) * 32 + 4;      // >= RAW and = 4 (mod 32)
    static constexpr int NPAR = net_params(KP), NP4 = (NPAR + 3) & ~3;
    static constexpr size_t smem_floats = (size_t)NP4 + H2 * H1 + H3 * H2 + OP * H3 + (size_t)TILE * ROW;
};

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// out[r][j] = act(b[j] + sum_k in[r][k] * Wt[k][j]) for the R rows of this thread; rows are `rs` floats apart
template <int K, int J, bool RELU, int R>
__device__ __forceinline__ void dense_fwd_t(const float *in, float *out, int rs, const float *__restrict__ Wt, const float *__restrict__ b) {
    constexpr int JC = (J < 32) ? J : 32;
#pragma unroll 1
    for (int jc = 0; jc < J; jc += JC) {
        float acc[R][JC];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < JC; ++j) acc[r][j] = b[jc + j];
#pragma unroll 2
        for (int k = 0; k < K; k += 4) {
            float a[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) { const float4 v = ld4(in + r * rs + k); a[r][0] = v.x; a[r][1] = v.y; a[r][2] = v.z; a[r][3] = v.w; }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int j4 = 0; j4 < JC / 4; ++j4) {
                    const float4 w = ld4(Wt + (k + kk) * J + jc + 4 * j4);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r][4 * j4 + 0] = fmaf(a[r][kk], w.x, acc[r][4 * j4 + 0]); acc[r][4 * j4 + 1] = fmaf(a[r][kk], w.y, acc[r][4 * j4 + 1]);
                        acc[r][4 * j4 + 2] = fmaf(a[r][kk], w.z, acc[r][4 * j4 + 2]); acc[r][4 * j4 + 3] = fmaf(a[r][kk], w.w, acc[r][4 * j4 + 3]);
                    }
                }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j4 = 0; j4 < JC / 4; ++j4) {
                float4 o = make_float4(acc[r][4 * j4], acc[r][4 * j4 + 1], acc[r][4 * j4 + 2], acc[r][4 * j4 + 3]);
                if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                st4(out + r * rs + jc + 4 * j4, o);
            }
    }
}

// backward-data for the R rows of this thread: din[k] = relu'(ain[k]) * sum_j dout[j] * W[j][k], written over ain
template <int K, int J, int R>
__device__ __forceinline__ void dense_bwd_t(float *ain_din, const float *dout, int rs, const float *__restrict__ W /* [J][K] */) {
    constexpr int KC = (K < 32) ? K : 32;
#pragma unroll 1
    for (int kc = 0; kc < K; kc += KC) {
        float acc[R][KC];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[r][k] = 0.f;
#pragma unroll 2
        for (int j = 0; j < J; j += 4) {
            float d[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r) { const float4 v = ld4(dout + r * rs + j); d[r][0] = v.x; d[r][1] = v.y; d[r][2] = v.z; d[r][3] = v.w; }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    const float4 w = ld4(W + (j + jj) * K + kc + 4 * k4);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r][4 * k4 + 0] = fmaf(d[r][jj], w.x, acc[r][4 * k4 + 0]); acc[r][4 * k4 + 1] = fmaf(d[r][jj], w.y, acc[r][4 * k4 + 1]);
                        acc[r][4 * k4 + 2] = fmaf(d[r][jj], w.z, acc[r][4 * k4 + 2]); acc[r][4 * k4 + 3] = fmaf(d[r][jj], w.w, acc[r][4 * k4 + 3]);
                    }
                }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                float *q = ain_din + r * rs + kc + 4 * k4;
                const float4 av = ld4(q);
                st4(q, make_float4(av.x > 0.f ? acc[r][4 * k4] : 0.f, av.y > 0.f ? acc[r][4 * k4 + 1] : 0.f,
                                   av.z > 0.f ? acc[r][4 * k4 + 2] : 0.f, av.w > 0.f ? acc[r][4 * k4 + 3] : 0.f));
            }
    }
}

#ifndef MH_WGRAD_UNROLL_N
#define MH_WGRAD_UNROLL_N 2
#endif
constexpr int MH_WGRAD_UNROLL = MH_WGRAD_UNROLL_N;      // rows in flight per thread in the weight-gradient loops
// weight-gradient register tile: acc[kk][j] += in[s][k0 + kk] * delta[s][j0 + j] over the rows [s0, s0 + ns) of the tile
template <int TJ>
__device__ __forceinline__ void wgrad_tile(float (&acc)[4][TJ], const float *__restrict__ rows, int row, int in_off, int dl_off, int s0, int ns) {
#pragma unroll 2
    for (int s = s0; s < s0 + ns; ++s) {
        const float *r = rows + (size_t)s * row;
        const float4 av = ld4(r + in_off);
        const float a[4] = {av.x, av.y, av.z, av.w};
        float d[TJ];
        if (TJ == 2) { const float2 v = *reinterpret_cast<const float2 *>(r + dl_off); d[0] = v.x; d[1] = v.y; }
        else {
#pragma unroll
            for (int j4 = 0; j4 < TJ / 4; ++j4) { const float4 v = ld4(r + dl_off + 4 * j4); d[4 * j4] = v.x; d[4 * j4 + 1] = v.y; d[4 * j4 + 2] = v.z; d[4 * j4 + 3] = v.w; }
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int j = 0; j < TJ; ++j) acc[kk][j] = fmaf(a[kk], d[j], acc[kk][j]);
    }
}

// ---- loss epilogue of one sample (shared by the FFMA and the tensor-core kernel) ------------------------
// HEAD: 0 = critic (MSE), 1 = Gaussian actor (clipped surrogate), 2 = categorical actor with the (M,M) broadcast.
// o: raw outputs of the net; returns dLoss/dz (4 padded outputs) and accumulates the loss and, for the critic,
// the advantage statistics.
struct LossAcc { double loss = 0.0, sA = 0.0, sAA = 0.0, cnt = 0.0; };
// the per-sample inputs of the loss (rtg for every head; V, action, old log-prob for the actors): can be loaded ahead of use
struct LossIn { float rtg, V, act, logp; };
template <int HEAD>
__device__ __forceinline__ LossIn load_loss_in(const LossArgs &la, int64_t s) {
    LossIn in; in.rtg = la.rtg[s]; in.V = 0.f; in.act = 0.f; in.logp = 0.f;
    if (HEAD != 0) { in.V = la.V[s]; in.logp = la.logp_old[s]; }
    if (HEAD == 1) in.act = la.act[s];
    return in;
}
template <int HEAD>
__device__ __forceinline__ void ppo_loss(const LossArgs &la, int64_t s, const LossIn &in, float4 o, float (&dz)[OP], LossAcc &acc) {
    if (HEAD == 0) {                                           // critic_loss = MSE(V, rtg), PY:808-809
        const float e = o.x - in.rtg;
        acc.loss += (double)e * (double)e * (double)la.inv_n;
        dz[0] = 2.0f * e * la.inv_n;
        if (la.V_out) {                                        // V = critic(s), A = rtgs - V (PY:785-786)
            la.V_out[s] = o.x;
            const double A = (double)(in.rtg - o.x);
            acc.sA += A; acc.sAA += A * A; acc.cnt += 1.0;
        }
        return;
    }
    const float An = ((in.rtg - in.V) - la.adv_mean) * la.adv_inv_std;     // PY:786-787
    if (HEAD == 1) {
        const float th = tanhf(o.x), mu = th * la.head.std + la.head.mean;
        const float a = in.act;
        const float lp = -((a - mu) * (a - mu)) * la.head.inv_2var - la.head.logp_c;   // MVN(mu, var I).log_prob, PY:795-800
        const float ratio = expf(lp - in.logp);                           // PY:803
        const float s1 = ratio * An, s2 = fminf(fmaxf(ratio, 0.8f), 1.2f) * An;   // PY:804-805
        acc.loss += (double)(-fminf(s1, s2)) * (double)la.inv_n;                 // PY:806
        if (s1 <= s2) dz[0] = -la.inv_n * An * ratio * (2.0f * la.head.inv_2var * (a - mu)) * (la.head.std * (1.0f - th * th));
    } else {
        // Categorical(probs [M,2]).log_prob(actions [M,1]) broadcasts to (M,M): lp[i,j] = log p_j(a_i)
        // (PY:834-842).  With two actions the mean over i collapses to the action frequencies f0, f1.
        const float m = fmaxf(o.x, o.y), e0 = expf(o.x - m), e1 = expf(o.y - m);
        const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
        const float inv_old = expf(-in.logp);
        float dp0 = 0.f, dp1 = 0.f;
        {
            const float rr = p0 * inv_old, s1 = rr * An, s2 = fminf(fmaxf(rr, 0.8f), 1.2f) * An;
            acc.loss += (double)(-fminf(s1, s2)) * (double)(la.f0 * la.inv_n);
            if (s1 <= s2) dp0 = -la.inv_n * la.f0 * An * inv_old;
        }
        {
            const float rr = p1 * inv_old, s1 = rr * An, s2 = fminf(fmaxf(rr, 0.8f), 1.2f) * An;
            acc.loss += (double)(-fminf(s1, s2)) * (double)(la.f1 * la.inv_n);
            if (s1 <= s2) dp1 = -la.inv_n * la.f1 * An * inv_old;
        }
        // softmax backward: dz_a = p_a * (dp_a - sum_b p_b dp_b)
        const float dot = p0 * dp0 + p1 * dp1;
        dz[0] = p0 * (dp0 - dot); dz[1] = p1 * (dp1 - dot);
    }
}

// ---- the weight-gradient register tiles of one thread (persist over all tiles of a CTA) ------------------
// The CTA's threads split into NG groups of 64; each group sums over 1/NG of the tile's rows and every thread of a
// group owns a 4 x TJ tile of every matrix; the groups are added once, after the last tile (finish).  `half` = group.
// Row layout: [x KP | a1/delta1 32 | a2/delta2 64 | a3/delta3 32 | dz 4], `row` floats apart.
template <int KP, int NG = 2>
struct WgradAcc {
    static constexpr int NKG1 = KP / 4;                                        // k-groups of layer 1
    static constexpr int JG1 = (NKG1 <= 4) ? 16 : ((NKG1 <= 8) ? 8 : 4);       // j-groups of layer 1 (NKG1 * JG1 <= 64 threads)
    static constexpr int TJ1 = H1 / JG1;
    static constexpr int X = 0, A1 = KP, A2 = KP + H1, A3 = KP + H1 + H2, D4 = KP + H1 + H2 + H3;
    float g1[4][TJ1], g2[4][8], g3[4][8], g4[2], gb4[2], gbias[2];
    int half, lt, k1, j1, k2, j2, k3, j3, k4, j4;
    bool l1_on;
    __device__ __forceinline__ void init(int tid) {
        half = tid >> 6; lt = tid & 63;
        l1_on = lt < NKG1 * JG1;
        k1 = (lt % NKG1) * 4; j1 = (lt / NKG1) * TJ1;                          // dW1t[k < KP][j < 32]
        k2 = (lt & 7) * 4; j2 = (lt >> 3) * 8;                                 // dW2t[k < 32][j < 64]
        k3 = (lt & 15) * 4; j3 = (lt >> 4) * 8;                                // dW3t[k < 64][j < 32]
        k4 = lt & 31; j4 = (lt >> 5) * 2;                                      // dW4t[k < 32][j < 4]
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int j = 0; j < TJ1; ++j) g1[kk][j] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { g2[kk][j] = 0.f; g3[kk][j] = 0.f; }
        }
        g4[0] = g4[1] = gb4[0] = gb4[1] = gbias[0] = gbias[1] = 0.f;
    }
    // layer 4: dW4t[k][j], b4[j] from (a3, dz) of the rows [s0, s0 + ns)
    __device__ __forceinline__ void layer4(const float *__restrict__ rows, int row, int s0, int ns) {
#pragma unroll 4
        for (int s = s0; s < s0 + ns; ++s) {
            const float *r = rows + (size_t)s * row;
            const float a = r[A3 + k4];
            const float2 d = *reinterpret_cast<const float2 *>(r + D4 + j4);
            g4[0] = fmaf(a, d.x, g4[0]); g4[1] = fmaf(a, d.y, g4[1]);
            gb4[0] += d.x; gb4[1] += d.y;
        }
    }
    __device__ __forceinline__ void layer3(const float *rows, int row, int s0, int ns) { wgrad_tile<8>(g3, rows, row, A2 + k3, A3 + j3, s0, ns); }
    __device__ __forceinline__ void layer2(const float *rows, int row, int s0, int ns) { wgrad_tile<8>(g2, rows, row, A1 + k2, A2 + j2, s0, ns); }
    // layer 1 and the bias gradients b1 | b2 | b3 (the deltas sit in a1, a2, a3 of every row: 128 consecutive columns)
    __device__ __forceinline__ void layer1_and_biases(const float *__restrict__ rows, int row, int s0, int ns) {
        if (l1_on) wgrad_tile<TJ1>(g1, rows, row, X + k1, A1 + j1, s0, ns);
#pragma unroll 4
        for (int s = s0; s < s0 + ns; ++s) {
            const float2 d = *reinterpret_cast<const float2 *>(rows + (size_t)s * row + A1 + 2 * lt);
            gbias[0] += d.x; gbias[1] += d.y;
        }
    }
    __device__ __forceinline__ void emit(float *dst, bool add) const {
        if (l1_on)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int j = 0; j < TJ1; ++j) { float &q = dst[(k1 + kk) * H1 + j1 + j]; q = add ? q + g1[kk][j] : g1[kk][j]; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float &q2 = dst[off_w2(KP) + (k2 + kk) * H2 + j2 + j]; q2 = add ? q2 + g2[kk][j] : g2[kk][j];
                float &q3 = dst[off_w3(KP) + (k3 + kk) * H3 + j3 + j]; q3 = add ? q3 + g3[kk][j] : g3[kk][j];
            }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float &q4 = dst[off_w4(KP) + k4 * OP + j4 + j]; q4 = add ? q4 + g4[j] : g4[j];
            if (k4 == 0) { float &qb = dst[off_b4(KP) + j4 + j]; qb = add ? qb + gb4[j] : gb4[j]; }
            const int c = 2 * lt + j;                   // column of (b1 | b2 | b3)
            const int o = (c < H1) ? (off_b1(KP) + c) : ((c < H1 + H2) ? (off_b2(KP) + c - H1) : (off_b3(KP) + c - H1 - H2));
            float &qq = dst[o]; qq = add ? qq + gbias[j] : gbias[j];
        }
    }
    // add the NG groups through `scratch` (>= net_params floats of shared memory no longer in use) and write the CTA's
    // partial gradient, loss and (critic) advantage statistics; `red` holds one double per warp of the CTA
    template <int HEAD>
    __device__ __forceinline__ void finish(float *scratch, float *__restrict__ gpartial, double *__restrict__ lpartial, const LossArgs &la,
                                           const LossAcc &acc, double *red) const {
        constexpr int NPAR = net_params(KP), NT = 64 * NG;
        const int tid = threadIdx.x;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {                  // fixed order: deterministic
            if (half == g) emit(scratch, g > 0);
            __syncthreads();
        }
        float *gp = gpartial + (size_t)blockIdx.x * NPAR;
        for (int i = tid; i < NPAR; i += NT) gp[i] = scratch[i];
        const double lt_sum = team_sum(acc.loss, red, tid, NT, 0);
        if (tid == 0) lpartial[blockIdx.x] = lt_sum;
        if (HEAD == 0 && la.spartial) {
            const double t0 = team_sum(acc.sA, red, tid, NT, 0), t1 = team_sum(acc.sAA, red, tid, NT, 0), t2 = team_sum(acc.cnt, red, tid, NT, 0);
            if (tid == 0) { la.spartial[blockIdx.x * 3 + 0] = t0; la.spartial[blockIdx.x * 3 + 1] = t1; la.spartial[blockIdx.x * 3 + 2] = t2; }
        }
    }
};

// Per tile: [thread = R samples] forward + loss + dz; then, layer by layer from the top, [register tiles] weight gradient from
// (input activation, delta) and [thread = R samples] backward-data which overwrites the activation with its delta.
constexpr int kTeams = 1;
template <int KP, int HEAD>
__global__ void __launch_bounds__(kMlpBlock, 1) k_ppo_grad(SampleSet ss, const float *__restrict__ net, LossArgs la,
                                                           float *__restrict__ gpartial /* [grid][net_params] */,
                                                           double *__restrict__ lpartial /* [grid] */) {
    resolve_adv_stats(la);
    extern __shared__ __align__(16) float smem[];
    typedef GradCfg<KP> G;
    constexpr int R = G::R, ROW = G::ROW;
    float *sw = smem;                                  // flat net (transposed weights)
    float *W2 = sw + G::NP4;                           // [H2][H1] original orientation, for backward-data
    float *W3 = W2 + H2 * H1;                          // [H3][H2]
    float *W4 = W3 + H3 * H2;                          // [OP][H3]
    float *rows = W4 + OP * H3;
    __shared__ double red[kMlpBlock / 32];
    stage_net<KP>(sw, net);
    __syncthreads();
    for (int i = threadIdx.x; i < H1 * H2; i += blockDim.x) { const int k = i / H2, j = i % H2; W2[j * H1 + k] = sw[off_w2(KP) + i]; }
    for (int i = threadIdx.x; i < H2 * H3; i += blockDim.x) { const int k = i / H3, j = i % H3; W3[j * H2 + k] = sw[off_w3(KP) + i]; }
    for (int i = threadIdx.x; i < H3 * OP; i += blockDim.x) { const int k = i / OP, j = i % OP; W4[j * H3 + k] = sw[off_w4(KP) + i]; }
    __syncthreads();

    const int tid = threadIdx.x;
    float *my = rows + (size_t)tid * ROW;              // row of this thread's first sample; the r-th is kMlpBlock rows further
    constexpr int RS = kMlpBlock * ROW;
    WgradAcc<KP> wg;
    wg.init(tid);
    constexpr int HALF = G::TILE / 2;
    const int s0 = wg.half * HALF;
    LossAcc acc;

    for (int64_t base = (int64_t)blockIdx.x * G::TILE; base < ss.Q; base += (int64_t)gridDim.x * G::TILE) {
        bool sel[R];
        int64_t sidx[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            sidx[r] = 0;
            sel[r] = map_sample(ss, base + tid + r * kMlpBlock, sidx[r]);
            float *x = my + r * RS;
#pragma unroll
            for (int k = 0; k < KP; k += 4) {
                float4 v;
                v.x = (sel[r] && k + 0 < ss.D) ? ss.x[(int64_t)(k + 0) * ss.S + sidx[r]] : 0.f;
                v.y = (sel[r] && k + 1 < ss.D) ? ss.x[(int64_t)(k + 1) * ss.S + sidx[r]] : 0.f;
                v.z = (sel[r] && k + 2 < ss.D) ? ss.x[(int64_t)(k + 2) * ss.S + sidx[r]] : 0.f;
                v.w = (sel[r] && k + 3 < ss.D) ? ss.x[(int64_t)(k + 3) * ss.S + sidx[r]] : 0.f;
                st4(x + k, v);
            }
        }
        dense_fwd_t<KP, H1, true, R>(my + G::X, my + G::A1, RS, sw, sw + off_b1(KP));
        dense_fwd_t<H1, H2, true, R>(my + G::A1, my + G::A2, RS, sw + off_w2(KP), sw + off_b2(KP));
        dense_fwd_t<H2, H3, true, R>(my + G::A2, my + G::A3, RS, sw + off_w3(KP), sw + off_b3(KP));
        dense_fwd_t<H3, OP, false, R>(my + G::A3, my + G::D4, RS, sw + off_w4(KP), sw + off_b4(KP));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float *row = my + r * RS;
            float dz[OP] = {0.f, 0.f, 0.f, 0.f};
            if (sel[r]) ppo_loss<HEAD>(la, sidx[r], load_loss_in<HEAD>(la, sidx[r]), ld4(row + G::D4), dz, acc);
            // an unselected row has x = 0 but its activations are relu(bias) != 0: its dz = 0 keeps every gradient term zero
            st4(row + G::D4, make_float4(dz[0], dz[1], dz[2], dz[3]));
        }
        __syncthreads();
        wg.layer4(rows, ROW, s0, HALF);
        __syncthreads();
        dense_bwd_t<H3, OP, R>(my + G::A3, my + G::D4, RS, W4);           // a3 now holds delta3
        __syncthreads();
        wg.layer3(rows, ROW, s0, HALF);
        __syncthreads();
        dense_bwd_t<H2, H3, R>(my + G::A2, my + G::A3, RS, W3);           // a2 now holds delta2
        __syncthreads();
        wg.layer2(rows, ROW, s0, HALF);
        __syncthreads();
        dense_bwd_t<H1, H2, R>(my + G::A1, my + G::A2, RS, W2);           // a1 now holds delta1
        __syncthreads();
        wg.layer1_and_biases(rows, ROW, s0, HALF);
        __syncthreads();
    }
    wg.template finish<HEAD>(rows, gpartial, lpartial, la, acc, red);
}

// ---- 4. deterministic reduction of the per-CTA partials ----------------------------------------------
__global__ void __launch_bounds__(256) k_reduce_partials(const float *__restrict__ gpartial, int nblocks, int npar, float *__restrict__ grad) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= npar) return;
    float acc = 0.f;
    for (int b = 0; b < nblocks; ++b) acc += gpartial[(size_t)b * npar + i];
    grad[i] = acc;
}
// gradient partials, loss partials and (optionally) the advantage-statistics partials of one ppo_grad launch in ONE kernel
__global__ void __launch_bounds__(256) k_reduce_all(const float *__restrict__ gpartial, int nblocks, int npar, float *__restrict__ grad,
                                                    const double *__restrict__ lpartial, double *__restrict__ loss,
                                                    const double *__restrict__ spartial, double *__restrict__ stats) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < npar) {
        float acc = 0.f;
        for (int b = 0; b < nblocks; ++b) acc += gpartial[(size_t)b * npar + i];
        grad[i] = acc;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= 252) {               // four idle-ish lanes of the last block: the scalars
        const int j = threadIdx.x - 252;                                    // 0: loss, 1..3: sum A, sum A^2, n
        if (j == 0) { double acc = 0.0; for (int b = 0; b < nblocks; ++b) acc += lpartial[b]; *loss = acc; }
        else if (stats) { double acc = 0.0; for (int b = 0; b < nblocks; ++b) acc += spartial[(size_t)b * 3 + (j - 1)]; stats[j - 1] = acc; }
    }
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

// both Adam steps of one epoch (actor and critic of a pair act on different nets, PY:810-815) in one launch
struct AdamArgs { float *p; const float *g; float *m, *v; int n; float step_size, inv_sqrt_bc2; };
__global__ void __launch_bounds__(256) k_adam2(AdamArgs a0, AdamArgs a1, float beta1, float beta2, float eps) {
    const AdamArgs a = blockIdx.y ? a1 : a0;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n) return;
    const float gi = a.g[i];
    const float mi = beta1 * a.m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * a.v[i] + (1.0f - beta2) * gi * gi;
    a.m[i] = mi; a.v[i] = vi;
    a.p[i] = a.p[i] - a.step_size * (mi / (sqrtf(vi) * a.inv_sqrt_bc2 + eps));
}

}  // namespace mhppo
