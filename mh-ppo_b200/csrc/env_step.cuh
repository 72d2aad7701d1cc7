// env_step.cuh -- the per-thread body of the env-step kernel (one thread = one env).
//
// Replaces Crosswalk_hybrid_multi_*.step (SC:789-878 and siblings) + in-kernel auto-reset.
//
// Shape of the computation (chosen for the SM, not copied from the Python control flow):
//  * the kernel is bound by latency, not by HBM and not by issue slots: a warp walks ~8k warp-level
//    instructions per step (976 bytes of state per env), an SM holds 20 warps (registers), and every
//    instruction costs ~7 cycles of its warp's life whether 3 or 32 lanes are active
//    (profiles/round2_env_step_notes.md).  What counts is (a) how many warps an SM can hold, (b) how
//    much SASS a warp walks per step -- instruction-cache misses were the second-largest stall of the
//    unrolled versions (profiles/round1_env_step_r1d_summary.txt) -- and (c) that lanes which take the
//    same rare path take it TOGETHER (one reconvergence point in front of every expensive block).
//    So the CARS of a CTA live in shared memory ([field][slot][thread], conflict-free) and every
//    loop over cars is ROLLED: one copy of each (pedestrian, car) body in SASS, few live
//    registers, 5 CTAs per SM instead of 3;
//  * pedestrians are STREAMED: one rolled loop loads pedestrian j from HBM, runs its state
//    machine, its danger detection against every car, its contribution to the wait rewards and
//    its observation row, and stores it back.  The reference runs these as four separate passes
//    over all pedestrians (SC:808-809, 841-846, 849-858, 869); the passes only communicate through
//    per-car running min/max and per-pedestrian fields, so interleaving them per pedestrian is
//    exactly equivalent (DESIGN.md "Step kernel");
//  * the coldest paths are single out-of-line copies: the episode reset, the Philox block function and
//    the exact fp64 critical gap (the sin walking profile and the gap judgement are inlined: with rolled
//    loops they have one or two call sites and the call's register shuffling cost more than it saved).  The gap-acceptance decision
//    `choix_pedestrian` runs every step for a waiting pedestrian, so its log10/pow/normal-draw
//    comparison is a filtered predicate (fp32 with an error bound, fp64 only when undecidable).
//    (Tried and dropped: queueing the judgements of a warp in shared memory and evaluating them one
//    per lane.  With rolled car loops the lanes already meet at a single call site; it was 3 % slower.);
//  * reward-shaping exponentials feed only fp32 outputs: their argument is formed in fp64 and the
//    exponential itself is exp2f (DESIGN.md "Precision"); everything that feeds state or a
//    threshold stays fp64 in the reference's operation order.
#pragma once
#include "../../include/mhppo.h"
#include "env_state.cuh"

namespace mhppo {

struct RngKey { uint32_t k0, k1; int64_t env_id0; };

// One-ahead prefetch of the next state group.  The loops over cars / pedestrians are rolled, so without it
// every iteration exposes a full HBM round trip (30 % of the warp stall samples sat on the first use of
// those loads, profiles/round2_env_step_notes.md); fetching group i+1 into L2 while group i is computed
// took 7 % off the step (L1 as the target measured the same; prefetching everything at kernel entry added nothing).
#ifndef MH_ENTRY_PF
#define MH_ENTRY_PF 1
#endif
#ifndef MH_PREFETCH
#define MH_PREFETCH 2
#endif
MH_HD void prefetch_next(const void *p) {
#if defined(__CUDA_ARCH__) && MH_PREFETCH == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#elif defined(__CUDA_ARCH__) && MH_PREFETCH == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

struct StepIO {
    mhppo_view actions, obs, rewards, reward_light, term_obs;
    uint8_t *done;
    int autoreset;
    int64_t n_begin, n_end;        // env range of this launch ([0, N) except for the sliced host-buffer path)
};

// ---------------------------------------------------------------------------------------------
// gap-acceptance decision

// Exact critical gap of pedestrian.CG_score (SC:419-428) from the two uniforms of its normal draw.
// Out of line: only reached when the fp32 filter below cannot decide (~1e-4 of the draws).
static MH_NOINLINE double cg_exact(double v0y, int gender, int age, double size, double u1, double u2) {
    double lv = 0.09 + log10(size / fabs(v0y + 10e-3));
    lv = lv + 0.0369 * (double)(gender == 1);
    lv = lv + -0.0355 * (double)(age == 0);
    lv = lv + -0.0221 * (double)(age == 1);
    lv = lv + -0.1810 * (double)(age == 2);
    lv = lv + (0.0 + 0.09 * (sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2)));
    return pow(10.0, lv);
}

// `car_time + light < CG` of SC:167-170, bit-exact at fp32 cost: the comparison is first evaluated
// in fp32 with a conservative error bound (filtered predicate); only an undecidable case recomputes
// both sides in fp64 exactly as the reference does.  `ctr` is the Philox block of this decision's
// normal draw (one block per CG_score call).  Inlined at its two call sites (3 % faster than one out-of-line copy:
// a 12-argument call costs more register moves than the body's second copy costs in instruction cache).
//
// Three levels, each sound on its own (mhppo_gap_selftest compares all of them with the exact decision on adversarial inputs):
//  0. no draw: z = sqrt(-2 ln(1 - u1)) cos(2 pi u2) with 1 - u1 >= 2^-53, so |z| <= 8.572 for EVERY block and CG lies in
//     base * 10^(+-0.7715) = base * [0.1692, 5.909].  A car that stands (lhs ~ dx / 0.01) or is about to pass is decided
//     here with a 3 % margin for the fp32 evaluation; the caller has consumed the draw either way;
//  1. fp32 draw from the top 24 bits of the two uniforms with the SFU intrinsics (device build).  Error budget of z:
//     1 - u1 truncated to 24 bits is exact in fp32 (a multiple of 2^-24 in (0, 1]) and off by < 2^-24; __logf is within
//     2^-21.4 absolute on [0.5, 2] (3 ulp elsewhere), so |d ln| <= 4.8e-7 and |d sqrt(-2 ln)| <= sqrt(2 * 4.8e-7) = 9.8e-4
//     (the worst case, 1 - u1 ~ 1; for 1 - u1 < 0.5 the error is <= 2^-24 / (1 - u1) / r <= 6e-5 down to the cut-off
//     1 - u1 >= 2^-12, below which this level is skipped); the cosine adds 1.1e-6 * r.  CG = base * 10^(0.09 z), so
//     |d CG| / CG <= 0.2072 * 9.8e-4 + 3e-6 = 2.06e-4 against the 5e-4 of `tol`;
//  2. exact: both sides in fp64 in the reference's operation order.
MH_HD bool gap_eval(double dx, double vden, double light, double size, double v0y, int gender, int age,
                                 uint32_t ctr, uint32_t env_lo, uint32_t env_hi, uint32_t k0, uint32_t k1, int *level = nullptr) {
    const double pden = fabs(v0y + 10e-3);
    // fp32 estimate: CG = size/|v0y+.01| * 10^(0.09 + gender/age terms + 0.09*z)
    const float adj = 0.09f + ((gender == 1) ? 0.0369f : 0.f) + ((age == 0) ? -0.0355f : ((age == 1) ? -0.0221f : -0.1810f));
    const float base = ((float)size / (float)pden) * exp2f(3.3219280948873623f * adj);            // CG at z = 0
    const float lhs = fabsf((float)dx / (float)vden) + (float)light;
    if (level) *level = 0;
    if (lhs > base * 6.09f) return false;
    if (lhs < base * 0.1642f) return true;
    const PhiloxBlock b = philox4x32_10(ctr, 0u, env_lo, env_hi, k0, k1);
    if (level) *level = 1;
#ifdef __CUDA_ARCH__
    const float wf = 1.0f - (float)(b.w0 >> 8) * 5.9604644775390625e-8f;                          // 1 - u1 from the top 24 bits: exact in fp32
    if (wf >= 2.44140625e-4f) {
        const float u2f = (float)(b.w2 >> 8) * 5.9604644775390625e-8f;
        const float cz = -__cosf(6.2831853f * (u2f - 0.5f));                                      // cos(2 pi u2), argument in [-pi, pi)
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(wf)));
        const float cg = base * exp2f(3.3219280948873623f * (0.09f * (r * cz)));
        const float tol = 5e-4f * (fabsf(lhs) + cg) + 1e-30f;
        if (fabsf(lhs - cg) > tol) return lhs < cg;
    }
    const double u1 = u53(b.w0, b.w1), u2 = u53(b.w2, b.w3);
#else
    const double u1 = u53(b.w0, b.w1), u2 = u53(b.w2, b.w3);
    {
        const float z = sqrtf(-2.0f * logf((float)(1.0 - u1))) * cosf(6.2831853f * (float)u2);
        const float cg = base * exp2f(3.3219280948873623f * (0.09f * z));
        const float tol = 2e-4f * (fabsf(lhs) + cg) + 1e-30f;
        if (fabsf(lhs - cg) > tol) return lhs < cg;
    }
#endif
    if (level) *level = 2;
    return (fabs(dx / vden) + light) < cg_exact(v0y, gender, age, size, u1, u2);
}

// ---------------------------------------------------------------------------------------------
// The cars of NT envs (one CTA; NT = 1 in the host build), [field][slot][thread].  Positions and
// speeds are the fp64 values after this step's move (they are rounded to fp32 only when stored);
// brake = Vc^2/(2b) and rVc = 1/Vc are the two per-car quotients every (pedestrian, car) pair reuses.
template <int MC, int NT>
struct CarSlots {
    double Sc[MC][NT], Vc[MC][NT], brake[MC][NT], rVc[MC][NT];
    float prevSc[MC][NT], light[MC][NT], pa[MC][NT], es[MC][NT], Ts[MC][NT], rl[MC][NT], wmin[MC][NT], wtmp[MC][NT];
    uint32_t bits[MC][NT];                                      // line | exist << 8
};
#define MH_LINE(S, i, t) ((int)((S).bits[i][t] & 255u))
MH_HD int ctz32(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

// pedestrian.choix_pedestrian, SC:139-174 / NA:138-177.  `seen` = slots handed to pedestrian.step
// (existing cars in scalable SC:803-806, leaders+followers in 4cars C4:796-799, all otherwise).
//
// The decision is split in two: choix_plan does everything that needs no normal draw (geometry,
// the follower rules, the shuffle's draws) and returns either the answer or the ordered list of
// cars whose gap must be judged (choix_fast then judges them in order; every lane of a warp reaches
// that one rolled loop, so lanes with work run it together).  SC:162-173 walks the cars in front in ascending order:
// a car standing on the crosswalk answers False, an upstream car answers False if its gap is
// refused; the two cases exclude each other (on the crosswalk means Sc > Sp_x), so the walk is
// "judge the upstream cars that come before the first car on the crosswalk, in order; False at the
// first refusal, else False if there was a car on the crosswalk, else True".
struct ChoixPlan { uint32_t gaps; int answer; bool tail_false; };   // answer: -1 = needs the gaps, else 0 / 1
template <int V, int MC, int NT>
MH_HD ChoixPlan choix_plan(const EnvConst &c, const Geo &g, const PedR &p, const CarSlots<MC, NT> &S, int t, uint32_t seen, Rng &rng) {
    typedef VT<V> T;
    ChoixPlan pl; pl.gaps = 0; pl.answer = -1; pl.tail_false = false;
    const int nseen = __builtin_popcount(seen);
    // is_in_front / is_crossing_in_front depend on the car only through its lane: one evaluation per lane
    uint32_t inf1_l = 0, cif_l = 0;
    const bool neg = p.dir == -1;
    const double y = neg ? -p.Spy : p.Spy;
#pragma unroll 1
    for (int l = 0; l < c.L; ++l) {
        const LanePred lp = lane_pred(g, y, neg, c.L, l, 1.0, 0.5);
        inf1_l |= (lp.in_front ? 1u : 0u) << l;
        cif_l |= (lp.crossing ? 1u : 0u) << l;
    }
    uint32_t inf1 = 0, on_cross = 0, behind = 0, blocked = 0;
#pragma unroll 1
    for (int i = 0; i < c.nC; ++i) {
        const double Sc = S.Sc[i][t];
        const int line = MH_LINE(S, i, t);
        const uint32_t f1 = (inf1_l >> line) & 1u;
        const uint32_t over = ((Sc < 4.0 + p.Spx) && (Sc > p.Spx)) ? 1u : 0u;        // car body on the crosswalk
        inf1 |= f1 << i;
        on_cross |= over << i;
        behind |= ((Sc < p.Spx) ? 1u : 0u) << i;
        blocked |= (over & f1 & ((cif_l >> line) & 1u)) << i;
    }
    inf1 &= seen; on_cross &= seen; behind &= seen; blocked &= seen;
    if (p.fl & PF_FOLLOW) {
        // NA:151 shuffles the visiting order for real, but both loops of NA:153-158 only ever return
        // False, so the order cannot change the outcome: only the n-1 draws are consumed.  SC:152-153
        // shuffles a temporary.
        if ((T::naif || T::burn_shuffle) && nseen > 1) rng.skip(nseen - 1);
        if (blocked) { pl.answer = 0; return pl; }                                   // SC:154-158 / NA:153-155
        for (uint32_t m = behind; m; m &= m - 1u) {                                  // ascending slots
            const float light = S.light[ctz32(m)][t];
            if (T::naif) { if (light < 0.f) { pl.answer = 0; return pl; } }          // NA:156-158
            else if (light != 0.f) { pl.answer = light > 0.f ? 1 : 0; return pl; }   // SC:159-161: first upstream car with a light
        }
    }
    const uint32_t oc = inf1 & on_cross;                                             // SC:162-173
    const uint32_t before = oc ? ((oc & (0u - oc)) - 1u) : 0xFFFFFFFFu;              // slots below the first car on the crosswalk
    pl.gaps = inf1 & behind & before;
    pl.tail_false = oc != 0;
    if (!pl.gaps) pl.answer = pl.tail_false ? 0 : 1;
    return pl;
}
// one judgement, sequential form: a pedestrian handed to choix_pedestrian is always `crossing`, so
// CG_score always draws (the CG = 0 branch of SC:427-428 is unreachable from here)
template <int MC, int NT>
MH_HD bool gap_refused(const Geo &g, const PedR &p, const CarSlots<MC, NT> &S, int t, int i, Rng &rng) {
    const double size = fabs((double)(p.lpos - MH_LINE(S, i, t))) * g.cross;
    const uint32_t ctr = rng.ctr++;
    return gap_eval(S.Sc[i][t] - p.Spx, S.Vc[i][t] + 10e-3, (double)S.light[i][t], size, p.v0y, p.gender, p.age, ctr,
                    rng.env_lo, rng.env_hi, rng.k0, rng.k1);
}
template <int V, int MC, int NT>
MH_HD bool choix_fast(const EnvConst &c, const Geo &g, const PedR &p, const CarSlots<MC, NT> &S, int t, uint32_t seen, Rng &rng) {
    const ChoixPlan pl = choix_plan<V, MC, NT>(c, g, p, S, t, seen, rng);
    if (pl.answer >= 0) return pl.answer != 0;
    for (uint32_t m = pl.gaps; m; m &= m - 1u)
        if (gap_refused<MC, NT>(g, p, S, t, ctz32(m), rng)) return false;
    return !pl.tail_false;
}

struct WalkOut { double pos, spd; };
// pedestrian.new_pedestrian_sin_y, SC:436-443 with the parameters of SC:94-102
MH_HD WalkOut walk_sin(double W, double v0y, double Spy, double dt, int step, int t0c, int dir) {
    const double PI_ = 3.141592653589793, Vm = 2.5;
    const double av = fabs(v0y);
    const double T = W / (av + 10e-3);
    const bool check = ((av * PI_) / 2.0 <= Vm);
    constexpr double kAd = 1.0 - (2.0 / 3.141592653589793);
    const double A = check ? (PI_ * av / 2.0 + 0.0) : (0.0 + div_rcp(Vm - av, kAd, 1.0 / kAd));
    const double B = check ? 0.0 : (Vm - A);
    const double w = PI_ / T;
    const double t = (double)step * dt + dt, t0 = (double)t0c * dt;
    double sn, cs;
    sincos(w * (t - t0), &sn, &cs);
    const double speed_p = A * sn + B;
    const double pos_p = (-W) / 2.0 + (A * (-cs + 1.0) / w);
    WalkOut o;
    if (pos_p >= 0.0 && speed_p < av) { o.pos = Spy + v0y * dt; o.spd = v0y; }
    else { o.pos = (double)dir * pos_p; o.spd = (double)dir * speed_p; }
    return o;
}

struct ViewOut {
    float *p; int64_t cs;
    MH_HD void operator()(int k, float v) const { p[(int64_t)k * cs] = v; }
};
struct NullOut {
    MH_HD void operator()(int, float) const {}
};

// reset (SC:884-946) of env n + first observation + store; used by the reset kernel and by the
// auto-reset branch of the step kernel.  Works on a scratch copy in local memory.
template <int V, int MC, int MP>
MH_NOINLINE void reset_and_store(const EnvArena &a, const EnvConst &c, int64_t n, Rng rng, mhppo_view obs) {
    EnvR<MC, MP> e;
    e.rng = rng;
    reset_env<V, MC, MP>(c, e);
    if (obs.ptr) { ViewOut out{obs.ptr + n * obs.env_stride, obs.comp_stride}; write_obs<V, MC, MP>(c, e, true, out); }
    else { NullOut nul; write_obs<V, MC, MP>(c, e, true, nul); }   // get_data still updates the running-min delta
    store_env<MC, MP>(a, c, n, e);
}

MH_HD float exp_f32(double x) {   // e^x for output-only terms: fp64 argument, fp32 exponential
    return exp2f((float)(x * 1.4426950408889634));
}
// 2^x of the reward-shaping terms: the bare MUFU.EX2 (results below 2^-126 flush to zero: an absolute difference of 1e-38 on
// an fp32 reward term); exp2f() wraps the same instruction in a three-instruction range fix-up, three times per (pedestrian, car) pair
MH_HD float exp2_shape(float x) {
#ifdef __CUDA_ARCH__
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}

// ---------------------------------------------------------------------------------------------
// pedestrian.step, SC:297-417, for the pedestrian currently held in registers
// boolean_ped_position (SC:266-275); returns whether this step opens with the kerb decision of SC:311-318
MH_HD bool ped_flags(const Geo &g, PedR &p) {
    const double dy = (double)p.dir * p.Spy;
    p.fl &= ~(PF_LEFT | PF_IN_CROSS);
    if (dy >= g.Hp) p.fl |= PF_LEFT;
    else if (dy > g.Hn) p.fl |= PF_IN_CROSS;
    return (p.fl & PF_CROSSING) && !(p.fl & PF_DECISION) && (p.fl & PF_AT_CROSSING);
}

// `kerb_choice`: the outcome of choix_pedestrian for a pedestrian whose step opens with the kerb decision
// (already taken by the caller, draws already consumed); ignored otherwise
template <int V, int MC, int NT>
MH_HD void ped_step_stream(const EnvConst &c, const Geo &g, PedR &p, const CarSlots<MC, NT> &S, int t, uint32_t seen, int step,
                           Rng &rng, bool kerb_choice) {
    typedef VT<V> T;
    const double dt = c.dt;
    const double pp_y = p.Spy + p.v0y * dt;                                          // SC:298
    const double dy = (double)p.dir * p.Spy;
    if (!(p.fl & PF_CROSSING)) return;                                               // SC:308

    const bool first_decision = !(p.fl & PF_DECISION) && (p.fl & PF_AT_CROSSING);    // SC:311
    // kerb arrival is evaluated after the decision block in the reference; it cannot fire in a step
    // that took the decision block (decision is then True), so it is decided up front
    const bool arrive = !first_decision && (dy < g.Hn) && (pp_y * (double)p.dir > g.Hn) && !(p.fl & PF_DECISION);  // SC:320
    const bool in_walk_block = !arrive && (first_decision || (fabs(p.Spy) <= g.Hp) || (p.fl & PF_DECISION));      // SC:331

    // phase A: decide whether this step needs the walking model / a gap-acceptance decision.  The bookkeeping of the three
    // cheap outcomes (decision taken, standing still, decision refused) is written as selects: with branches the lanes that
    // took the kerb decision and the lanes already walking reached the uniform draw and the walking model below as two
    // groups, and both ran twice per warp.
    const bool accept = first_decision && kerb_choice;                               // SC:311-318 (first_decision implies in_walk_block)
    const bool stopped = in_walk_block && (p.tstop != 0);                            // SC:335-339
    // a refused kerb decision (every step of a waiting pedestrian): the uniform of SC:346 and the randint of SC:406 are
    // drawn, but `u < 0.98 and choice` is False whatever u is and SC:410 overwrites t_stop -- the two blocks are
    // consumed without being computed
    const bool refuse = first_decision && !kerb_choice && !stopped;
    const bool draw = in_walk_block && !stopped && !refuse;
    {
        const bool still = stopped || refuse;
        uint32_t fl = p.fl;
        fl = accept ? (fl & ~PF_AT_CROSSING) : fl;
        fl = (first_decision && !refuse) ? (fl | PF_DECISION) : fl;                  // SC:317, undone by SC:408 when refused
        p.fl = fl;
        p.lpos = accept ? ((p.dir < 0) ? (c.L - 1) : 0) : p.lpos;
        p.t0c = (first_decision ? step : p.t0c) + (still ? 1 : 0);
        p.tstop -= stopped ? 1 : 0;
        p.waitc += refuse ? 1 : 0;                                                   // SC:408-411 (t_stop is already 0)
        if (refuse) rng.skip(2);
        p.Vpx = still ? 0.0 : p.Vpx; p.Vpy = still ? 0.0 : p.Vpy;
    }
    bool walk_try = false;
    if (draw) {
        const double u = rng.random();                                               // SC:346
        if (u < 0.98) walk_try = true;
        else {                                                                       // SC:405-413
            p.tstop = rng.randint(T::rs_lo, T::rs_hi);
            p.Vpx = 0.0; p.Vpy = 0.0; p.t0c += 1;
        }
    }
    if (walk_try) {                                                                  // SC:347-402
        p.fl &= ~PF_DECISION;
        double ny, nv;
        if (c.sin_model) { const WalkOut wo = walk_sin(g.W, p.v0y, p.Spy, dt, step, p.t0c, p.dir); ny = wo.pos; nv = wo.spd; }
        else { ny = p.Spy + p.v0y * dt; nv = p.v0y; }                                // SC:433-434
        bool change_line = false;                                                    // will_change_line SC:281-286
        if (fabs(ny) < g.Hp) {
            const double nl = floor((ny + g.Hp) / g.cross);
            if (nl != (double)p.lpos && fabs(p.Spy) < g.Hp) change_line = true;
        }
        double dtc = ((double)(c.L - p.lpos - 1) * g.cross) * (double)(p.dir > 0);   // SC:350-351
        dtc += ((double)p.lpos * g.cross) * (double)(p.dir < 0);
        bool new_choice = false;
        if (change_line && (dtc > 0.0 && dtc < g.W)) {                               // SC:353-357
            new_choice = choix_fast<V, MC, NT>(c, g, p, S, t, seen, rng);
            if (new_choice) p.fl &= ~PF_STOP;
        }
        if (p.fl & PF_STOP) {                                                        // SC:363-369
            p.Vpx = 0.0; p.Vpy = 0.0; p.t0c += 1;
            if (change_line) p.waitc += 1;
        } else if (T::has_need_to_stop && (p.fl & PF_NEED_STOP) && p.Spy < p.cstop && pp_y > p.cstop) {
            p.tstop = rng.randint(T::nts_lo, T::nts_hi);                             // SC:371-380
            p.fl &= ~PF_NEED_STOP;
            p.Vpx = 0.0; p.Vpy = 0.0; p.t0c += 1;
        } else if (!change_line || new_choice) {                                     // SC:382-387
            const double ratio = p.v0x / ((p.fl & PF_RATIO_EPS) ? (p.v0y + 1e-3) : p.v0y);   // SC:73 (SC:111 after reset_ped)
            p.Spy = ny; p.Vpy = nv;
            p.Spx = p.Spx + p.Vpy * ratio * dt;
            p.Vpx = p.Vpy * ratio;
            p.crossc += 1;
            // apply_change_line SC:288-294: change_line implies |ny| < Hp here, so its kerb branch is dead.  (Carrying the
            // quotient of will_change_line down here instead of dividing again costs a register the kernel does not have: it spilled.)
            if (change_line) p.lpos = (int)floor((ny + g.Hp) / g.cross);
        } else {                                                                     // SC:389-397
            p.fl |= PF_STOP;
            const double d = fabs(((double)p.dir * (g.W - dtc) - (double)p.dir * g.W / 2.0) - p.Spy);
            const double px = p.Vpx * d / fabs(p.Vpy + 10e-3);
            p.Vpx = div_rcp(px, dt, c.rdt);
            p.Spx = p.Spx + px;
            p.Vpy = div_rcp((double)p.dir * d, dt, c.rdt);
            p.Spy = (double)p.dir * ((g.W - dtc) - g.W / 2.0);
        }
    } else if (arrive) {                                                             // SC:320-328
        const double px = (p.Vpx * dt) * (fabs(g.Hn - dy) / fabs(p.Vpy * dt + 10e-3));
        p.Vpx = div_rcp(px, dt, c.rdt);
        p.Spx = p.Spx + px;
        p.Vpy = div_rcp((double)p.dir * fabs(-dy - g.Hp), dt, c.rdt);
        p.Spy = (double)(-p.dir) * g.W / 2.0;
        p.tstop = 0;
        p.fl |= PF_AT_CROSSING;
    } else if (!in_walk_block) {                                                     // SC:415-417
        p.Spx = p.Spx + p.v0x * dt; p.Vpx = p.v0x;
        p.Spy = p.Spy + p.v0y * dt; p.Vpy = p.v0y;
    }
}

// ---------------------------------------------------------------------------------------------
template <int V, int MC, int MP, int NT>
MH_HD void env_step_thread(const EnvArena &a, const EnvConst &c, const RngKey &key, const StepIO &io, int64_t n,
                           CarSlots<MC, NT> &S, int t) {
    typedef VT<V> T;
#if MH_ENTRY_PF
    // the first car's rows and action words start their trip together with the env word (they do not depend on it): 2 % of the
    // step (prefetching the first pedestrian or the second car here as well measured the same)
    prefetch_next(&a.car_a[n]); prefetch_next(&a.car_b[n]);
    prefetch_next(io.actions.ptr + n * io.actions.env_stride);
    prefetch_next(io.actions.ptr + n * io.actions.env_stride + (int64_t)(c.nA / 2) * io.actions.comp_stride);
#endif
    // ---- env word
    const float4 ee = a.env_e[n];
    const double cross = words_to_double(f2u(ee.x), f2u(ee.y));
    const uint32_t tr = f2u(ee.z);
    int step = (int)(tr & 255u);
    const int ped_traffic = (int)((tr >> 8) & 255u), car_traffic = (int)((tr >> 16) & 255u);
    Rng rng;
    {
        const uint64_t gid = (uint64_t)(key.env_id0 + n);
        rng.ctr = f2u(ee.w); rng.env_lo = (uint32_t)gid; rng.env_hi = (uint32_t)(gid >> 32); rng.k0 = key.k0; rng.k1 = key.k1;
    }
    const Geo g = make_geo(cross, c.L);
    prefetch_next(&a.ped_a[n]); prefetch_next(&a.ped_b[n]); prefetch_next(&a.ped_c[n]);
    const bool done = (step >= c.done_idx) || (ped_traffic <= 0);                    // SC:874
    const bool will_reset = done && io.autoreset;
    const mhppo_view ov = will_reset ? io.term_obs : io.obs;
    float *const op = ov.ptr ? ov.ptr + n * ov.env_stride : nullptr;
    const int ocs = (int)ov.comp_stride;                 // the API rejects component strides >= 2^31 (one register instead of two)

    // ---- cars: load, act (SC:797-802 / C4:791-795 / C42:807-811 / CO:753-755), observation row
    // (car.get_data SC:652-655) and position/speed store, one slot at a time.  Leaders come before the
    // slots that follow them (odd slots in scalable SC:799-801, the second half in 4cars C4:793-795), so a
    // follower reads its leader's already-moved position from the slots.
    uint32_t seen = 0, lead_ok = 0, lead_green = 0;   // seen: cars handed to pedestrian.step; lead_ok: cars detection may touch
    {
        const float *ap = io.actions.ptr + n * io.actions.env_stride;
        const int half = c.nA / 2;
        const int64_t cs = io.actions.comp_stride;
#pragma unroll 1
        for (int i = 0; i < c.nC; ++i) {
            const float4 ka = a.car_a[(int64_t)i * a.N + n], kb = a.car_b[(int64_t)i * a.N + n];
            if (i + 1 < c.nC) {
                prefetch_next(&a.car_a[(int64_t)(i + 1) * a.N + n]); prefetch_next(&a.car_b[(int64_t)(i + 1) * a.N + n]);
                prefetch_next(&ap[(int64_t)(i + 1) * cs]); prefetch_next(&ap[(int64_t)(half + i + 1) * cs]);
            }
            CarR k;
            k.Vc = (double)ka.x; k.Sc = (double)ka.y; k.light = (double)ka.z; k.Ac = (double)ka.w;
            const uint32_t b = f2u(kb.w);
            k.line = (int)(b & 255u); k.exist = (int)((b >> 8) & 1u);
            if (T::scal) {
                double acc = (double)ap[(int64_t)i * cs];
                const int l = i > 0 ? i - 1 : 0;
                if ((i & 1) && ((S.bits[l][t] >> 8) & 1u) && k.exist) acc = dmin(idm(c, k, S.Sc[l][t], S.Vc[l][t]), acc);
                else acc = dmin(2.0, acc);
                car_move<V>(c, k, acc, (double)ap[(int64_t)(half + i) * cs]);
            } else if (T::four && i >= c.nb_car) {
                const int l = i - c.nb_car;                                          // follower of leader l
                const double a_idm = idm(c, k, S.Sc[l][t], S.Vc[l][t]);
                if (V == V_4CARS2) car_move<V>(c, k, dmin(a_idm, (double)ap[(int64_t)i * cs]), (double)ap[(int64_t)(half + i) * cs]);
                else car_move<V>(c, k, a_idm, (double)S.light[l][t]);
            } else {
                car_move<V>(c, k, (double)ap[(int64_t)i * cs], (double)ap[(int64_t)(half + i) * cs]);
            }
            S.prevSc[i][t] = ka.y; S.Sc[i][t] = k.Sc; S.Vc[i][t] = k.Vc; S.light[i][t] = (float)k.light;
            // Vc^2 / (2b): b = 4 in every reference driver, a power of two, where the product with the reciprocal IS the quotient
            S.brake[i][t] = (c.brake_inv != 0.0) ? (k.Vc * k.Vc) * c.brake_inv : div_pos(k.Vc * k.Vc, c.brake_den);
            // 1/Vc is only read for Vc >= 0.01 (SC:190 `Vc < 0.05`, SC:482 / ST:478 thresholds); below that a
            // finite stand-in keeps a stopped car (Vc = 0) out of the division's slow path
            S.rVc[i][t] = 1.0 / opaque((k.Vc >= 0.01) ? k.Vc : 1.0);
            S.pa[i][t] = kb.x; S.es[i][t] = kb.y; S.Ts[i][t] = kb.z; S.rl[i][t] = 0.f; S.wmin[i][t] = 0.f;
            S.bits[i][t] = b;
            const bool ex = !T::scal || k.exist;
            if (ex) seen |= 1u << i;
            if (i < c.nlead && ex) { lead_ok |= 1u << i; if (k.light > 0.0) lead_green |= 1u << i; }
            if (op) {
                float *q = op + (int64_t)(T::car_w * i) * (int64_t)ocs;
                q[(int64_t)0 * ocs] = ex ? (float)k.Ac : 0.f; q[(int64_t)1 * ocs] = ex ? (float)k.Vc : 0.f;
                q[(int64_t)2 * ocs] = ex ? (float)(10.0 - k.Vc) : 10.f; q[(int64_t)3 * ocs] = ex ? (float)k.Sc : -1000.f;
                q[(int64_t)4 * ocs] = ex ? (float)k.light : 0.f; q[(int64_t)5 * ocs] = (float)k.line;
                if (T::scal) q[(int64_t)6 * ocs] = ex ? 1.f : 0.f;
            }
            if (!will_reset) a.car_a[(int64_t)i * a.N + n] = make_float4((float)k.Vc, (float)k.Sc, (float)k.light, (float)k.Ac);
        }
    }
    const float green = (float)__builtin_popcount(lead_green);                       // SC:246
    const int env_w = T::scal ? 4 : 3;
    const int ped_o = T::car_w * c.nC + env_w;
    const double time_braking = c.time_braking;                                      // SC:580
    constexpr uint32_t ANY_EXIST = 1u << 31;             // "a pedestrian exists so far" rides in a spare bit of `lead_green`, whose readers mask it with car bits (no register of its own)

    // ---- pedestrians, streamed
#pragma unroll 1
    for (int j = 0; j < c.nP; ++j) {
        PedR p;
        {
            const float4 pa = a.ped_a[(int64_t)j * a.N + n], pb = a.ped_b[(int64_t)j * a.N + n], pc = a.ped_c[(int64_t)j * a.N + n];
            if (j + 1 < c.nP) {
                prefetch_next(&a.ped_a[(int64_t)(j + 1) * a.N + n]); prefetch_next(&a.ped_b[(int64_t)(j + 1) * a.N + n]);
                prefetch_next(&a.ped_c[(int64_t)(j + 1) * a.N + n]);
            }
            p.Spx = (double)pa.x; p.Spy = (double)pa.y; p.Vpx = (double)pa.z; p.Vpy = (double)pa.w;
            p.v0x = (double)pb.x; p.v0y = (double)pb.y; p.cstop = (double)pb.z; p.delta = (double)pb.w;
            p.wdl = (double)pc.x;
            unpack_ped_counts(f2u(pc.y), p);
            const uint32_t bits = f2u(pc.z);
            unpack_ped_bits(bits, p);
            if (bits & (PB_Y_KERB | PB_Y_LANE)) p.Spy = (bits & PB_Y_KERB) ? sym_kerb(g, p) : sym_lane(g, p);
        }
        {                                                                            // SC:808-809
            const bool kerb = ped_flags(g, p);
            const bool kerb_choice = kerb ? choix_fast<V, MC, NT>(c, g, p, S, t, seen, rng) : true;
            ped_step_stream<V, MC, NT>(c, g, p, S, t, seen, step, rng, kerb_choice);
        }
        const bool left = (p.fl & PF_LEFT) != 0;
        const bool pex = (p.fl & PF_EXIST) != 0;

        // ---- one pass over the (pedestrian, car) pairs.  The reference touches every pair three times: pedestrian.get_data ->
        // delta_l_all (SC:449-460, 508-514), detection (SC:176-264) and new_reward_wait_safety (SC:478-506).  All three are
        // built on the same distance d_raw = |Sc - Sp_x| - Vc^2/2b and the same two lane predicates, so they share one rolled
        // loop here; what forces an order is kept: `behind` of every car is needed before the first pair (n_wait, SC:206),
        // and the accident penalty of the wait reward needs the flag after ALL cars' detection (SC:841-846 finish before
        // SC:849), so the loop leaves the running minimum WITHOUT the penalty per car and a short second loop applies it
        // (x -> x - penalty is monotonic in fp32, so min and subtraction commute bit for bit).
        uint32_t inf_l = 0, cif_l = 0, behind = 0;    // lane predicates (they depend on the car only through its lane); behind: Sc < Sp_x
        {
            const bool neg = p.dir == -1;
            const double y = neg ? -p.Spy : p.Spy;
#pragma unroll 1
            for (int l = 0; l < c.L; ++l) {
                const LanePred lp = lane_pred(g, y, neg, c.L, l, 0.0, 0.0);
                inf_l |= (lp.in_front ? 1u : 0u) << l;
                cif_l |= (lp.crossing ? 1u : 0u) << l;
            }
        }
#pragma unroll 1
        for (int i = 0; i < c.nlead; ++i) behind |= ((S.Sc[i][t] < p.Spx) ? 1u : 0u) << i;
        const double wait_t = (double)p.waitc * c.dt, cross_t = (double)p.crossc * c.dt;
        const double nwait = (double)__builtin_popcount(lead_green & behind);        // SC:206
        const float ts_new = (float)(T::naif ? (((wait_t + 10.0 * cross_t) - time_braking) + 1.0)
                                             : ((((1.0 + nwait) * wait_t + 2.0 * cross_t) - time_braking) + 1.0));
        // The pair body is written branch-free on purpose: lanes (envs) disagree on every one of these
        // conditions, so each pair is evaluated once with selects instead of serialising the sides.
        // The running min/max accumulators (possible_accident, error_scenario, Ts) are fp32: they are stored
        // as fp32 and only ever updated by min/max with a new candidate, and rounding is monotonic, so
        // min(fp32 old, fp32(candidate)) == fp32(min(old, candidate)) bit for bit.
        uint32_t fl = p.fl;
        const bool counts = (p.fl & PF_CROSSING) && (!T::scal || pex);               // SC:844-845
        const bool guard_p = pex && !left && (p.fl & PF_CROSSING);
        double dlmin = T::far;                                                       // delta_l_all, SC:508-514
        float wrun = __builtin_huge_valf();                                          // running min of the wait terms, penalty not yet applied
#pragma unroll 1
        for (int i = 0; i < c.nlead; ++i) {
            const double Sc = S.Sc[i][t], Vc = S.Vc[i][t], rVc = S.rVc[i][t];
            const float light = S.light[i][t];
            const int line = MH_LINE(S, i, t);
            const bool f0 = (inf_l >> line) & 1u, ci = (cif_l >> line) & 1u;
            const bool gi = f0 && ((lead_ok >> i) & 1u);                             // SC:180
            const bool bi = (behind >> i) & 1u;
            const bool ahead = (Sc > p.Spx);
            const double raw = fabs(Sc - p.Spx) - S.brake[i][t];
            const double d = raw - 1.0 * Vc;                                         // delta_l SC:516-520
            const bool dl_ok = (bool)((seen >> i) & 1u) & (Sc <= p.Spx) & f0 & !left & (light >= 0.f);       // one predicate, one select
            dlmin = dl_ok ? dmin(dlmin, d) : dlmin;
            const double wdl = (ahead || left) ? T::far : raw;                       // worst_delta_l SC:522-527
            const bool wneg = wdl < 0.0;
            const bool acc0 = (fl & PF_ACCIDENT) != 0;
            bool wa = (fl & PF_WORST_ACC) != 0;
            const bool ped_accident = T::naif ? (!acc0 && wneg) : (!acc0 && wa);     // SC:181-182 / NA:181-185
            wa = gi ? wneg : wa;
            const bool hit = gi && ped_accident && ci && ((double)S.prevSc[i][t] < p.Spx) && ahead;   // SC:184-185
            fl = (fl & ~PF_WORST_ACC) | (wa ? PF_WORST_ACC : 0u) | (hit ? PF_ACCIDENT : 0u);
            float paf = S.pa[i][t], esf = S.es[i][t], Tsf = S.Ts[i][t];
            {                                                                        // SC:187-201
                const bool slow = Vc < 0.05;
                const double dl64 = slow ? T::far : wdl * rVc;
                const float dl = (float)dl64;
                const bool pos = slow ? (T::far > 0.0) : (wdl > 0.0);
                // -dl-1 (ST:197, NA:200) cancels near dl = -1: that one difference is formed in fp64
                const float lin = T::neg_dl ? (float)(-1.0 * dl64 - 1.0) : (1.0f * dl - 1.0f);
                const float pa = pos ? -exp2_shape(-4.0f * 1.4426950408889634f * dl) : lin;
                paf = (gi && ci) ? fminf(paf, pa) : paf;
            }
            Tsf = (gi && bi) ? fmaxf(ts_new, Tsf) : Tsf;                             // SC:207-208
            {                                                                        // SC:216-237
                const bool red = light < 0.f, grn = light > 0.f;
                const double gap = p.Spx - Sc;
                const float gapf = (float)gap;
                const bool use_exp = red ? (Tsf < 0.f) : (gap > 0.0);
                const float arg = red ? (4.0f * Tsf) : (-4.0f * gapf);
                const float lin = red ? (-1.0f * (1.0f + Tsf)) : (-1.0f - (float)(Sc - p.Spx));
                const float ne = use_exp ? -exp2_shape(1.4426950408889634f * arg) : lin;
                esf = (gi && (red || grn)) ? fminf(ne, esf) : esf;
                if (!T::naif) fl |= (gi && red && ci && bi) ? PF_NOT_WAITING : 0u;
            }
            S.pa[i][t] = paf; S.es[i][t] = esf; S.Ts[i][t] = Tsf;
            {                                                                        // res: SC:250-263 (select, not a branch: lanes disagree)
                float r = paf + esf;
                if (T::danger_sign != 0) {
                    const float extra = ((light < 0.f) && (Tsf > 0.f)) ? 0.5f * green : 0.f;
                    r = (T::danger_sign > 0) ? (r + extra) : (r - extra);
                }
                if (T::scal && !((S.bits[i][t] >> 8) & 1u)) r = 0.f;
                const float rl0 = S.rl[i][t];
                S.rl[i][t] = counts ? rl0 + r : rl0;
            }
            {                                                                        // new_reward_wait_safety SC:478-506, before the penalty
                const float dlw = (float)(d * rVc);
                const float soft = fmaxf(-20.0f * exp2_shape(1.4426950408889634f * (-4.0f * dlw - 4.0f)), -20.0f);
                const float e0 = (Vc < T::wait_thr) ? 0.0f : ((d >= -Vc) ? soft : 20.0f * dlw);
                const bool guard = guard_p && (light > 0.f) && bi && f0;
                wrun = guard ? fminf(wrun, e0) : wrun;
                S.wtmp[i][t] = wrun;
            }
        }
        p.fl = fl;
        if (T::four) {                                                               // the followers are handed to get_data too (C4:796-799)
#pragma unroll 1
            for (int i = c.nlead; i < c.nC; ++i) {
                const double Sc = S.Sc[i][t];
                if (((seen >> i) & 1u) && (Sc <= p.Spx) && ((inf_l >> MH_LINE(S, i, t)) & 1u) && !left && (S.light[i][t] >= 0.f))
                    dlmin = dmin(dlmin, (fabs(Sc - p.Spx) - S.brake[i][t]) - 1.0 * S.Vc[i][t]);
            }
        }
        {                                                                            // pedestrian.get_data SC:449-460 (a select: lanes disagree on pex)
            const double gate = ((p.fl & PF_CROSSING) && !left) ? 1.0 : 0.0;
            const double dnew = dmin(dlmin * gate, p.delta);
            p.delta = pex ? dnew : p.delta;
        }
        // ---- wait-reward contribution (SC:855-857): the reference loops cars outside / pedestrians inside, but worst_dl is
        // per pedestrian and only sees the cars in ascending order, which this loop preserves
        {
            const float acc_pen = (p.fl & PF_ACCIDENT) ? 20.0f : 0.0f;
            const float pw0 = (float)p.wdl;
            float pwdl = pw0;
#pragma unroll 4
            for (int i = 0; i < c.nlead; ++i) {
                pwdl = fminf(pw0, S.wtmp[i][t] - acc_pen);
                const bool grn = pex && (S.light[i][t] > 0.f);
                const float wm = S.wmin[i][t];
                S.wmin[i][t] = grn ? ((!(lead_green & ANY_EXIST) || pwdl < wm) ? pwdl : wm) : wm;
            }
            p.wdl = (double)pwdl;
            if (pex) lead_green |= ANY_EXIST;
        }

        // ---- observation row
        if (op) {
            float *q = op + (int64_t)(ped_o + 9 * j) * (int64_t)ocs;
            q[(int64_t)0 * ocs] = pex ? (float)p.Vpx : 0.f; q[(int64_t)1 * ocs] = pex ? (float)p.Vpy : 0.f;
            q[(int64_t)2 * ocs] = pex ? (float)p.Spx : 0.f; q[(int64_t)3 * ocs] = pex ? (float)p.Spy : 0.f;
            q[(int64_t)4 * ocs] = pex ? (float)p.delta : 0.f; q[(int64_t)5 * ocs] = (pex && left) ? 1.f : 0.f;
            q[(int64_t)6 * ocs] = (pex && (p.fl & PF_IN_CROSS)) ? 1.f : 0.f; q[(int64_t)7 * ocs] = pex ? 1.f : 0.f;
            q[(int64_t)8 * ocs] = pex ? (float)p.dir : 0.f;
        }
        // ---- store pedestrian j
        if (!will_reset) {
            a.ped_a[(int64_t)j * a.N + n] = make_float4((float)p.Spx, (float)p.Spy, (float)p.Vpx, (float)p.Vpy);
            a.ped_b[(int64_t)j * a.N + n] = make_float4((float)p.v0x, (float)p.v0y, (float)p.cstop, (float)p.delta);
            a.ped_c[(int64_t)j * a.N + n] = make_float4((float)p.wdl, u2f(pack_ped_counts(p)),
                                                        u2f(pack_ped_bits(p) | sym_tag(g, p, (float)p.Spy)), 0.f);
        }
    }

    // ---- rewards (SC:849-858), reward_light (SC:846), shaping accumulators back to HBM, done
    {
        float *rp = io.rewards.ptr ? io.rewards.ptr + n * io.rewards.env_stride : nullptr;
        float *lp = io.reward_light.ptr ? io.reward_light.ptr + n * io.reward_light.env_stride : nullptr;
#pragma unroll 1
        for (int i = 0; i < c.nC; ++i) {
            if (i < c.nlead) {
                const double d = S.Vc[i][t] - 10.0;
                // SC:657-665; an fp32 OUTPUT only (never state, never a threshold): the quotient by 100 is taken as a product,
                // one fp64 ulp away at most and far inside the 1e-5 of an fp32 reward (the exact division is ~35 instructions)
                double r = (-10.0 * (d * d)) * 0.01;
                if ((S.light[i][t] > 0.f) && (lead_green & ANY_EXIST)) r += (double)S.wmin[i][t];
                if (rp) rp[(int64_t)i * io.rewards.comp_stride] = (float)r;
                if (lp) lp[(int64_t)i * io.reward_light.comp_stride] = S.rl[i][t];
            }
            if (!will_reset) a.car_b[(int64_t)i * a.N + n] = make_float4(S.pa[i][t], S.es[i][t], S.Ts[i][t], u2f(S.bits[i][t]));
        }
        if (io.done) io.done[n] = done ? 1 : 0;
    }
    // ---- observation: env row (SC:871)
    if (op) {
        float *q = op + (int64_t)(T::car_w * c.nC) * (int64_t)ocs;
        int o = 0;
        q[(int64_t)(o++) * ocs] = (float)(cross * (double)c.L / 2.0);
        q[(int64_t)(o++) * ocs] = (float)ped_traffic;
        if (T::scal) q[(int64_t)(o++) * ocs] = (float)car_traffic;
        q[(int64_t)(o++) * ocs] = (float)c.L;
    }
    step += 1;                                                                       // SC:875
    if (will_reset) {                                                                // in-kernel auto-reset
        reset_and_store<V, MC, MP>(a, c, n, rng, io.obs);
        return;
    }
    a.env_e[n] = make_float4(ee.x, ee.y, u2f(pack_env_word(step, ped_traffic, car_traffic)), u2f(rng.ctr));
}

}  // namespace mhppo
