// env_step.cuh -- the per-thread body of the env-step kernel (one thread = one env).
//
// Replaces Crosswalk_hybrid_multi_*.step (SC:789-878 and siblings) + in-kernel auto-reset.
//
// Shape of the computation (chosen for the SM, not copied from the Python control flow):
//  * cars live in registers (slots fully unrolled); pedestrians are STREAMED: one rolled loop
//    loads pedestrian j from HBM, runs its state machine, its danger detection against every car,
//    its contribution to the wait rewards and its observation row, and stores it back.  The
//    reference runs these as four separate passes over all pedestrians (SC:808-809, 841-846,
//    849-858, 869); the passes only communicate through per-car running min/max and per-pedestrian
//    fields, so interleaving them per pedestrian is exactly equivalent (DESIGN.md "Step kernel").
//    It keeps one pedestrian in registers instead of P, and the hot loop body exists once in SASS
//    (the first version unrolled everything: 28k instructions, instruction-fetch bound).
//  * bulky paths are single out-of-line copies: the sin walking profile (fp64 sincos), the
//    episode reset, and the exact fp64 critical gap.  The gap-acceptance decision
//    `choix_pedestrian` runs every step for a waiting pedestrian, so its log10/pow/normal-draw
//    comparison is a filtered predicate: fp32 with an error bound, fp64 only when undecidable.
//  * geometry predicates is_in_front / is_crossing_in_front are evaluated once per (ped, car)
//    into bit masks and reused by detection, rewards and the observation.
//  * reward-shaping exponentials feed only fp32 outputs: their argument is formed in fp64 and the
//    exponential itself is exp2f (DESIGN.md "Precision"); everything that feeds state or a
//    threshold stays fp64 in the reference's operation order.
#pragma once
#include "../../include/mhppo.h"
#include "env_state.cuh"

namespace mhppo {

struct RngKey { uint32_t k0, k1; int64_t env_id0; };

struct StepIO {
    mhppo_view actions, obs, rewards, reward_light, term_obs;
    uint8_t *done;
    int autoreset;
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
// both sides in fp64 exactly as the reference does.  One Philox block per call, like CG_score.
// Returns (new_ctr << 1) | refused.  Out of line (one copy): a few of these run per waiting pedestrian.
static MH_NOINLINE uint64_t gap_refused_ol(double Spx, double v0y, int dlines, int gender, int age, bool crossing,
                                           double Sc, double Vc, double light, double cross, Rng rng) {
    const double dx = Sc - Spx, vden = Vc + 10e-3;
    bool refused;
    if (!crossing) refused = (fabs(dx / vden) + light) < 0.0;                        // CG = 0., no draw (SC:427-428)
    else {
        const PhiloxBlock b = rng.next();
        const double u1 = u53(b.w0, b.w1), u2 = u53(b.w2, b.w3);
        const double size = fabs((double)dlines) * cross;
        const double pden = fabs(v0y + 10e-3);
        // fp32 estimate: CG = size/|v0y+.01| * 10^(0.09 + gender/age terms + 0.09*z)
        const float adj = 0.09f + ((gender == 1) ? 0.0369f : 0.f) + ((age == 0) ? -0.0355f : ((age == 1) ? -0.0221f : -0.1810f));
#ifdef __CUDA_ARCH__
        const float cz = cospif(2.0f * (float)u2);
#else
        const float cz = cosf(6.2831853f * (float)u2);
#endif
        const float z = sqrtf(-2.0f * logf((float)(1.0 - u1))) * cz;
        const float cg = ((float)size / (float)pden) * exp2f(3.3219280948873623f * (adj + 0.09f * z));
        const float lhs = fabsf((float)dx / (float)vden) + (float)light;
        const float tol = 2e-4f * (fabsf(lhs) + cg) + 1e-30f;
        if (fabsf(lhs - cg) > tol) refused = lhs < cg;
        else refused = (fabs(dx / vden) + light) < cg_exact(v0y, gender, age, size, u1, u2);
    }
    return ((uint64_t)rng.ctr << 1) | (refused ? 1u : 0u);
}
MH_HD bool gap_refused(const PedR &p, const CarR &k, double cross, Rng &rng) {
    const uint64_t r = gap_refused_ol(p.Spx, p.v0y, p.lpos - k.line, p.gender, p.age, (p.fl & PF_CROSSING) != 0, k.Sc, k.Vc,
                                      k.light, cross, rng);
    rng.ctr = (uint32_t)(r >> 1);
    return (r & 1u) != 0;
}

// pedestrian.choix_pedestrian, SC:139-174 / NA:138-177.  `seen` = slots handed to pedestrian.step
// (existing cars in scalable SC:803-806, leaders+followers in 4cars C4:796-799, all otherwise).
template <int V, int MC>
MH_HD bool choix_fast(const Geo &g, const PedR &p, const CarR (&car)[MC], uint32_t seen, Rng &rng) {
    typedef VT<V> T;
    const int nseen = __builtin_popcount(seen);
    // per-car predicates of this decision
    uint32_t inf1 = 0, on_cross = 0, behind = 0, blocked = 0;
#pragma unroll
    for (int i = 0; i < MC; ++i) {
        if (!((seen >> i) & 1u)) continue;
        const bool f1 = in_front(g, p, car[i].line, 1.0);
        const bool over = (car[i].Sc < 4.0 + p.Spx) && (car[i].Sc > p.Spx);          // car body on the crosswalk
        if (f1) inf1 |= 1u << i;
        if (over) on_cross |= 1u << i;
        if (car[i].Sc < p.Spx) behind |= 1u << i;
        if (over && f1 && crossing_in_front(g, p, car[i].line, 0.5)) blocked |= 1u << i;
    }
    if (p.fl & PF_FOLLOW) {
        if (T::naif) {
            // NA:151 shuffles the visiting order for real, but both loops below only ever return False
            // (NA:153-158), so the order cannot change the outcome: only the n-1 draws are consumed
            if (nseen > 1) rng.skip(nseen - 1);
            if (blocked) return false;                                               // NA:153-155
#pragma unroll
            for (int i = 0; i < MC; ++i)                                             // NA:156-158
                if (((seen & behind) >> i) & 1u) { if (car[i].light < 0.0) return false; }
        } else {
            if (T::burn_shuffle && nseen > 1) rng.skip(nseen - 1);                   // SC:152-153: shuffles a temporary
            if (blocked) return false;                                               // SC:154-158
#pragma unroll
            for (int i = 0; i < MC; ++i)                                             // SC:159-161: first upstream car with a light
                if (((seen & behind) >> i) & 1u) { if (car[i].light != 0.0) return car[i].light > 0.0; }
        }
    }
#pragma unroll
    for (int i = 0; i < MC; ++i) {                                                   // SC:162-173
        if (!(((seen & inf1) >> i) & 1u)) continue;
        if ((on_cross >> i) & 1u) return false;
        if ((behind >> i) & 1u) { if (gap_refused(p, car[i], g.cross, rng)) return false; }
    }
    return true;
}

struct WalkOut { double pos, spd; };
// pedestrian.new_pedestrian_sin_y, SC:436-443 with the parameters of SC:94-102
static MH_NOINLINE WalkOut walk_sin(double W, double v0y, double Spy, double dt, int step, int t0c, int dir) {
    const double PI_ = 3.141592653589793, Vm = 2.5;
    const double av = fabs(v0y);
    const double T = W / (av + 10e-3);
    const bool check = ((av * PI_) / 2.0 <= Vm);
    const double A = check ? (PI_ * av / 2.0 + 0.0) : (0.0 + (Vm - av) / (1.0 - (2.0 / PI_)));
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

// ---------------------------------------------------------------------------------------------
// pedestrian.step, SC:297-417, for the pedestrian currently held in registers
template <int V, int MC>
MH_HD void ped_step_stream(const EnvConst &c, const Geo &g, PedR &p, const CarR (&car)[MC], uint32_t seen, int step, Rng &rng) {
    typedef VT<V> T;
    const double dt = c.dt;
    const double pp_y = p.Spy + p.v0y * dt;                                          // SC:298
    const double dy = (double)p.dir * p.Spy;                                         // boolean_ped_position SC:266-275
    p.fl &= ~(PF_LEFT | PF_IN_CROSS);
    if (dy >= g.Hp) p.fl |= PF_LEFT;
    else if (dy > g.Hn) p.fl |= PF_IN_CROSS;
    if (!(p.fl & PF_CROSSING)) return;                                               // SC:308

    auto choix = [&]() -> bool { return choix_fast<V, MC>(g, p, car, seen, rng); };

    bool choose = true;
    const bool first_decision = !(p.fl & PF_DECISION) && (p.fl & PF_AT_CROSSING);    // SC:311
    // kerb arrival is evaluated after the decision block in the reference; it cannot fire in a step
    // that took the decision block (decision is then True), so it is decided up front
    const bool arrive = !first_decision && (dy < g.Hn) && (pp_y * (double)p.dir > g.Hn) && !(p.fl & PF_DECISION);  // SC:320
    const bool in_walk_block = !arrive && (first_decision || (fabs(p.Spy) <= g.Hp) || (p.fl & PF_DECISION));      // SC:331

    // phase A: decide whether this step needs the walking model / a gap-acceptance decision
    bool walk_try = false;
    if (in_walk_block) {
        if (first_decision) {
            choose = choix();
            if (choose) { p.lpos = (p.dir < 0) ? (c.L - 1) : 0; p.fl &= ~PF_AT_CROSSING; }
            p.fl |= PF_DECISION;
            p.t0c = step;
        }
        if (p.tstop != 0) {                                                          // SC:335-339
            p.Vpx = 0.0; p.Vpy = 0.0; p.tstop -= 1; p.t0c += 1;
        } else {
            const double u = rng.random();                                           // SC:346: always drawn
            if ((u < 0.98) && choose) walk_try = true;
            else {                                                                   // SC:405-413
                p.tstop = rng.randint(T::rs_lo, T::rs_hi);
                if (!choose) { p.fl &= ~PF_DECISION; p.tstop = 0; p.waitc += 1; }
                p.Vpx = 0.0; p.Vpy = 0.0; p.t0c += 1;
            }
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
            new_choice = choix();
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
            const double ratio = p.v0x / p.v0y;                                      // SC:73
            p.Spy = ny; p.Vpy = nv;
            p.Spx = p.Spx + p.Vpy * ratio * dt;
            p.Vpx = p.Vpy * ratio;
            p.crossc += 1;
            if (change_line) {                                                       // apply_change_line SC:288-294
                if (fabs(ny) >= g.Hp) p.lpos = (p.dir < 0) ? c.L : ((p.dir > 0) ? -1 : 0);
                else p.lpos = (int)floor((ny + g.Hp) / g.cross);
            }
        } else {                                                                     // SC:389-397
            p.fl |= PF_STOP;
            const double d = fabs(((double)p.dir * (g.W - dtc) - (double)p.dir * g.W / 2.0) - p.Spy);
            const double px = p.Vpx * d / fabs(p.Vpy + 10e-3);
            p.Vpx = px / dt;
            p.Spx = p.Spx + px;
            p.Vpy = (double)p.dir * d / dt;
            p.Spy = (double)p.dir * ((g.W - dtc) - g.W / 2.0);
        }
    } else if (arrive) {                                                             // SC:320-328
        const double px = (p.Vpx * dt) * (fabs(g.Hn - dy) / fabs(p.Vpy * dt + 10e-3));
        p.Vpx = px / dt;
        p.Spx = p.Spx + px;
        p.Vpy = (double)p.dir * fabs(-dy - g.Hp) / dt;
        p.Spy = (double)(-p.dir) * g.W / 2.0;
        p.tstop = 0;
        p.fl |= PF_AT_CROSSING;
    } else if (!in_walk_block) {                                                     // SC:415-417
        p.Spx = p.Spx + p.v0x * dt; p.Vpx = p.v0x;
        p.Spy = p.Spy + p.v0y * dt; p.Vpy = p.v0y;
    }
}

// ---------------------------------------------------------------------------------------------
template <int V, int MC, int MP>
MH_HD void env_step_thread(const EnvArena &a, const EnvConst &c, const RngKey &key, const StepIO &io, int64_t n) {
    typedef VT<V> T;
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

    // ---- cars: load, act (SC:797-802 / C4:791-795 / C42:807-811 / CO:753-755)
    CarR car[MC];
    double prevSc[MC];
    // running min/max shaping accumulators (possible_accident, error_scenario, Ts): kept in fp32.  They
    // are stored as fp32 and only ever updated by min/max with a new candidate, and rounding is
    // monotonic, so min(fp32 old, fp32(candidate)) == fp32(min(old, candidate)) bit for bit.
    float paf[MC], esf[MC], Tsf[MC];
#pragma unroll
    for (int i = 0; i < MC; ++i) {
        if (i >= c.nC) continue;
        const float4 ka = a.car_a[(int64_t)i * a.N + n], kb = a.car_b[(int64_t)i * a.N + n];
        CarR &k = car[i];
        k.Vc = (double)ka.x; k.Sc = (double)ka.y; k.light = (double)ka.z; k.Ac = (double)ka.w;
        paf[i] = kb.x; esf[i] = kb.y; Tsf[i] = kb.z;
        const uint32_t b = f2u(kb.w);
        k.line = (int)(b & 255u); k.exist = (int)((b >> 8) & 1u);
        prevSc[i] = k.Sc;
    }
    {
        const float *ap = io.actions.ptr + n * io.actions.env_stride;
        const int half = c.nA / 2;
        const int64_t cs = io.actions.comp_stride;
        if (T::scal) {
#pragma unroll
            for (int i = 0; i < MC; ++i) {
                if (i >= c.nC) continue;
                double acc = (double)ap[(int64_t)i * cs];
                const double lt = (double)ap[(int64_t)(half + i) * cs];
                if ((i & 1) && car[i > 0 ? i - 1 : 0].exist && car[i].exist)
                    acc = dmin(idm(c, car[i], car[i > 0 ? i - 1 : 0].Sc, car[i > 0 ? i - 1 : 0].Vc), acc);
                else acc = dmin(2.0, acc);
                car_move<V>(c, car[i], acc, lt);
            }
        } else if (T::four) {
            const int nl = c.nb_car;
#pragma unroll
            for (int i = 0; i < MC / 2; ++i) {
                if (i >= nl) continue;
                car_move<V>(c, car[i], (double)ap[(int64_t)i * cs], (double)ap[(int64_t)(half + i) * cs]);
            }
#pragma unroll
            for (int i = 0; i < MC / 2; ++i) {
                if (i >= nl) continue;
                CarR &f = car[MC / 2 + i];                  // follower of leader i (MC == 2*nb_car for these classes)
                const double a_idm = idm(c, f, car[i].Sc, car[i].Vc);
                if (V == V_4CARS2) car_move<V>(c, f, dmin(a_idm, (double)ap[(int64_t)(nl + i) * cs]), (double)ap[(int64_t)(half + nl + i) * cs]);
                else car_move<V>(c, f, a_idm, car[i].light);
            }
        } else {
#pragma unroll
            for (int i = 0; i < MC; ++i) {
                if (i >= c.nC) continue;
                car_move<V>(c, car[i], (double)ap[(int64_t)i * cs], (double)ap[(int64_t)(half + i) * cs]);
            }
        }
    }
    // ---- per-car quantities reused by every pedestrian
    double brake[MC], rVc[MC];
    uint32_t seen = 0, lead_ok = 0;     // seen: cars handed to pedestrian.step; lead_ok: cars detection may touch
    double green = 0.0;                 // SC:246
#pragma unroll
    for (int i = 0; i < MC; ++i) {
        if (i >= c.nC) continue;
        brake[i] = car[i].Vc * car[i].Vc / (-2.0 * c.acc_lo);
        rVc[i] = 1.0 / car[i].Vc;
        if (!T::scal || car[i].exist) seen |= 1u << i;
        if (i < c.nlead && (!T::scal || car[i].exist)) {
            lead_ok |= 1u << i;
            if (car[i].light > 0.0) green += 1.0;
        }
    }
    const bool done = (step >= c.done_idx) || (ped_traffic <= 0);                    // SC:874
    const bool will_reset = done && io.autoreset;
    const mhppo_view ov = will_reset ? io.term_obs : io.obs;
    float *const op = ov.ptr ? ov.ptr + n * ov.env_stride : nullptr;
    const int64_t ocs = ov.comp_stride;
    const int env_w = T::scal ? 4 : 3;
    const int ped_o = T::car_w * c.nC + env_w;
    const double time_braking = -(10.0 / (2.0 * c.acc_lo)) + 1.0;                    // SC:580

    float rl[MC], wmin[MC];
#pragma unroll
    for (int i = 0; i < MC; ++i) { rl[i] = 0.f; wmin[i] = 0.f; }
    bool any_exist = false;

    // ---- pedestrians, streamed
#pragma unroll 1
    for (int j = 0; j < c.nP; ++j) {
        PedR p;
        {
            const float4 pa = a.ped_a[(int64_t)j * a.N + n], pb = a.ped_b[(int64_t)j * a.N + n], pc = a.ped_c[(int64_t)j * a.N + n];
            p.Spx = (double)pa.x; p.Spy = (double)pa.y; p.Vpx = (double)pa.z; p.Vpy = (double)pa.w;
            p.v0x = (double)pb.x; p.v0y = (double)pb.y; p.cstop = (double)pb.z; p.delta = (double)pb.w;
            p.wdl = (double)pc.x;
            unpack_ped_counts(f2u(pc.y), p);
            const uint32_t bits = f2u(pc.z);
            unpack_ped_bits(bits, p);
            if (bits & (PB_Y_KERB | PB_Y_LANE)) p.Spy = (bits & PB_Y_KERB) ? sym_kerb(g, p) : sym_lane(g, p);
        }
        ped_step_stream<V, MC>(c, g, p, car, seen, step, rng);                       // SC:808-809

        // geometry predicates of this pedestrian against every car lane, and the shared gaps
        uint32_t inf = 0, cif = 0, behind = 0;    // behind: Sc < Sp_x
        double raw[MC];                           // |Sc - Sp_x| - Vc^2/(2b)   (SC:527)
#pragma unroll
        for (int i = 0; i < MC; ++i) {
            if (i >= c.nC) continue;
            if (in_front(g, p, car[i].line, 0.0)) inf |= 1u << i;
            if (crossing_in_front(g, p, car[i].line, 0.0)) cif |= 1u << i;
            if (car[i].Sc < p.Spx) behind |= 1u << i;
            raw[i] = fabs(car[i].Sc - p.Spx) - brake[i];
        }
        const bool left = (p.fl & PF_LEFT) != 0;
        const double wait_t = (double)p.waitc * c.dt, cross_t = (double)p.crossc * c.dt;

        // ---- detection (SC:176-264); placeholders run it too (SC:842-843)
        double nwait = 0.0;                                                          // SC:206
#pragma unroll
        for (int i = 0; i < MC; ++i)
            if (((lead_ok & behind) >> i) & 1u) { if (car[i].light > 0.0) nwait += 1.0; }
        const float ts_new = (float)(T::naif ? (((wait_t + 10.0 * cross_t) - time_braking) + 1.0)
                                             : ((((1.0 + nwait) * wait_t + 2.0 * cross_t) - time_braking) + 1.0));
        // Written branch-free on purpose: lanes (envs) disagree on every one of these conditions, so
        // each pair is evaluated once with selects instead of serialising the sides of the branches.
        uint32_t fl = p.fl;
#pragma unroll
        for (int i = 0; i < MC; ++i) {
            if (i >= c.nlead) continue;
            CarR &k = car[i];
            const bool gi = ((lead_ok & inf) >> i) & 1u;                             // SC:180
            const bool bi = (behind >> i) & 1u, ci = (cif >> i) & 1u;
            const bool ahead = (k.Sc > p.Spx);
            const double wdl = (ahead || left) ? T::far : raw[i];                    // worst_delta_l SC:522-527
            const bool wneg = wdl < 0.0;
            const bool acc0 = (fl & PF_ACCIDENT) != 0;
            bool wa = (fl & PF_WORST_ACC) != 0;
            const bool ped_accident = T::naif ? (!acc0 && wneg) : (!acc0 && wa);     // SC:181-182 / NA:181-185
            wa = gi ? wneg : wa;
            const bool hit = gi && ped_accident && ci && (prevSc[i] < p.Spx) && ahead;   // SC:184-185
            fl = (fl & ~PF_WORST_ACC) | (wa ? PF_WORST_ACC : 0u) | (hit ? PF_ACCIDENT : 0u);
            {                                                                        // SC:187-201
                const bool slow = k.Vc < 0.05;
                const double dl64 = slow ? T::far : wdl * rVc[i];
                const float dl = (float)dl64;
                const bool pos = slow ? (T::far > 0.0) : (wdl > 0.0);
                // -dl-1 (ST:197, NA:200) cancels near dl = -1: that one difference is formed in fp64
                const float lin = T::neg_dl ? (float)(-1.0 * dl64 - 1.0) : (1.0f * dl - 1.0f);
                const float pa = pos ? -exp2f(-4.0f * 1.4426950408889634f * dl) : lin;
                paf[i] = (gi && ci) ? fminf(paf[i], pa) : paf[i];
            }
            Tsf[i] = (gi && bi) ? fmaxf(ts_new, Tsf[i]) : Tsf[i];                    // SC:207-208
            {                                                                        // SC:216-237
                const bool red = k.light < 0.0, grn = k.light > 0.0;
                const double gap = p.Spx - k.Sc;
                const float gapf = (float)gap;
                const bool use_exp = red ? (Tsf[i] < 0.f) : (gap > 0.0);
                const float arg = red ? (4.0f * Tsf[i]) : (-4.0f * gapf);
                const float lin = red ? (-1.0f * (1.0f + Tsf[i])) : (-1.0f - (float)(k.Sc - p.Spx));
                const float ne = use_exp ? -exp2f(1.4426950408889634f * arg) : lin;
                esf[i] = (gi && (red || grn)) ? fminf(ne, esf[i]) : esf[i];
                if (!T::naif) fl |= (gi && red && ci && bi) ? PF_NOT_WAITING : 0u;
            }
        }
        p.fl = fl;
        if ((p.fl & PF_CROSSING) && (!T::scal || (p.fl & PF_EXIST))) {               // SC:844-845, res: SC:250-263
#pragma unroll
            for (int i = 0; i < MC; ++i) {
                if (i >= c.nlead) continue;
                float r = paf[i] + esf[i];
                if (T::danger_sign != 0) {
                    const float extra = ((car[i].light < 0.0) && (Tsf[i] > 0.f)) ? 0.5f * (float)green : 0.f;
                    r = (T::danger_sign > 0) ? (r + extra) : (r - extra);
                }
                if (T::scal && !car[i].exist) r = 0.f;
                rl[i] += r;
            }
        }

        // ---- wait-reward contribution (SC:855-857 with new_reward_wait_safety SC:478-506); the
        // reference loops cars outside / pedestrians inside, but worst_dl is per pedestrian and only
        // sees the cars in ascending order, which this loop preserves
        {
            const bool ex = (p.fl & PF_EXIST) != 0;
            const bool guard_p = ex && !left && (p.fl & PF_CROSSING);
            const float acc_pen = (p.fl & PF_ACCIDENT) ? 20.0f : 0.0f;
            float pwdl = (float)p.wdl;
#pragma unroll
            for (int i = 0; i < MC; ++i) {
                if (i >= c.nlead) continue;
                const CarR &k = car[i];
                const bool grn = ex && (k.light > 0.0);
                const bool guard = grn && guard_p && (((behind & inf) >> i) & 1u);
                const double d = raw[i] - 1.0 * k.Vc;                                // delta_l SC:516-520
                const float dl = (float)(d * rVc[i]);
                const float soft = fmaxf(-20.0f * exp2f(1.4426950408889634f * (-4.0f * dl - 4.0f)), -20.0f);
                float e = (k.Vc < T::wait_thr) ? 0.0f : ((d >= -k.Vc) ? soft : 20.0f * dl);
                e = e - acc_pen;
                pwdl = (guard && e < pwdl) ? e : pwdl;
                wmin[i] = grn ? ((!any_exist || pwdl < wmin[i]) ? pwdl : wmin[i]) : wmin[i];
            }
            p.wdl = (double)pwdl;
            any_exist = any_exist || ex;
        }

        // ---- observation row (pedestrian.get_data SC:449-460, delta_l_all SC:508-514)
        if (p.fl & PF_EXIST) {
            double dl = T::far;
#pragma unroll
            for (int i = 0; i < MC; ++i) {
                if (!((seen >> i) & 1u)) continue;                                   // SC:803-806 existing cars / C4:796-799 all
                if ((car[i].Sc <= p.Spx) && ((inf >> i) & 1u) && !left && (car[i].light >= 0.0))
                    dl = dmin(dl, raw[i] - 1.0 * car[i].Vc);
            }
            const double gate = ((p.fl & PF_CROSSING) && !left) ? 1.0 : 0.0;
            p.delta = dmin(dl * gate, p.delta);
        }
        if (op) {
            float *q = op + (int64_t)(ped_o + 9 * j) * ocs;
            const bool ex = (p.fl & PF_EXIST) != 0;
            q[0 * ocs] = ex ? (float)p.Vpx : 0.f; q[1 * ocs] = ex ? (float)p.Vpy : 0.f;
            q[2 * ocs] = ex ? (float)p.Spx : 0.f; q[3 * ocs] = ex ? (float)p.Spy : 0.f;
            q[4 * ocs] = ex ? (float)p.delta : 0.f; q[5 * ocs] = (ex && left) ? 1.f : 0.f;
            q[6 * ocs] = (ex && (p.fl & PF_IN_CROSS)) ? 1.f : 0.f; q[7 * ocs] = ex ? 1.f : 0.f;
            q[8 * ocs] = ex ? (float)p.dir : 0.f;
        }
        // ---- store pedestrian j
        if (!will_reset) {
            a.ped_a[(int64_t)j * a.N + n] = make_float4((float)p.Spx, (float)p.Spy, (float)p.Vpx, (float)p.Vpy);
            a.ped_b[(int64_t)j * a.N + n] = make_float4((float)p.v0x, (float)p.v0y, (float)p.cstop, (float)p.delta);
            a.ped_c[(int64_t)j * a.N + n] = make_float4((float)p.wdl, u2f(pack_ped_counts(p)),
                                                        u2f(pack_ped_bits(p) | sym_tag(g, p, (float)p.Spy)), 0.f);
        }
    }

    // ---- rewards (SC:849-858), reward_light (SC:846), done
    {
        float *rp = io.rewards.ptr ? io.rewards.ptr + n * io.rewards.env_stride : nullptr;
        float *lp = io.reward_light.ptr ? io.reward_light.ptr + n * io.reward_light.env_stride : nullptr;
#pragma unroll
        for (int i = 0; i < MC; ++i) {
            if (i >= c.nlead) continue;
            const double d = car[i].Vc - 10.0;
            double r = (-10.0 * (d * d)) / 100.0;                                    // SC:657-665
            if ((car[i].light > 0.0) && any_exist) r += (double)wmin[i];
            if (rp) rp[(int64_t)i * io.rewards.comp_stride] = (float)r;
            if (lp) lp[(int64_t)i * io.reward_light.comp_stride] = rl[i];
        }
        if (io.done) io.done[n] = done ? 1 : 0;
    }
    // ---- observation: car rows (car.get_data SC:652-655) and env row (SC:871)
    if (op) {
#pragma unroll
        for (int i = 0; i < MC; ++i) {
            if (i >= c.nC) continue;
            const CarR &k = car[i];
            float *q = op + (int64_t)(T::car_w * i) * ocs;
            const bool ex = !T::scal || k.exist;
            q[0 * ocs] = ex ? (float)k.Ac : 0.f; q[1 * ocs] = ex ? (float)k.Vc : 0.f;
            q[2 * ocs] = ex ? (float)(10.0 - k.Vc) : 10.f; q[3 * ocs] = ex ? (float)k.Sc : -1000.f;
            q[4 * ocs] = ex ? (float)k.light : 0.f; q[5 * ocs] = (float)k.line;
            if (T::scal) q[6 * ocs] = ex ? 1.f : 0.f;
        }
        float *q = op + (int64_t)(T::car_w * c.nC) * ocs;
        int o = 0;
        q[(o++) * ocs] = (float)(cross * (double)c.L / 2.0);
        q[(o++) * ocs] = (float)ped_traffic;
        if (T::scal) q[(o++) * ocs] = (float)car_traffic;
        q[(o++) * ocs] = (float)c.L;
    }
    step += 1;                                                                       // SC:875
    if (will_reset) {                                                                // in-kernel auto-reset
        reset_and_store<V, MC, MP>(a, c, n, rng, io.obs);
        return;
    }
    // ---- store cars + env word
#pragma unroll
    for (int i = 0; i < MC; ++i) {
        if (i >= c.nC) continue;
        const CarR &k = car[i];
        a.car_a[(int64_t)i * a.N + n] = make_float4((float)k.Vc, (float)k.Sc, (float)k.light, (float)k.Ac);
        a.car_b[(int64_t)i * a.N + n] = make_float4(paf[i], esf[i], Tsf[i],
                                                    u2f(((uint32_t)k.line & 255u) | (((uint32_t)k.exist & 1u) << 8)));
    }
    a.env_e[n] = make_float4(ee.x, ee.y, u2f(pack_env_word(step, ped_traffic, car_traffic)), u2f(rng.ctr));
}

}  // namespace mhppo
