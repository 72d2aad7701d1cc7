// env_core.cuh -- one crosswalk env advanced by one thread.
//
// This is the device-side semantics of `Crosswalk_hybrid_multi_{stop,naif,coop,coop_4cars,
// coop_4cars2,coop_scalable}.step/reset` (reference: Environments/Env_hybrid_multi_*.py; citation
// abbreviations SC/CO/ST/NA/C4/C42 as in include/mhppo.h).  One templated body covers the six
// classes; the behavioural deltas between the files are compile-time switches (`VT<V>`).
//
// Design (B200-first, not a translation of the Python object graph):
//  * thread = env, lane = env inside a warp; the reference's filtered Python lists become bit masks
//    over car slots (order preserved); the step itself (env_step.cuh) keeps the cars of a CTA in
//    shared memory and streams the pedestrians through one rolled loop;
//  * persisted state is fp32 / packed integers in HBM (env_state.cuh), arithmetic is fp64 in the
//    reference's operation order (compile with -fmad=false) so threshold flags are bit-exact and
//    values differ from the fp64 reference only by the final fp32 rounding;
//  * time, t0, waiting_time and crossing_time are integer multiples of dt;
//  * the data-dependent RNG consumption of the reference is reproduced with a per-env Philox
//    cursor (philox.cuh); throw-away draws (SC:153) just advance the counter;
//  * auto-reset runs inside the step kernel through a non-inlined reset on a scratch copy.
//
// This header holds the shared pieces: behaviour table, register structs, geometry predicates,
// car dynamics, the full observation writer and the episode reset.
//
// The file is also compilable as plain C++ (tests/hostsim) so the kernel logic can be unit-tested
// against the golden fixtures on a machine without a GPU; the product never runs that build.
#pragma once
#include <math.h>
#include <stdint.h>

#include "philox.cuh"

#ifdef __CUDACC__
#define MH_NOINLINE __host__ __device__ __noinline__
#else
#define MH_NOINLINE __attribute__((noinline))
#endif

namespace mhppo {

enum { V_STOP = 0, V_NAIF = 1, V_COOP = 2, V_4CARS = 3, V_4CARS2 = 4, V_SCAL = 5 };

// pedestrian flag bits: positions are those of the canonical state dump (include/mhppo.h)
enum : uint32_t {
    PF_EXIST = 1u << 0, PF_CROSSING = 1u << 1, PF_DECISION = 1u << 2, PF_AT_CROSSING = 1u << 3,
    PF_LEFT = 1u << 4, PF_IN_CROSS = 1u << 5, PF_NOT_WAITING = 1u << 6, PF_ACCIDENT = 1u << 7,
    PF_WORST_ACC = 1u << 8, PF_FOLLOW = 1u << 9, PF_STOP = 1u << 10, PF_NEED_STOP = 1u << 11,
    PF_RATIO_EPS = 1u << 12     // pedestrian.reset_ped replaced the constructor's ratio v0x / v0y (SC:73) by v0x / (v0y + 1e-3) (SC:111)
};

// compile-time behaviour table of the six env files ("Behavioural deltas")
template <int V>
struct VT {
    static constexpr bool scal = (V == V_SCAL);
    static constexpr bool four = (V == V_4CARS || V == V_4CARS2);
    static constexpr bool naif = (V == V_NAIF);
    static constexpr bool stop = (V == V_STOP);
    static constexpr bool burn_shuffle = (V == V_STOP || V == V_COOP || V == V_SCAL);  // SC:153 ST:152 CO:152
    static constexpr bool has_need_to_stop = (V == V_STOP || V == V_4CARS2 || V == V_SCAL);  // SC:371 ST:366 C42:448
    static constexpr bool draws_need_to_stop = has_need_to_stop;                     // SC:91
    static constexpr bool draws_cross_stop = (V == V_STOP || V == V_COOP || V == V_4CARS2 || V == V_SCAL);  // SC:92 CO:91
    static constexpr bool placeholder_keeps_x = (V == V_COOP || V == V_SCAL);        // CO:68 vs ST:68
    static constexpr int nts_lo = stop ? 2 : 5, nts_hi = stop ? 15 : 35;             // ST:367, SC:372
    static constexpr int rs_lo = (V == V_4CARS2) ? 5 : 2;                            // C42:480
    static constexpr int rs_hi = stop ? 15 : ((V == V_4CARS2) ? 35 : 5);             // ST:402, SC:406
    static constexpr double far = scal ? 100.0 : 0.0;                                // SC:191,509,525 vs CO:190,493,509
    static constexpr bool neg_dl = (V == V_STOP || V == V_NAIF);                     // ST:197 NA:200: -dl-1
    static constexpr int danger_sign = (V == V_STOP || V == V_COOP || V == V_4CARS2) ? +1
                                       : ((V == V_4CARS || V == V_SCAL) ? -1 : 0);   // SC:258 CO:256 C4:345 NA:252
    static constexpr double wait_thr = stop ? 0.01 : 0.05;                           // ST:478 / SC:482
    static constexpr double Ts0 = scal ? 0.0 : -10.0;                                // SC:555 / CO:539
    static constexpr int car_w = scal ? 7 : 6;                                       // SC:655 / CO:612
};

struct CarR {
    double Vc, Sc, light, Ac, pa, es, Ts;
    int line, exist;
};
struct PedR {
    double Spx, Spy, Vpx, Vpy, v0x, v0y, cstop, delta, wdl;
    int t0c, waitc, crossc, tstop, lpos, dir, gender, age;
    uint32_t fl;
};
template <int MC, int MP>
struct EnvR {
    double cross;
    int step, ped_traffic, car_traffic;
    Rng rng;
    CarR car[MC];
    PedR ped[MP];
};

// runtime configuration shared by every env of a handle (kernel argument, by value)
struct EnvConst {
    int nC, nP, L, nb_car, nlead, nA, nobs, done_idx, sin_model;
    double dt, acc_lo, acc_hi;       // car_b[0,0], car_b[1,0]
    double pb[8];                    // ped_b row-major
    double cross_lo, cross_hi;       // cross_b
    // derived on the host by env_const_finish (same IEEE double arithmetic as on the device): the three
    // per-handle constants the reference recomputes on every call
    double brake_den;                // -2 * acc_lo                      (SC:527, 518: Vc^2 / (2b))
    double idm_den;                  // 2 * sqrt(-acc_lo * acc_hi)       (SC:611)
    double time_braking;             // -(10 / (2 * acc_lo)) + 1         (SC:580)
    double brake_inv;                // 1 / brake_den if brake_den is a power of two (then x * brake_inv == x / brake_den exactly), else 0
    double idm_rden, rdt;            // RN(1 / idm_den), RN(1 / dt): correctly rounded reciprocals for div_rcp
};
inline void env_const_finish(EnvConst &c) {
    c.brake_den = -2.0 * c.acc_lo;
    c.idm_den = 2.0 * sqrt(-c.acc_lo * c.acc_hi);
    c.time_braking = -(10.0 / (2.0 * c.acc_lo)) + 1.0;
    int ex = 0;
    c.brake_inv = (c.brake_den > 0.0 && frexp(c.brake_den, &ex) == 0.5) ? 1.0 / c.brake_den : 0.0;
    c.idm_rden = 1.0 / c.idm_den;
    c.rdt = 1.0 / c.dt;
}

struct Geo {
    double cross, W, Hn, Hp, Lf;     // Hn = -W/2, Hp = W/2
};
MH_HD Geo make_geo(double cross, int L) {
    Geo g; g.cross = cross; g.Lf = (double)L; g.W = g.Lf * cross; g.Hn = (-g.W) / 2.0; g.Hp = g.W / 2.0;
    return g;
}

MH_HD double dmin(double a, double b) { return b < a ? b : a; }   // Python min(a, b)
// x / den for den > 0, bit-identical to the plain quotient.  A zero numerator (a stopped car: Vc = 0, a = 0) sends
// the device's fp64 division into its out-of-line slow path; a third of the cars stand still at any time, so the
// quotient is formed from a stand-in numerator and the (signed) zero is passed through instead.
// keeps the compiler from folding a select of the operand back through the division (it turned
// `(z ? 1 : x) / den` into `z ? 1 / den : x / den`, and the zero was back in the divider)
MH_HD double opaque(double x) {
#if defined(__CUDA_ARCH__) && !defined(MH_NO_DIVFIX)
    asm volatile("" : "+d"(x));
#endif
    return x;
}
MH_HD double div_pos(double x, double den) {
    const bool z = (x == 0.0);
    const double q = opaque(z ? 1.0 : x) / den;
    return z ? x : q;
}
MH_HD double dmax(double a, double b) { return b > a ? b : a; }   // Python max(a, b)
// x / d from the correctly rounded reciprocal rd = RN(1 / d) of a divisor that is a constant of the kernel (dt, 10, the IDM
// denominator, 1 - 2/pi) or of the env (the crosswalk width): q0 = RN(x * rd) is within an ulp or two of the quotient, the
// remainder r = x - d * q0 is exact in one FMA, and RN(q0 + r * rd) is the correctly rounded quotient (Markstein's theorem:
// the last step of every software division, here without the steps that first have to FIND the reciprocal).  Three
// instructions instead of ~35 with a slow-path branch; checked against `/` on 1.3e9 operands incl. near-exact quotients
// (tools/div_rcp_check.c), and by the host-build parity tests, which run this code against the oracle's plain divisions.
// Preconditions (hold for every call site): d > 0 finite; x = 0 or 2^-900 < |x| < 2^900 (kinematic values built from fp32
// state by a few operations).  A zero numerator keeps its sign (the FMA correction would turn -0 into +0).
MH_HD double div_rcp(double x, double d, double rd) {
    const double q0 = x * rd;
#ifdef __CUDA_ARCH__
    const double q = __fma_rn(__fma_rn(-d, q0, x), rd, q0);
#else
    const double q = fma(fma(-d, q0, x), rd, q0);
#endif
    return (x == 0.0) ? x : q;
}

// pedestrian.is_in_front, SC:463-468
MH_HD bool in_front(const Geo &g, const PedR &p, int line, double nl) {
    if (p.dir == -1) return p.Spy >= (g.Hp - g.cross * ((g.Lf - 0.5 * nl) - (double)line)) - 0.001;
    return p.Spy <= (g.Hn + g.cross * (((double)line - 0.5 * nl) + 1.0)) + 0.001;
}
// pedestrian.is_crossing_in_front, SC:470-476
MH_HD bool crossing_in_front(const Geo &g, const PedR &p, int line, double pl) {
    if (p.dir == -1) return p.Spy < g.Hp - g.cross * (((g.Lf - (double)line) - 1.0) - pl);
    return p.Spy > g.Hn + g.cross * ((double)line - pl);
}

// Both predicates for one lane without the branch on the walking direction (lanes of a warp disagree on it, so the two
// sides above run one after the other at half occupancy).  With y = -Sp_y for dir == -1 and y = Sp_y otherwise:
//   in front:          y <= (Hn + cross * a) + 0.001,   a = (dir == -1 ? L - line : line + 1) - 0.5 * nl
//   crossing in front: y >   Hn + cross * a',           a' = (dir == -1 ? L - line - 1 : line) - pl
// Bit-identical to the two functions above: Hn = -Hp exactly, negation commutes with rounding ((Hp - b) - 0.001 =
// -((Hn + b) + 0.001)), and a, a' are small integers or half-integers, exact in either operation order.
struct LanePred { bool in_front, crossing; };
MH_HD LanePred lane_pred(const Geo &g, double y, bool neg, int L, int line, double nl, double pl) {
    const int ai = neg ? (L - line) : (line + 1);
    LanePred r;
    r.in_front = y <= (g.Hn + g.cross * ((double)ai - 0.5 * nl)) + 0.001;
    r.crossing = y > g.Hn + g.cross * ((double)(ai - 1) - pl);
    return r;
}

// pedestrian.CG_score, SC:419-428 -- one normal draw, only for crossing pedestrians
MH_HD double cg_score(const PedR &p, double size, Rng &rng) {
    if (!(p.fl & PF_CROSSING)) return 0.0;
    double lv = 0.09 + log10(size / fabs(p.v0y + 10e-3));
    lv = lv + 0.0369 * (double)(p.gender == 1);
    lv = lv + -0.0355 * (double)(p.age == 0);
    lv = lv + -0.0221 * (double)(p.age == 1);
    lv = lv + -0.1810 * (double)(p.age == 2);
    lv = lv + rng.normal(0.0, 0.09);
    return pow(10.0, lv);
}

// car.sigma SC:592-602
MH_HD double sigma_lim(double Vc, double a, double dt) {
    if (Vc == 0.0) return (a > 0.0) ? 1.0 : 0.0;       // max(0, a/|a|); a == 0 -> 0
    if (a > 0.0) return 1.0;
    return dmax(dmin(-Vc / (dt * a), 1.0), 0.0);
}
// car.follow_action SC:604-624 (IDM; car_follower.follow_action C4:109-129)
MH_HD double idm(const EnvConst &c, const CarR &k, double leadSc, double leadVc) {
    const double dd = leadSc - k.Sc;
    const double dv = k.Vc - leadVc;
    const double s = (2.0 + (k.Vc * 2.0)) + div_rcp(k.Vc * dv, c.idm_den, c.idm_rden);
    const double r4 = div_rcp(k.Vc, 10.0, 0.1), r2 = s / dd;               // (the literal 0.1 is RN(1 / 10))
    return c.acc_hi * ((1.0 - (r4 * r4) * (r4 * r4)) - r2 * r2);
}
// car.step SC:627-650 (ST:599-618 adds the clamp; car_follower.transform C4:79-93)
template <int V>
MH_HD void car_move(const EnvConst &c, CarR &k, double action, double light) {
    double a = dmin(dmax(action, c.acc_lo), c.acc_hi);                               // acceleration() SC:589-590
    const double sg = sigma_lim(k.Vc, a, c.dt);
    if (VT<V>::stop && sg > 0.0) a = dmax(a, -k.Vc / (c.dt * sg));                   // ST:604-605
    const double fa = a * sg;                                                        // discount [1,0,0] is the identity
    const double v = k.Vc + c.dt * fa;
    const double s = ((fa * (c.dt * c.dt)) / 2.0 + (k.Vc * c.dt)) + k.Sc;
    k.Ac = fa; k.Vc = v; k.Sc = s; k.light = light;
}

// ---------------------------------------------------------------------------------------------
// observation writer: car.get_data SC:652-655, env row SC:871, pedestrian.get_data SC:449-460 with
// delta_l_all SC:508-514.  Out(k, value) stores component k of this env's flat observation
// (key order car,(car_follow),env,ped).  Mutates the running-min `delta` exactly like get_data.
template <int V, int MC, int MP, class Out>
MH_HD void write_obs(const EnvConst &c, EnvR<MC, MP> &e, bool at_reset, Out &out) {
    typedef VT<V> T;
    const Geo g = make_geo(e.cross, c.L);
    int o = 0;
#pragma unroll
    for (int i = 0; i < MC; ++i) {
        if (i >= c.nC) continue;
        const CarR &k = e.car[i];
        if (T::scal && !k.exist) {
            out(o + 0, 0.f); out(o + 1, 0.f); out(o + 2, 10.f); out(o + 3, -1000.f); out(o + 4, 0.f);
            out(o + 5, (float)k.line); out(o + 6, 0.f);
        } else {
            out(o + 0, (float)k.Ac); out(o + 1, (float)k.Vc); out(o + 2, (float)(10.0 - k.Vc)); out(o + 3, (float)k.Sc);
            out(o + 4, (float)k.light); out(o + 5, (float)k.line);
            if (T::scal) out(o + 6, 1.f);
        }
        o += T::car_w;
    }
    out(o++, (float)(e.cross * (double)c.L / 2.0));
    out(o++, (float)e.ped_traffic);
    if (T::scal) out(o++, (float)e.car_traffic);
    out(o++, (float)c.L);
    // cars handed to get_data: every slot of self.cars at reset (SC:922-929; leaders only C4:886-893);
    // existing cars in step (SC:803-806), leaders + followers in 4cars (C4:796-799)
    const int upto = at_reset ? c.nlead : c.nC;
#pragma unroll
    for (int j = 0; j < MP; ++j) {
        if (j >= c.nP) continue;
        PedR &p = e.ped[j];
        if (!(p.fl & PF_EXIST)) {
#pragma unroll
            for (int q = 0; q < 9; ++q) out(o + q, 0.f);
        } else {
            double dl = T::far;
            const bool left = (p.fl & PF_LEFT) != 0;
#pragma unroll
            for (int i = 0; i < MC; ++i) {
                if (i >= upto) continue;
                const CarR &k = e.car[i];
                if (T::scal && !at_reset && !k.exist) continue;
                if ((k.Sc <= p.Spx) && in_front(g, p, k.line, 0.0) && !left && (k.light >= 0.0)) {
                    const double nd = (fabs(k.Sc - p.Spx) - (k.Vc * k.Vc / c.brake_den)) - 1.0 * k.Vc;
                    dl = dmin(dl, nd);
                }
            }
            const double gate = ((p.fl & PF_CROSSING) && !left) ? 1.0 : 0.0;
            p.delta = dmin(dl * gate, p.delta);
            out(o + 0, (float)p.Vpx); out(o + 1, (float)p.Vpy); out(o + 2, (float)p.Spx); out(o + 3, (float)p.Spy);
            out(o + 4, (float)p.delta); out(o + 5, left ? 1.f : 0.f); out(o + 6, (p.fl & PF_IN_CROSS) ? 1.f : 0.f);
            out(o + 7, 1.f); out(o + 8, (float)p.dir);
        }
        o += 9;
    }
}

// ---------------------------------------------------------------------------------------------
// reset: SC:884-946 (CO:838-892, C4:850-911, C42:866-927); pedestrian.__init__ SC:15-104 and
// car.__init__ SC:531-581 consume the stream in the reference's order.
template <int V>
MH_HD void ped_init(const EnvConst &c, double W, double cross, PedR &p, bool real, Rng &rng) {
    typedef VT<V> T;
    p.fl = real ? (PF_EXIST | PF_CROSSING) : 0u;
    (void)rng.randint(0, 20);                                                        // time_to_remove SC:43 (dead)
    const int fr = rng.randint(0, 9);
    if (T::naif ? (fr < 10) : (fr < 3)) p.fl |= PF_FOLLOW;                           // SC:48 / NA:46
    p.dir = 2 * rng.randint(0, 1) - 1;                                               // SC:54
    p.lpos = (p.dir < 0) ? c.L : -1;                                                 // SC:55
    p.v0x = rng.uniform(c.pb[0], c.pb[4]);                                           // SC:56
    p.v0y = rng.uniform(c.pb[1], c.pb[5]) * (double)p.dir;                           // SC:57
    p.Spx = rng.uniform(c.pb[2], c.pb[6]);                                           // SC:61
    p.Spy = (rng.uniform(c.pb[3], c.pb[7]) - W / 2.0) * (double)p.dir;               // SC:62
    if (!real) {                                                                     // SC:66-68
        p.v0x = 0.0; p.v0y = 0.0;
        p.Spy = c.pb[3] * (double)p.dir;
        if (!T::placeholder_keeps_x) p.Spx = c.pb[2];
    }
    p.Vpx = p.v0x; p.Vpy = p.v0y;
    p.gender = rng.randint(0, 1);                                                    // SC:84
    p.age = rng.randint(0, 2);                                                       // SC:85
    if (real) rng.skip(1);                                                           // SC:86: CG normal draw, value dead
    p.delta = 0.0; p.wdl = 0.0; p.cstop = 0.0;
    p.t0c = 0; p.waitc = 0; p.crossc = 0; p.tstop = 0;
    if (T::draws_need_to_stop) { if (rng.uniform(0.0, 1.0) < 0.5) p.fl |= PF_NEED_STOP; }   // SC:91
    else if (V == V_COOP) p.fl |= PF_NEED_STOP;                                      // CO:90
    if (T::draws_cross_stop) p.cstop = rng.uniform(-W / 2.0 + 0.2, W / 2.0 - 0.2);   // SC:92
    (void)cross;
}

template <int V>
MH_HD void car_init(const EnvConst &c, double W, CarR &k, int arg, int exist, Rng &rng) {
    typedef VT<V> T;
    const double v0 = 10.0;
    const double mean_speed_ped = c.pb[1] + c.pb[5] / 2.0;                           // SC:565
    const double fct = (W * v0) / mean_speed_ped;                                    // SC:570
    const double lo = (c.pb[3] * v0) / c.pb[1], hi = (c.pb[7] * v0) / c.pb[5];       // SC:573-574
    const double u = rng.uniform(lo - fct, hi);                                      // SC:576
    k.line = T::scal ? arg / 2 : arg;                                                // SC:537
    k.Sc = T::scal ? (u - 20.0 * (double)(arg % 2)) : u;
    k.Vc = v0; k.Ac = 0.0; k.light = 0.0; k.pa = 0.0; k.es = 0.0; k.Ts = T::Ts0; k.exist = exist;
}

template <int V, int MC, int MP>
MH_NOINLINE void reset_env(const EnvConst &c, EnvR<MC, MP> &e) {
    typedef VT<V> T;
    Rng &rng = e.rng;
    e.cross = rng.uniform(c.cross_lo, c.cross_hi);                                   // SC:890 (kept fp64 in HBM)
    const double W = (double)c.L * e.cross;
    for (int j = 0; j < c.nP; ++j) ped_init<V>(c, W, e.cross, e.ped[j], false, rng); // SC:897-899
    if (T::scal) {
        for (int i = 0; i < c.nC; ++i) car_init<V>(c, W, e.car[i], i / 2, 0, rng);   // SC:900-901
        e.car_traffic = rng.randint(1, c.nb_car);                                    // SC:902
        uint64_t pool = 0xFEDCBA9876543210ull;                                       // random.sample SC:903
        const int n = c.nC;
        for (int i = 0; i < e.car_traffic; ++i) {
            int j = i + (int)floor(rng.random() * (double)(n - i));
            if (j > n - 1) j = n - 1;
            const uint64_t a = (pool >> (4 * i)) & 15u, b = (pool >> (4 * j)) & 15u;
            pool &= ~((15ull << (4 * i)) | (15ull << (4 * j)));
            pool |= (b << (4 * i)) | (a << (4 * j));
        }
        for (int q = 0; q < e.car_traffic; ++q) {                                    // SC:904-906
            const int s = (int)((pool >> (4 * q)) & 15u);
            car_init<V>(c, W, e.car[s], s / 2, 1, rng);
        }
    } else {
        for (int i = 0; i < c.nb_car; ++i) car_init<V>(c, W, e.car[i], i % c.L, 1, rng);   // CO:852-853
        e.car_traffic = c.nb_car;
        if (T::four)
            for (int i = 0; i < c.nb_car; ++i) car_init<V>(c, W, e.car[c.nb_car + i], e.car[i].line, 1, rng);  // C4:868
    }
    e.ped_traffic = rng.randint(1, c.nP);                                            // SC:912
    for (int j = 0; j < e.ped_traffic; ++j) ped_init<V>(c, W, e.cross, e.ped[j], true, rng);   // SC:913-915
    if (T::four)
        for (int i = 0; i < c.nb_car; ++i) {                                         // C4:878-879 / C42:894-895
            CarR &f = e.car[c.nb_car + i];
            const double gap = (V == V_4CARS2) ? rng.uniform(10.0, 30.0) : 15.0;
            f.Vc = 10.0; f.Sc = e.car[i].Sc - gap; f.light = 0.0; f.line = e.car[i].line;
        }
    e.step = 0;
}

// ---------------------------------------------------------------------------------------------
// State injection: Crosswalk_*.reset_pedestrian (SC:948-955; CO:895, ST:904, C4:913, C42:929) rebuilds EVERY
// pedestrian as a placeholder (the constructors' draws are consumed) and then runs pedestrian.reset_ped (SC:106-137) on slot
// num_ped with q = (speed_x, speed_y, pos_x, pos_y, dl, leave, CZ, exist, direction); `leave` / `CZ` are kept as truth values.
// naif (NA:897-898 -> NA:100-135): no rebuild, q = (speed_x, speed_y, pos_x, pos_y, dl, direction, cross); the new cross becomes
// the env's (its only caller, reset_distrib NA:903-913, sets env.cross to the same value and resets every pedestrian).
template <int V, int MC, int MP>
MH_HD void inject_pedestrian(const EnvConst &c, EnvR<MC, MP> &e, int num_ped, const float *q) {
    typedef VT<V> T;
    PedR &p = e.ped[num_ped];
    if (T::naif) {
        e.cross = (double)q[6];                                                      // NA:101-102
        const double W = (double)c.L * e.cross;
        p.dir = (int)q[5];                                                           // NA:103
        p.v0x = (double)q[0]; p.v0y = (double)q[1] * (double)p.dir;                  // NA:104
        p.Spx = (double)q[2]; p.Spy = ((double)q[3] - W / 2.0) * (double)p.dir;      // NA:106
        p.delta = (double)q[4];
        p.fl = (p.fl | PF_EXIST) & ~(PF_LEFT | PF_IN_CROSS);                         // NA:116-120
    } else {
        const double W = (double)c.L * e.cross;
        for (int j = 0; j < c.nP; ++j) ped_init<V>(c, W, e.cross, e.ped[j], false, e.rng);   // SC:949-951
        p.v0x = (double)q[0]; p.v0y = (double)q[1]; p.Spx = (double)q[2]; p.Spy = (double)q[3];   // SC:107-110
        p.dir = (int)q[8];                                                           // SC:115
        p.delta = (double)q[4];
        p.fl &= ~(PF_EXIST | PF_LEFT | PF_IN_CROSS);
        if (q[7] != 0.f) p.fl |= PF_EXIST;
        if (q[5] != 0.f) p.fl |= PF_LEFT;
        if (q[6] != 0.f) p.fl |= PF_IN_CROSS;                                        // SC:117-123
    }
    p.Vpx = p.v0x; p.Vpy = p.v0y;
    p.fl |= PF_RATIO_EPS | PF_CROSSING;                                              // SC:111, 124
    p.lpos = (p.dir < 0) ? c.L : ((p.dir > 0) ? -1 : 0);                             // SC:116
    e.rng.skip(1);                                                                   // SC:125: CG_score draws, the value is dead
    p.t0c = 0;                                                                       // SC:127
}
// reset_cars -> car.reset_car (SC:957-958, 583-587): q = (speed_x, pos_x, light, line)
MH_HD void inject_car(CarR &k, const float *q) {
    k.Sc = (double)q[1]; k.Vc = (double)q[0]; k.light = (double)q[2]; k.line = (int)q[3];
}

}  // namespace mhppo
