/* mhppo_oracle.c -- CPU restatement of the MH-PPO crosswalk envs.  TEST INFRASTRUCTURE.
 *
 * See mhppo_oracle.h for the role of this file (checker + CPU baseline; never used by the
 * product) and for the citation abbreviations (SC/CO/ST/NA/C4/C42).  Every function cites the
 * reference lines it restates.  Arithmetic is IEEE double in the reference's operation order
 * (compile with -ffp-contract=off); `store_f32` rounds the persisted fields to fp32 after every
 * reset/step, which is the only difference between "reference semantics" (store_f32=0, used to
 * pin this file against the Python reference) and "HBM semantics" (store_f32=1, used to check
 * the CUDA kernels, whose state lives in fp32).
 *
 * Deliberate deviations from the reference, all outside its defined behaviour:
 *  - RNG: CPython's Mersenne Twister is replaced by the project's Philox contract
 *    (oracle/refshim/philox.py); the reference is driven with the same contract when compared.
 *  - `sigma(a)` with Vc == 0 and a == 0 raises ZeroDivisionError in the reference (SC:594);
 *    here sg = 0 (what numpy-float operands give: max(0., nan) == 0.).
 *  - `CG_score(0)` raises ValueError (math.log10(0), SC:423); here log10(0) = -inf -> CG = 0.
 *  - time, t0, waiting_time, crossing_time are kept as integer multiples of dt (the reference
 *    accumulates dt in fp64; the difference is a few ulp and never feeds an equality test).
 */
#include "mhppo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define PI 3.141592653589793 /* math.pi */

enum { F_EXIST = 0, F_IS_CROSSING, F_DECISION, F_AT_CROSSING, F_PED_LEFT, F_PED_IN_CROSS,
       F_NOT_WAITING, F_ACCIDENT, F_WORST_ACC, F_FOLLOW_RULE, F_STOP, F_NEED_TO_STOP, F_RATIO_EPS };

typedef struct {
    double Ac, Vc, Sc, light, pa, es, Ts;
    int line, exist;
} OCar;

typedef struct {
    double Vpx, Vpy, Spx, Spy, v0x, v0y, cross_stop, delta, worst_dl;
    int t0c, waitc, crossc, time_stop, line_pos, dir, gender, age;
    int exist, is_crossing, decision, at_crossing, ped_left, ped_in_cross, not_waiting, accident,
        worst_acc, follow_rule, stop, need_to_stop;
    int ratio_eps;          /* set by reset_ped: ratio = v0x / (v0y + 1e-3) (SC:111) instead of the constructor's v0x / v0y (SC:73) */
} OPed;

typedef struct {
    double cross;
    int step_idx, ped_traffic, car_traffic;
    uint32_t ctr;
    uint64_t env_id;
    OCar car[MHO_MAXC];
    OPed ped[MHO_MAXP];
} OEnv;

typedef struct {
    mho_cfg c;
    int C, nlead, A, nobs, done_idx;
    int64_t N;
    OEnv *env;
} OHandle;

/* ------------------------------------------------------------------ RNG contract */
/* oracle/refshim/philox.py is the normative statement of the stream. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void mho_philox(uint32_t ctr, uint64_t env_id, uint64_t seed, uint32_t out[4]) {
    out[0] = ctr; out[1] = 0; out[2] = (uint32_t)env_id; out[3] = (uint32_t)(env_id >> 32);
    philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}

typedef struct { const OHandle *h; OEnv *e; } Ctx;

static inline double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}
static double rng_random(Ctx *x) {
    uint32_t w[4];
    mho_philox(x->e->ctr++, x->e->env_id, x->h->c.seed, w);
    return u53(w[0], w[1]);
}
static double rng_uniform(Ctx *x, double a, double b) { return a + (b - a) * rng_random(x); }
static int rng_randint(Ctx *x, int a, int b) {
    int n = b - a + 1;
    int k = (int)floor(rng_random(x) * n);
    if (k > n - 1) k = n - 1;
    return a + k;
}
static double rng_normal(Ctx *x, double mu, double sigma) {
    uint32_t w[4];
    mho_philox(x->e->ctr++, x->e->env_id, x->h->c.seed, w);
    double u1 = u53(w[0], w[1]), u2 = u53(w[2], w[3]);
    return mu + sigma * (sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * PI * u2));
}
static void rng_shuffle(Ctx *x, int *v, int n) {
    for (int i = n - 1; i > 0; --i) {
        int j = (int)floor(rng_random(x) * (i + 1));
        if (j > i) j = i;
        int t = v[i]; v[i] = v[j]; v[j] = t;
    }
}

/* ------------------------------------------------------------------ small helpers */
static inline double f32r(double v) { return (double)(float)v; }
static inline double pymin(double a, double b) { return b < a ? b : a; } /* Python min(a,b) */
static inline double pymax(double a, double b) { return b > a ? b : a; } /* Python max(a,b) */

/* CPython float.__floordiv__ (Objects/floatobject.c float_floor_div) */
static double py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0 && ((wx < 0) != (mod < 0))) div -= 1.0;
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return copysign(0.0, vx / wx);
}

int mho_n_slots(const mho_cfg *c) {
    if (c->variant == MHO_SCALABLE) return 2 * c->nb_lines;                 /* SC:900-901 */
    if (c->variant == MHO_4CARS || c->variant == MHO_4CARS2) return 2 * c->nb_car; /* C4:865-869 */
    return c->nb_car;                                                       /* CO:852-853 */
}
int mho_n_lead(const mho_cfg *c) {
    if (c->variant == MHO_4CARS || c->variant == MHO_4CARS2) return c->nb_car;
    return mho_n_slots(c);
}
int mho_n_action(const mho_cfg *c) {
    if (c->variant == MHO_SCALABLE) return 4 * c->nb_lines;                 /* SC:798-802 */
    if (c->variant == MHO_4CARS2) return 4 * c->nb_car;                     /* C42:809-811 */
    return 2 * c->nb_car;                                                   /* CO:754-755 */
}
int mho_n_obs(const mho_cfg *c) {
    int carw = c->variant == MHO_SCALABLE ? 7 : 6;                          /* SC:655, CO:612 */
    int envw = c->variant == MHO_SCALABLE ? 4 : 3;                          /* SC:871, CO:825 */
    return carw * mho_n_slots(c) + envw + 9 * c->nb_ped;                    /* followers: car_follow, C4:838 */
}

/* ------------------------------------------------------------------ geometry predicates */
/* pedestrian.is_in_front, SC:463-468 */
static int is_in_front(const Ctx *x, const OPed *p, int car_line, double next_line) {
    double cross = x->e->cross, L = x->h->c.nb_lines, W = L * cross;
    double line_1 = (-W / 2) + cross * (car_line - 0.5 * next_line + 1);
    double line_2 = (W / 2) - cross * (L - 0.5 * next_line - car_line);
    if (p->dir == -1) return p->Spy >= line_2 - 0.001;
    return p->Spy <= line_1 + 0.001;
}
/* pedestrian.is_crossing_in_front, SC:470-476 */
static int is_crossing_in_front(const Ctx *x, const OPed *p, int car_line, double prev_line) {
    double cross = x->e->cross, L = x->h->c.nb_lines, W = L * cross;
    double line_1 = (-W / 2) + cross * (car_line - prev_line);
    double line_2 = (W / 2) - cross * (L - car_line - 1 - prev_line);
    if (p->dir == -1) return p->Spy < line_2;
    return p->Spy > line_1;
}

/* pedestrian.CG_score, SC:419-428 (one normalvariate draw when is_crossing) */
static double cg_score(Ctx *x, const OPed *p, double crossing_size) {
    if (!p->is_crossing) return 0.0;
    double gamma = log10(crossing_size / fabs(p->v0y + 10e-3));
    double lv = 0.09 + gamma;
    lv = lv + 0.0369 * (p->gender == 1);
    lv = lv + -0.0355 * (p->age == 0);
    lv = lv + -0.0221 * (p->age == 1);
    lv = lv + -0.1810 * (p->age == 2);
    lv = lv + rng_normal(x, 0.0, 0.09);
    return pow(10.0, lv);
}

/* pedestrian.choix_pedestrian, SC:139-174 (NA:138-177 for the naif rule) */
static int choix(Ctx *x, OPed *p, int n, const double *Vc, const double *Sc, const int *line,
                 const double *light) {
    const int v = x->h->c.variant;
    const double car_size = 4;
    if (p->follow_rule) {
        int order[MHO_MAXC];
        for (int i = 0; i < n; ++i) order[i] = i;
        if (n > 1) {
            if (v == MHO_NAIF) {
                rng_shuffle(x, order, n);            /* NA:151 really permutes the visiting order */
            } else if (v == MHO_STOP || v == MHO_COOP || v == MHO_SCALABLE) {
                int tmp[MHO_MAXC];                   /* SC:153 shuffles a temporary: burns n-1 draws */
                for (int i = 0; i < n; ++i) tmp[i] = i;
                rng_shuffle(x, tmp, n);
            }                                        /* C4:259-260, C42: commented out */
        }
        for (int k = 0; k < n; ++k) {                /* SC:154-158 */
            int i = order[k];
            if (is_crossing_in_front(x, p, line[i], 0.5) * is_in_front(x, p, line[i], 1.0))
                if ((Sc[i] < car_size + p->Spx) * (Sc[i] > p->Spx)) return 0;
        }
        for (int k = 0; k < n; ++k) {                /* SC:159-161 / NA:157-158 */
            int i = order[k];
            if (v == MHO_NAIF) {
                if (Sc[i] < p->Spx && light[i] < 0) return 0;
            } else {
                if (Sc[i] < p->Spx && light[i] != 0) return light[i] > 0.;
            }
        }
    }
    for (int i = 0; i < n; ++i) {                    /* SC:162-173 */
        if (is_in_front(x, p, line[i], 1.0)) {
            if ((Sc[i] < car_size + p->Spx) * (Sc[i] > p->Spx)) return 0;
            if (Sc[i] < p->Spx) {
                double car_time = fabs((Sc[i] - p->Spx) / (Vc[i] + 10e-3));
                double CG = cg_score(x, p, fabs((double)(p->line_pos - line[i])) * x->e->cross);
                if (car_time + light[i] < CG) return 0;
            }
        }
    }
    return 1;
}

/* sin-profile parameters, SC:94-102 (recomputed from the persisted v0y / cross) */
static void sin_params(const Ctx *x, const OPed *p, double *A, double *B, double *w) {
    double W = x->h->c.nb_lines * x->e->cross;
    double abs_speed = fabs(p->v0y), Vm = 2.5;
    double T = W / (abs_speed + 10e-3);
    int check = ((abs_speed * PI) / 2.0 <= Vm);
    *A = check * PI * abs_speed / 2.0 + (!check) * (Vm - abs_speed) / (1.0 - (2.0 / PI));
    *B = (!check) * (Vm - *A);
    *w = PI / T;
}

/* pedestrian.function_step: new_pedestrian_sin_y SC:436-443 / new_pedestrian_unif_y SC:433-434 */
static void function_step(const Ctx *x, const OPed *p, double time, double *pos, double *spd) {
    const double dt = x->h->c.dt;
    if (x->h->c.sin_model && p->is_crossing) {
        double A, B, w, W = x->h->c.nb_lines * x->e->cross;
        sin_params(x, p, &A, &B, &w);
        double t = time + dt, t0 = p->t0c * dt;
        double speed_p = (A * sin(w * (t - t0)) + B);
        double pos_p = ((-W / 2.) + (A * (-cos(w * (t - t0)) + cos(w * 0.0)) / w));
        if (pos_p >= 0.0 && speed_p < fabs(p->v0y)) {
            *pos = p->Spy + p->v0y * dt; *spd = p->v0y;
            return;
        }
        *pos = p->dir * pos_p; *spd = p->dir * speed_p;
        return;
    }
    *pos = p->Spy + p->v0y * dt; *spd = p->v0y;
}

/* pedestrian.step, SC:297-417 */
static void ped_step(Ctx *x, OPed *p, double time, int n, const double *Vc, const double *Sc,
                     const int *line, const double *light) {
    const mho_cfg *c = &x->h->c;
    const double dt = c->dt, cross = x->e->cross, W = c->nb_lines * cross;
    const int L = c->nb_lines, v = c->variant;
    double pp_y = p->Spy + p->v0y * dt;                                   /* SC:298 */
    /* boolean_ped_position, SC:266-275 */
    if (p->dir * p->Spy >= W / 2) { p->ped_in_cross = 0; p->ped_left = 1; }
    else if (p->dir * p->Spy > -W / 2) { p->ped_in_cross = 1; p->ped_left = 0; }
    else { p->ped_in_cross = 0; p->ped_left = 0; }
    if (!p->is_crossing) return;                                          /* SC:308 */
    int choose = 1;
    if ((!p->decision) * (p->at_crossing)) {                              /* SC:311-318 */
        choose = choix(x, p, n, Vc, Sc, line, light);
        if (choose) { p->line_pos = (L - 1) * (p->dir < 0); p->at_crossing = 0; }
        p->decision = 1;
        p->t0c = x->e->step_idx;                                          /* t0 = time */
    }
    if ((p->Spy * p->dir < -W / 2.) * (pp_y * p->dir > -W / 2.) * (!p->decision)) { /* SC:320-328 */
        double pos_p_x = (p->Vpx * dt) * (fabs(-W / 2. - p->Spy * p->dir) / fabs(p->Vpy * dt + 10e-3));
        p->Vpx = pos_p_x / dt;
        p->Spx = p->Spx + pos_p_x;
        p->Vpy = p->dir * fabs(-p->Spy * p->dir - (W / 2.)) / dt;
        p->Spy = -p->dir * W / 2.;
        p->time_stop = 0;
        p->at_crossing = 1;
    } else if ((fabs(p->Spy) <= W / 2) + p->decision) {                   /* SC:331 */
        if (p->time_stop != 0) {                                          /* SC:335-339 */
            p->Vpx = 0.0; p->Vpy = 0.0;
            p->time_stop -= 1;
            p->t0c += 1;
        } else if ((rng_uniform(x, 0, 1) < 0.98) * choose) {              /* SC:346 */
            p->decision = 0;
            double new_spy, new_vpy;
            function_step(x, p, time, &new_spy, &new_vpy);
            int change_line = 0;                                          /* will_change_line SC:281-286 */
            if (fabs(new_spy) < W / 2) {
                double new_line = py_floordiv(new_spy + W / 2, cross);
                if (new_line != p->line_pos)
                    if (fabs(p->Spy) < W / 2) change_line = 1;
            }
            double dtc = (L - p->line_pos - 1) * cross * (p->dir > 0);    /* SC:350-351 */
            dtc += (p->line_pos) * cross * (p->dir < 0);
            int new_choice;
            if (change_line && (dtc > 0. && dtc < W)) {                   /* SC:353-361 */
                new_choice = choix(x, p, n, Vc, Sc, line, light);
                if (new_choice && p->stop) p->stop = 0;
            } else new_choice = 0;
            int has_nts = (v == MHO_STOP || v == MHO_4CARS2 || v == MHO_SCALABLE);
            if (p->stop) {                                                /* SC:363-369 */
                p->Vpx = 0.0; p->Vpy = 0.0;
                p->t0c += 1;
                if (change_line) p->waitc += 1;
            } else if (has_nts && p->need_to_stop && p->Spy < p->cross_stop && pp_y > p->cross_stop) {
                /* SC:371-380, ST:366-375 (2,15), C42:448-457 */
                p->time_stop = (v == MHO_STOP) ? rng_randint(x, 2, 15) : rng_randint(x, 5, 35);
                p->need_to_stop = 0;
                p->Vpx = 0.0; p->Vpy = 0.0;
                p->t0c += 1;
            } else if ((!change_line) || (change_line && new_choice)) {   /* SC:382-387 */
                double ratio = p->ratio_eps ? p->v0x / (p->v0y + 1e-3) : p->v0x / p->v0y;   /* SC:73 / SC:111 */
                p->Spy = new_spy; p->Vpy = new_vpy;
                p->Spx = p->Spx + p->Vpy * ratio * dt;
                p->Vpx = p->Vpy * ratio;
                p->crossc += 1;
                if (change_line && new_choice) {                          /* apply_change_line SC:288-294 */
                    if (fabs(new_spy) >= W / 2) p->line_pos = L * (p->dir < 0) - 1 * (p->dir > 0);
                    else {
                        double new_line = py_floordiv(new_spy + W / 2, cross);
                        if (new_line != p->line_pos) p->line_pos = (int)new_line;
                    }
                }
            } else {                                                      /* SC:389-397 */
                p->stop = 1;
                double distance = fabs(p->dir * (W - dtc) - p->dir * W / 2. - p->Spy);
                double pos_p_x = (p->Vpx) * (distance) / fabs(p->Vpy + 10e-3);
                p->Vpx = pos_p_x / dt;
                p->Spx = p->Spx + pos_p_x;
                p->Vpy = p->dir * distance / dt;
                p->Spy = p->dir * ((W - dtc) - W / 2.);
            }
        } else {                                                          /* SC:405-413 */
            int lo = 2, hi = 5;                                           /* CO:390, NA:384, C4:464, SC:406 */
            if (v == MHO_STOP) { lo = 2; hi = 15; }                       /* ST:402 */
            if (v == MHO_4CARS2) { lo = 5; hi = 35; }                     /* C42:480 */
            p->time_stop = rng_randint(x, lo, hi);
            if (!choose) { p->decision = 0; p->time_stop = 0; p->waitc += 1; }
            p->Vpx = 0.0; p->Vpy = 0.0;
            p->t0c += 1;
        }
    } else {                                                              /* SC:415-417 */
        p->Spx = p->Spx + p->v0x * dt; p->Vpx = p->v0x;
        p->Spy = p->Spy + p->v0y * dt; p->Vpy = p->v0y;
    }
}

/* pedestrian.worst_delta_l, SC:522-527 (fallback 100 in scalable, 0 elsewhere: CO:509) */
static double worst_delta_l(const Ctx *x, const OPed *p, double car_pos, double car_speed, int car_line) {
    const double b = -2.0 * x->h->c.car_b[0];
    if (car_pos > p->Spx || p->ped_left || (!is_in_front(x, p, car_line, 0)))
        return x->h->c.variant == MHO_SCALABLE ? 100.0 : 0.0;
    return fabs(car_pos - p->Spx) - (car_speed * car_speed / b);
}

/* pedestrian.detection, SC:176-264 (CO:176-259, ST, NA:178-254, C4:284-347, C42) */
static void detection(Ctx *x, OPed *p, int ncars, const double *prev_Sc, double *res) {
    const int v = x->h->c.variant;
    OCar *cars = x->e->car;
    const double time_braking = -(10.0 / (2.0 * x->h->c.car_b[0])) + 1.0;  /* SC:580, Vc = 10 at init */
    for (int i = 0; i < ncars; ++i) {
        OCar *ci = &cars[i];
        int guard = is_in_front(x, p, ci->line, 0);
        if (v == MHO_SCALABLE) guard = guard && ci->exist;                 /* SC:180 */
        if (!guard) continue;
        int ped_accident;
        if (v == MHO_NAIF) {                                               /* NA:181-185: uses the NEW flag */
            p->worst_acc = worst_delta_l(x, p, ci->Sc, ci->Vc, ci->line) < 0 ? 1 : 0;
            ped_accident = (!p->accident) * (p->worst_acc);
        } else {                                                           /* SC:181-182 */
            ped_accident = (!p->accident) * (p->worst_acc);
            p->worst_acc = worst_delta_l(x, p, ci->Sc, ci->Vc, ci->line) < 0 ? 1 : 0;
        }
        if (ped_accident * (is_crossing_in_front(x, p, ci->line, 0) * (prev_Sc[i] < p->Spx) * (ci->Sc > p->Spx)))
            p->accident = 1;                                               /* SC:184-185 */
        if (is_crossing_in_front(x, p, ci->line, 0)) {                     /* SC:187-201 */
            double dl;
            if (ci->Vc < 0.05) dl = (v == MHO_SCALABLE) ? 100. : 0.;       /* SC:191 / CO:190 */
            else dl = worst_delta_l(x, p, ci->Sc, ci->Vc, ci->line) / ci->Vc;
            double possible_accident;
            if (dl > 0) possible_accident = -1. * exp(-4. * dl);
            else if (v == MHO_STOP || v == MHO_NAIF) possible_accident = -1. * dl - 1; /* ST:197, NA:200 */
            else possible_accident = 1. * dl - 1;                          /* SC:198, CO:197 */
            ci->pa = pymin(ci->pa, possible_accident);
        }
        if (v == MHO_NAIF) {                                               /* NA:207-208 */
            if (ci->Sc < p->Spx)
                ci->Ts = pymax(p->waitc * x->h->c.dt + 10. * (p->crossc * x->h->c.dt) - time_braking + 1., ci->Ts);
        } else {                                                           /* SC:206-208 */
            double cars_light_waiting = 0;
            for (int k = 0; k < ncars; ++k)
                if (cars[k].light > 0. && cars[k].Sc < p->Spx && (v != MHO_SCALABLE || cars[k].exist))
                    cars_light_waiting += 1.0;
            if (ci->Sc < p->Spx)
                ci->Ts = pymax((1. + cars_light_waiting) * (p->waitc * x->h->c.dt) +
                                   2. * (p->crossc * x->h->c.dt) - time_braking + 1., ci->Ts);
        }
        if (ci->light < 0.0) {                                             /* SC:216-228 */
            double new_error;
            if (ci->Ts < 0) new_error = -1. * exp(4. * ci->Ts);
            else new_error = -1. * (1 + ci->Ts);
            if (v != MHO_NAIF)
                if (!p->not_waiting && is_crossing_in_front(x, p, ci->line, 0) && (ci->Sc < p->Spx))
                    p->not_waiting = 1;
            ci->es = pymin(new_error, ci->es);
        }
        if (ci->light > 0.0) {                                             /* SC:230-237 */
            double new_error;
            if (p->Spx - ci->Sc > 0) new_error = -1. * exp(-4. * (p->Spx - ci->Sc));
            else new_error = -1. * (1 + ci->Sc - p->Spx);
            ci->es = pymin(new_error, ci->es);
        }
    }
    double cars_light_green = 0;                                           /* SC:246 */
    for (int k = 0; k < ncars; ++k)
        if (cars[k].light > 0. && (v != MHO_SCALABLE || cars[k].exist)) cars_light_green += 1.0;
    for (int i = 0; i < ncars; ++i) {                                      /* SC:250-263 */
        double r = cars[i].pa + cars[i].es;
        double extra = 0.5 * cars_light_green * (cars[i].light < 0.) * (cars[i].Ts > 0);
        if (v == MHO_STOP || v == MHO_COOP || v == MHO_4CARS2) r = r + extra;  /* ST:256, CO:256, C42:350 */
        else if (v == MHO_4CARS || v == MHO_SCALABLE) r = r - extra;           /* C4:345, SC:258 */
        /* NA:252: no extra term */
        if (v == MHO_SCALABLE && !cars[i].exist) r = 0.;
        res[i] = r;
    }
}

/* pedestrian.new_reward_wait_safety, SC:478-506 (speed threshold 0.01 in ST:478) */
static double reward_wait_safety(const Ctx *x, OPed *p, double car_speed, double car_pos, int car_line) {
    const double b = -2.0 * x->h->c.car_b[0];
    if ((!p->ped_left) * (p->is_crossing) * (car_pos < p->Spx) * is_in_front(x, p, car_line, 0)) {
        double exp_dl;
        double thr = x->h->c.variant == MHO_STOP ? 0.01 : 0.05;
        if (car_speed < thr) exp_dl = 0.;
        else {
            /* delta_l SC:516-520: the guard above makes its 0.0 fallback unreachable */
            double d = fabs(car_pos - p->Spx) - (car_speed * car_speed / b) - 1.0 * (car_speed);
            double dl = d / (car_speed);
            if (dl >= -1.) exp_dl = pymax(-20. * exp(-4. * (dl) - 4.), -20.0);
            else exp_dl = 20. * dl;
        }
        exp_dl = exp_dl - (p->accident) * 20;
        if (exp_dl < p->worst_dl) p->worst_dl = exp_dl;
    }
    return p->worst_dl;
}

/* pedestrian.get_data SC:449-460 + delta_l_all SC:508-514; writes 9 floats */
static void ped_get_data(const Ctx *x, OPed *p, int n, const double *Sc, const double *Vc,
                         const int *line, const double *light, float *out) {
    if (!p->exist) { for (int k = 0; k < 9; ++k) out[k] = 0.f; return; }
    const double b = -2.0 * x->h->c.car_b[0];
    double delta_l = x->h->c.variant == MHO_SCALABLE ? 100.0 : 0.0;         /* SC:509 / CO:493 */
    for (int i = 0; i < n; ++i)
        if ((Sc[i] <= p->Spx) && is_in_front(x, p, line[i], 0) && (!p->ped_left) && (light[i] >= 0)) {
            double nd = fabs(Sc[i] - p->Spx) - (Vc[i] * Vc[i] / b) - 1.0 * (Vc[i]);
            delta_l = pymin(delta_l, nd);
        }
    p->delta = pymin(delta_l * (p->is_crossing) * (!p->ped_left), p->delta);
    out[0] = (float)p->Vpx; out[1] = (float)p->Vpy; out[2] = (float)p->Spx; out[3] = (float)p->Spy;
    out[4] = (float)p->delta; out[5] = (float)p->ped_left; out[6] = (float)p->ped_in_cross;
    out[7] = (float)p->exist; out[8] = (float)p->dir;
}

/* car.get_data SC:652-655 / CO:609-612 */
static int car_get_data(const Ctx *x, const OCar *c, float *out) {
    if (x->h->c.variant == MHO_SCALABLE) {
        if (!c->exist) {
            out[0] = 0.f; out[1] = 0.f; out[2] = 10.f; out[3] = -1000.f; out[4] = 0.f;
            out[5] = (float)c->line; out[6] = 0.f;
        } else {
            out[0] = (float)c->Ac; out[1] = (float)c->Vc; out[2] = (float)(10 - c->Vc); out[3] = (float)c->Sc;
            out[4] = (float)c->light; out[5] = (float)c->line; out[6] = 1.f;
        }
        return 7;
    }
    out[0] = (float)c->Ac; out[1] = (float)c->Vc; out[2] = (float)(10 - c->Vc); out[3] = (float)c->Sc;
    out[4] = (float)c->light; out[5] = (float)c->line;
    return 6;
}

/* observation dict flattened in sorted key order: car, (car_follow), env, ped (SC:868-871, C4:834-838) */
static void write_obs(Ctx *x, float *obs, int at_reset) {
    const OHandle *h = x->h; OEnv *e = x->e;
    const int v = h->c.variant, P = h->c.nb_ped;
    float *o = obs;
    for (int i = 0; i < h->C; ++i) o += car_get_data(x, &e->car[i], o);  /* leaders then followers */
    *o++ = (float)(e->cross * h->c.nb_lines / 2.);
    *o++ = (float)e->ped_traffic;
    if (v == MHO_SCALABLE) *o++ = (float)e->car_traffic;
    *o++ = (float)h->c.nb_lines;
    /* car list handed to ped.get_data: at reset every slot of self.cars (SC:922-929; leaders only in
     * C4:886-893); in step existing cars only in scalable (SC:803-806), leaders+followers in 4cars */
    double Sc[MHO_MAXC], Vc[MHO_MAXC], light[MHO_MAXC]; int line[MHO_MAXC]; int n = 0;
    int upto = (at_reset ? h->nlead : h->C);
    for (int i = 0; i < upto; ++i) {
        if (!at_reset && v == MHO_SCALABLE && !e->car[i].exist) continue;
        Sc[n] = e->car[i].Sc; Vc[n] = e->car[i].Vc; light[n] = e->car[i].light; line[n] = e->car[i].line; ++n;
    }
    for (int j = 0; j < P; ++j) { ped_get_data(x, &e->ped[j], n, Sc, Vc, line, light, o); o += 9; }
}

/* ------------------------------------------------------------------ cars */
/* car.sigma SC:592-602 */
static double sigma(const Ctx *x, double Vc, double a) {
    if (Vc == 0.) { if (a == 0.) return 0.; return pymax(0., a / fabs(a)); }
    if (a > 0) return 1;
    return pymax(pymin(-Vc / (x->h->c.dt * a), 1.), 0.);
}
/* car.follow_action SC:604-624 / car_follower.follow_action C4:109-129 */
static double follow_action(const Ctx *x, const OCar *c, double lead_Sc, double lead_Vc) {
    const double *cb = x->h->c.car_b;
    double speed_car = c->Vc;
    double diff_dist = lead_Sc - c->Sc;
    double delta_v = speed_car - lead_Vc;
    double s = 2. + (speed_car * 2.0) + (speed_car * delta_v) / (2 * sqrt(-cb[0] * cb[2]));
    return cb[2] * (1 - pow(speed_car / 10., 4) - pow(s / diff_dist, 2));
}
/* car.step SC:627-650 (ST:599-618 with the extra clamp; car_follower.transform C4:79-93) */
static void car_move(const Ctx *x, OCar *c, double action, double action_light) {
    const double *cb = x->h->c.car_b; const double dt = x->h->c.dt;
    double acc = pymin(pymax(action, cb[0]), cb[2]);                       /* acceleration() SC:589-590 */
    double sg = sigma(x, c->Vc, acc);
    if (x->h->c.variant == MHO_STOP)
        if (sg > 0.) acc = pymax(acc, -c->Vc / (dt * sg));                 /* ST:604-605 */
    double final_acc = 0.0 + 1.0 * acc;                                    /* discount_array [1,0,0] SC:641-644 */
    final_acc = final_acc * sg;
    double speed = c->Vc + dt * final_acc;
    double pos = (final_acc * pow(dt, 2.0) / 2.0) + (c->Vc * dt) + (c->Sc);
    c->Ac = final_acc; c->Vc = speed; c->Sc = pos; c->light = action_light;
}

/* ------------------------------------------------------------------ reset */
/* pedestrian.__init__, SC:15-104 (draw order is part of the contract) */
static void ped_init(Ctx *x, OPed *p, int is_crossing, int exist) {
    const mho_cfg *c = &x->h->c; const int v = c->variant, L = c->nb_lines;
    const double *pb = c->ped_b; const double W = L * x->e->cross;
    memset(p, 0, sizeof(*p));
    p->is_crossing = is_crossing; p->exist = exist;
    (void)rng_randint(x, 0, 20);                                           /* time_to_remove SC:43 (dead) */
    p->follow_rule = (v == MHO_NAIF) ? (rng_randint(x, 0, 9) < 10) : (rng_randint(x, 0, 9) < 3); /* SC:48, NA:46 */
    p->dir = 2 * rng_randint(x, 0, 1) - 1;                                 /* SC:54 */
    p->line_pos = L * (p->dir < 0) - 1 * (p->dir > 0);                     /* SC:55 */
    p->v0x = rng_uniform(x, pb[0], pb[4]);                                 /* SC:56 */
    p->v0y = rng_uniform(x, pb[1], pb[5]) * p->dir;                        /* SC:57 */
    p->Spx = rng_uniform(x, pb[2], pb[6]);                                 /* SC:61 */
    p->Spy = (rng_uniform(x, pb[3], pb[7]) - W / 2.) * p->dir;             /* SC:62 */
    if (!exist) {                                                          /* SC:66-68 */
        p->v0x = 0.; p->v0y = 0.;
        if (v == MHO_COOP || v == MHO_SCALABLE) p->Spy = pb[3] * p->dir;   /* CO:68 */
        else { p->Spx = pb[2]; p->Spy = pb[3] * p->dir; }                  /* ST:68, NA:66, C4:188 */
    } else if (!is_crossing) { p->v0x = 0.; p->v0y = 0.; p->dir = 0; }     /* SC:69-71 */
    p->Vpx = p->v0x; p->Vpy = p->v0y;                                      /* SC:74 */
    p->gender = rng_randint(x, 0, 1);                                      /* SC:84 */
    p->age = rng_randint(x, 0, 2);                                         /* SC:85 */
    (void)cg_score(x, p, x->e->cross);                                     /* SC:86: self.CG is dead, the draw is not */
    p->delta = 0.0; p->worst_dl = 0.0;
    if (v == MHO_STOP || v == MHO_4CARS2 || v == MHO_SCALABLE)
        p->need_to_stop = rng_uniform(x, 0, 1) < 0.5;                      /* SC:91, ST:90, C42:212 */
    else if (v == MHO_COOP) p->need_to_stop = 1;                           /* CO:90 (no branch reads it) */
    if (v == MHO_STOP || v == MHO_COOP || v == MHO_4CARS2 || v == MHO_SCALABLE)
        p->cross_stop = rng_uniform(x, -W / 2 + 0.2, W / 2 - 0.2);         /* SC:92, CO:91 */
}

/* car.__init__, SC:531-581; `arg` is the constructor's line / index argument */
static void car_init(Ctx *x, OCar *c, int arg, int exist) {
    const mho_cfg *g = &x->h->c; const double *pb = g->ped_b;
    const double W = g->nb_lines * x->e->cross, initial_speed = 10;
    memset(c, 0, sizeof(*c));
    c->line = (g->variant == MHO_SCALABLE) ? arg / 2 : arg;                /* SC:537 vs CO:521 */
    c->Vc = initial_speed;
    c->Ts = (g->variant == MHO_SCALABLE) ? 0. : -10.;                      /* SC:555 / CO:539 */
    double mean_speed_ped = pb[1] + pb[5] / 2;                             /* SC:565 */
    double finish_crosslines_time = (W * initial_speed) / (mean_speed_ped);
    double low_car_range = (pb[3] * initial_speed) / pb[1];
    double high_car_range = (pb[7] * initial_speed) / pb[5];
    double u = rng_uniform(x, low_car_range - finish_crosslines_time, high_car_range);
    c->Sc = (g->variant == MHO_SCALABLE) ? u - 20.0 * (arg % 2) : u;        /* SC:576 / CO:560 */
    c->exist = exist;
}

static void store_round(const OHandle *h, OEnv *e) {
    if (!h->c.store_f32) return;   /* cross stays fp64 in HBM (env_state.cuh) */
    for (int i = 0; i < h->C; ++i) {
        OCar *c = &e->car[i];
        c->Ac = f32r(c->Ac); c->Vc = f32r(c->Vc); c->Sc = f32r(c->Sc); c->light = f32r(c->light);
        c->pa = f32r(c->pa); c->es = f32r(c->es); c->Ts = f32r(c->Ts);
    }
    for (int j = 0; j < h->c.nb_ped; ++j) {
        OPed *p = &e->ped[j];
        /* HBM format rule "symbolic positions" (mh-ppo_b200/csrc/env_state.cuh): an Sp_y whose fp32
         * rounding equals that of the kerb snap (SC:326) or of the lane-edge snap (SC:392-397, dtc
         * SC:350-351) is kept as that exact fp64 value; anything else is rounded to fp32. */
        const double L = h->c.nb_lines, W = L * e->cross;
        const double kerb = -p->dir * W / 2.;
        double dtc = (L - p->line_pos - 1) * e->cross * (p->dir > 0);
        dtc += (p->line_pos) * e->cross * (p->dir < 0);
        const double lane = p->dir * ((W - dtc) - W / 2.);
        if ((float)p->Spy == (float)kerb) p->Spy = kerb;
        else if ((float)p->Spy == (float)lane) p->Spy = lane;
        else p->Spy = f32r(p->Spy);
        p->Vpx = f32r(p->Vpx); p->Vpy = f32r(p->Vpy); p->Spx = f32r(p->Spx);
        p->v0x = f32r(p->v0x); p->v0y = f32r(p->v0y); p->cross_stop = f32r(p->cross_stop);
        p->delta = f32r(p->delta); p->worst_dl = f32r(p->worst_dl);
    }
}

/* reset, SC:884-946 (CO:838-892, C4:850-911, C42:866-927) */
static void env_reset(const OHandle *h, OEnv *e, float *obs) {
    Ctx x = { h, e };
    const mho_cfg *c = &h->c; const int v = c->variant, P = c->nb_ped, L = c->nb_lines;
    e->cross = rng_uniform(&x, c->cross_b[0], c->cross_b[1]);              /* SC:890 */
    for (int j = 0; j < P; ++j) ped_init(&x, &e->ped[j], 0, 0);            /* SC:897-899 */
    if (v == MHO_SCALABLE) {
        for (int i = 0; i < 2 * L; ++i) car_init(&x, &e->car[i], i / 2, 0); /* SC:900-901 */
        e->car_traffic = rng_randint(&x, 1, c->nb_car);                    /* SC:902 */
        int pool[MHO_MAXC], n = 2 * L;                                     /* random.sample SC:903 */
        for (int i = 0; i < n; ++i) pool[i] = i;
        for (int i = 0; i < e->car_traffic; ++i) {
            int j = i + (int)floor(rng_random(&x) * (n - i));
            if (j > n - 1) j = n - 1;
            int t = pool[i]; pool[i] = pool[j]; pool[j] = t;
        }
        for (int k = 0; k < e->car_traffic; ++k)                           /* SC:904-906 */
            car_init(&x, &e->car[pool[k]], pool[k] / 2, 1);
    } else {
        for (int i = 0; i < c->nb_car; ++i) car_init(&x, &e->car[i], i % L, 1); /* CO:852-853 */
        e->car_traffic = c->nb_car;
        if (v == MHO_4CARS || v == MHO_4CARS2)                             /* C4:868-869 */
            for (int i = 0; i < c->nb_car; ++i) car_init(&x, &e->car[c->nb_car + i], e->car[i].line, 1);
    }
    e->ped_traffic = rng_randint(&x, 1, P);                                /* SC:912 */
    for (int j = 0; j < e->ped_traffic; ++j) ped_init(&x, &e->ped[j], 1, 1); /* SC:913-915 */
    if (v == MHO_4CARS || v == MHO_4CARS2)                                 /* C4:878-879, C42:894-895 */
        for (int i = 0; i < c->nb_car; ++i) {
            OCar *f = &e->car[c->nb_car + i];
            double gap = (v == MHO_4CARS2) ? rng_uniform(&x, 10, 30) : 15.;
            f->Vc = 10; f->Sc = e->car[i].Sc - gap; f->light = 0; f->line = e->car[i].line;
        }
    e->step_idx = 0;
    if (obs) write_obs(&x, obs, 1);
    else { float tmp[MHO_MAXC * 7 + 4 + MHO_MAXP * 9]; write_obs(&x, tmp, 1); } /* get_data mutates delta */
    store_round(h, e);
}

/* State injection.
 * reset_pedestrian, SC:948-955 (CO:895-902, ST:904-911, C4:913-920, C42:929-936): EVERY pedestrian is rebuilt as a placeholder
 * (constructor draws consumed, SC:949-951), then pedestrian.reset_ped (SC:106-137) on slot num_ped with
 * q = (speed_x, speed_y, pos_x, pos_y, dl, leave, CZ, exist, direction); `leave` / `CZ` are stored as their truth value.
 * naif (NA:897-898 -> NA:100-135): no rebuild, q = (speed_x, speed_y, pos_x, pos_y, dl, direction, cross): the pedestrian's cross
 * becomes `cross` (here: the env's, which every pedestrian shares), speed_y and pos_y are relative to the direction / kerb. */
static void env_reset_pedestrian(const OHandle *h, OEnv *e, int num_ped, const double *q) {
    Ctx x = { h, e };
    const mho_cfg *c = &h->c; const int L = c->nb_lines;
    OPed *p = &e->ped[num_ped];
    if (c->variant == MHO_NAIF) {
        e->cross = q[6];                                                   /* NA:101-102 */
        const double W = L * e->cross;
        p->dir = (int)q[5];                                                /* NA:103 */
        p->v0x = q[0]; p->v0y = q[1] * p->dir;                             /* NA:104 */
        p->Spx = q[2]; p->Spy = (q[3] - W / 2.) * p->dir;                  /* NA:106 */
        p->delta = q[4]; p->exist = 1; p->ped_left = 0; p->ped_in_cross = 0;   /* NA:115-120 */
    } else {
        for (int j = 0; j < c->nb_ped; ++j) ped_init(&x, &e->ped[j], 0, 0);    /* SC:949-951 */
        p->v0x = q[0]; p->v0y = q[1]; p->Spx = q[2]; p->Spy = q[3];        /* SC:107-110 */
        p->dir = (int)q[8];                                                /* SC:115 */
        p->delta = q[4]; p->exist = q[7] != 0.0; p->ped_left = q[5] != 0.0; p->ped_in_cross = q[6] != 0.0;   /* SC:117-123 */
    }
    p->Vpx = p->v0x; p->Vpy = p->v0y;
    p->ratio_eps = 1;                                                      /* SC:111 */
    p->line_pos = L * (p->dir < 0) - 1 * (p->dir > 0);                     /* SC:116 */
    p->is_crossing = 1;                                                    /* SC:124 */
    (void)cg_score(&x, p, L * e->cross);                                   /* SC:125: the value is dead, the draw is not */
    p->t0c = 0;                                                            /* SC:127 */
    store_round(h, e);
}
/* reset_cars -> car.reset_car, SC:957-958, 583-587: q = (speed_x, pos_x, light, line) */
static void env_reset_car(const OHandle *h, OEnv *e, int num_car, const double *q) {
    OCar *c = &e->car[num_car];
    c->Sc = q[1]; c->Vc = q[0]; c->light = q[2]; c->line = (int)q[3];
    store_round(h, e);
}
void mho_reset_pedestrian(void *hh, int num_ped, const double *params, const uint8_t *mask) {
    OHandle *h = (OHandle *)hh;
    for (int64_t n = 0; n < h->N; ++n) if (!mask || mask[n]) env_reset_pedestrian(h, &h->env[n], num_ped, params + n * 9);
}
void mho_reset_cars(void *hh, int num_car, const double *params, const uint8_t *mask) {
    OHandle *h = (OHandle *)hh;
    for (int64_t n = 0; n < h->N; ++n) if (!mask || mask[n]) env_reset_car(h, &h->env[n], num_car, params + n * 4);
}
/* get_state, SC:960-969: the observation of the current state (all car slots are handed to get_data, as in reset) */
void mho_observe(void *hh, float *obs) {
    OHandle *h = (OHandle *)hh;
    for (int64_t n = 0; n < h->N; ++n) { Ctx x = { h, &h->env[n] }; write_obs(&x, obs + n * h->nobs, 1); store_round(h, &h->env[n]); }
}

/* step, SC:789-878 (CO:745-832, C4:783-844, C42:799-860) */
static int env_step(const OHandle *h, OEnv *e, const double *act, float *obs, double *rewards,
                    double *reward_light) {
    Ctx x = { h, e };
    const mho_cfg *c = &h->c; const int v = c->variant, P = c->nb_ped, L = c->nb_lines, C = h->C;
    const double time = e->step_idx * c->dt;
    double prev_Sc[MHO_MAXC];
    for (int i = 0; i < h->nlead; ++i) prev_Sc[i] = e->car[i].Sc;          /* SC:797 */
    if (v == MHO_SCALABLE) {                                               /* SC:798-802 */
        for (int i = 0; i < 2 * L; ++i) {
            double a_idm = 2.;
            if (i % 2 == 1 && e->car[i - 1].exist && e->car[i].exist)
                a_idm = follow_action(&x, &e->car[i], e->car[i - 1].Sc, e->car[i - 1].Vc);
            car_move(&x, &e->car[i], pymin(a_idm, act[i]), act[i + 2 * L]);
        }
    } else if (v == MHO_4CARS || v == MHO_4CARS2) {
        const int n = c->nb_car;
        for (int i = 0; i < n; ++i)                                        /* C4:792-793, C42:808-809 */
            car_move(&x, &e->car[i], pymin(2., act[i]), act[i + (v == MHO_4CARS2 ? 2 * n : n)]);
        for (int i = 0; i < n; ++i) {                                      /* C4:794-795, C42:810-811 */
            OCar *f = &e->car[n + i];
            double a_idm = follow_action(&x, f, e->car[i].Sc, e->car[i].Vc);
            if (v == MHO_4CARS2) car_move(&x, f, pymin(a_idm, act[i + n]), act[i + 3 * n]); /* C42:75-79 */
            else car_move(&x, f, a_idm, e->car[i].light);                  /* C4:75-77 */
        }
    } else {
        for (int i = 0; i < C; ++i)                                        /* CO:754-755 */
            car_move(&x, &e->car[i], pymin(2., act[i]), act[i + C]);
    }
    /* lists handed to pedestrian.step / get_data: SC:803-806 */
    double Sc[MHO_MAXC], Vc[MHO_MAXC], light[MHO_MAXC]; int line[MHO_MAXC]; int n = 0;
    for (int i = 0; i < C; ++i) {
        if (v == MHO_SCALABLE && !e->car[i].exist) continue;
        Sc[n] = e->car[i].Sc; Vc[n] = e->car[i].Vc; light[n] = e->car[i].light; line[n] = e->car[i].line; ++n;
    }
    for (int j = 0; j < P; ++j) ped_step(&x, &e->ped[j], time, n, Vc, Sc, line, light); /* SC:808-809 */
    for (int i = 0; i < h->nlead; ++i) reward_light[i] = 0.;               /* SC:841-846 */
    for (int j = 0; j < P; ++j) {
        double det[MHO_MAXC];
        detection(&x, &e->ped[j], h->nlead, prev_Sc, det);
        int add = e->ped[j].is_crossing && (v != MHO_SCALABLE || e->ped[j].exist);
        if (add) for (int i = 0; i < h->nlead; ++i) reward_light[i] += det[i];
    }
    for (int i = 0; i < h->nlead; ++i) {                                   /* SC:849-858 */
        OCar *ci = &e->car[i];
        double d = (ci->Vc - 10);
        double reward = -10. * pow(d, 2) / 100;                            /* SC:657-665 */
        if (!(ci->light <= 0.0)) {
            int any = 0; double m = 0;
            for (int j = 0; j < P; ++j) {
                if (!e->ped[j].exist) continue;
                double r = reward_wait_safety(&x, &e->ped[j], ci->Vc, ci->Sc, ci->line);
                if (!any || r < m) { m = r; any = 1; }
            }
            if (any) reward += m;
        }
        rewards[i] = reward;
    }
    write_obs(&x, obs, 0);                                                 /* SC:868-871 */
    int done = (e->step_idx >= h->done_idx) || (e->ped_traffic <= 0);      /* SC:874 */
    e->step_idx += 1;                                                      /* SC:875 */
    store_round(h, e);
    return done;
}

/* ------------------------------------------------------------------ C API */
int mho_create(const mho_cfg *cfg, int64_t n_envs, int64_t env_id0, void **handle) {
    if (!cfg || !handle || n_envs <= 0) return -1;
    OHandle *h = (OHandle *)calloc(1, sizeof(OHandle));
    h->c = *cfg;
    h->C = mho_n_slots(cfg); h->nlead = mho_n_lead(cfg); h->A = mho_n_action(cfg); h->nobs = mho_n_obs(cfg);
    if (h->C > MHO_MAXC || cfg->nb_ped > MHO_MAXP || cfg->nb_ped < 1 || h->C < 1) { free(h); return -2; }
    if (cfg->variant == MHO_SCALABLE && cfg->nb_car > 2 * cfg->nb_lines) { free(h); return -3; }
    /* done = time >= (max_episode-1)*dt with time the fp64 running sum of dt (SC:874-875,945) */
    double t = 0.0, lim = (cfg->max_episode - 1) * cfg->dt; int k = 0;
    while (!(t >= lim) && k < 1000000) { t = t + cfg->dt; ++k; }
    h->done_idx = k;
    h->N = n_envs;
    h->env = (OEnv *)calloc((size_t)n_envs, sizeof(OEnv));
    for (int64_t i = 0; i < n_envs; ++i) h->env[i].env_id = (uint64_t)(env_id0 + i);
    *handle = h;
    return 0;
}

void mho_destroy(void *hh) {
    OHandle *h = (OHandle *)hh;
    if (!h) return;
    free(h->env); free(h);
}

/* batched calls: static partition of the env range over pthreads (no OpenMP dependency) */
typedef struct {
    OHandle *h; int64_t lo, hi; int op;
    const uint8_t *mask; const double *actions; float *obs, *term_obs; double *rewards, *reward_light;
    uint8_t *done; int autoreset;
} Job;

static void job_run(Job *j) {
    OHandle *h = j->h;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        if (j->op == 0) {
            if (j->mask && !j->mask[i]) continue;
            env_reset(h, &h->env[i], j->obs ? j->obs + i * h->nobs : NULL);
        } else {
            float *o = j->obs + i * h->nobs;
            int d = env_step(h, &h->env[i], j->actions + i * h->A, o, j->rewards + i * h->nlead,
                             j->reward_light + i * h->nlead);
            j->done[i] = (uint8_t)d;
            if (j->term_obs) memcpy(j->term_obs + i * h->nobs, o, sizeof(float) * h->nobs);
            if (d && j->autoreset) env_reset(h, &h->env[i], o);
        }
    }
}
/* Persistent worker pool (one per process, grown on demand): the workers sleep on a condition variable between
 * batched calls, so a step does not pay thread creation; envs are handed out in chunks from a shared cursor. */
#define POOL_MAX 256
#define POOL_CHUNK 256
typedef struct {
    pthread_mutex_t mu; pthread_cond_t go, fin;
    pthread_t th[POOL_MAX];
    int n_threads, active, pending; unsigned long gen;
    Job job; int64_t N; volatile int64_t cursor;
} Pool;
static Pool g_pool = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER };

static void pool_work(Pool *p) {
    for (;;) {
        const int64_t lo = __atomic_fetch_add(&p->cursor, POOL_CHUNK, __ATOMIC_RELAXED);
        if (lo >= p->N) break;
        Job j = p->job;
        j.lo = lo; j.hi = lo + POOL_CHUNK < p->N ? lo + POOL_CHUNK : p->N;
        job_run(&j);
    }
}
static void *pool_thread(void *arg) {
    Pool *p = &g_pool; const int id = (int)(intptr_t)arg;
    unsigned long seen = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (p->gen == seen) pthread_cond_wait(&p->go, &p->mu);
        seen = p->gen;
        if (id >= p->active) continue;
        pthread_mutex_unlock(&p->mu);
        pool_work(p);
        pthread_mutex_lock(&p->mu);
        if (--p->pending == 0) pthread_cond_signal(&p->fin);
    }
    return NULL;
}

static void run_parallel(Job *proto) {
    OHandle *h = proto->h;
    int nt = h->c.n_threads > 0 ? h->c.n_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > POOL_MAX) nt = POOL_MAX;
    if ((int64_t)nt * POOL_CHUNK > h->N) nt = (int)((h->N + POOL_CHUNK - 1) / POOL_CHUNK);
    if (nt <= 1) { proto->lo = 0; proto->hi = h->N; job_run(proto); return; }
    Pool *p = &g_pool;
    pthread_mutex_lock(&p->mu);
    while (p->n_threads < nt - 1) {              /* the caller is worker 0 */
        pthread_create(&p->th[p->n_threads], NULL, pool_thread, (void *)(intptr_t)p->n_threads);
        pthread_detach(p->th[p->n_threads]);
        p->n_threads++;
    }
    p->job = *proto; p->N = h->N; p->cursor = 0; p->active = nt - 1; p->pending = nt - 1; p->gen++;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->mu);
    pool_work(p);
    pthread_mutex_lock(&p->mu);
    while (p->pending != 0) pthread_cond_wait(&p->fin, &p->mu);
    pthread_mutex_unlock(&p->mu);
}

void mho_reset(void *hh, const uint8_t *mask, float *obs) {
    Job j; memset(&j, 0, sizeof(j));
    j.h = (OHandle *)hh; j.op = 0; j.mask = mask; j.obs = obs;
    run_parallel(&j);
}

void mho_step(void *hh, const double *actions, float *obs, double *rewards, double *reward_light,
              uint8_t *done, int autoreset, float *term_obs) {
    Job j; memset(&j, 0, sizeof(j));
    j.h = (OHandle *)hh; j.op = 1; j.actions = actions; j.obs = obs; j.rewards = rewards;
    j.reward_light = reward_light; j.done = done; j.autoreset = autoreset; j.term_obs = term_obs;
    run_parallel(&j);
}

static unsigned pack_flags(const OPed *p) {
    unsigned f = 0;
    f |= (unsigned)(p->exist != 0) << F_EXIST; f |= (unsigned)(p->is_crossing != 0) << F_IS_CROSSING;
    f |= (unsigned)(p->decision != 0) << F_DECISION; f |= (unsigned)(p->at_crossing != 0) << F_AT_CROSSING;
    f |= (unsigned)(p->ped_left != 0) << F_PED_LEFT; f |= (unsigned)(p->ped_in_cross != 0) << F_PED_IN_CROSS;
    f |= (unsigned)(p->not_waiting != 0) << F_NOT_WAITING; f |= (unsigned)(p->accident != 0) << F_ACCIDENT;
    f |= (unsigned)(p->worst_acc != 0) << F_WORST_ACC; f |= (unsigned)(p->follow_rule != 0) << F_FOLLOW_RULE;
    f |= (unsigned)(p->stop != 0) << F_STOP; f |= (unsigned)(p->need_to_stop != 0) << F_NEED_TO_STOP;
    f |= (unsigned)(p->ratio_eps != 0) << F_RATIO_EPS;
    return f;
}

void mho_get_state(void *hh, double *car_f, int32_t *car_i, double *ped_f, int32_t *ped_i,
                   double *env_f, int64_t *env_i) {
    OHandle *h = (OHandle *)hh; const int C = h->C, P = h->c.nb_ped;
    for (int64_t n = 0; n < h->N; ++n) {
        const OEnv *e = &h->env[n];
        for (int i = 0; i < C; ++i) {
            const OCar *c = &e->car[i]; double *f = car_f + (n * C + i) * 7; int32_t *q = car_i + (n * C + i) * 2;
            f[0] = c->Ac; f[1] = c->Vc; f[2] = c->Sc; f[3] = c->light; f[4] = c->pa; f[5] = c->es; f[6] = c->Ts;
            q[0] = c->line; q[1] = c->exist;
        }
        for (int j = 0; j < P; ++j) {
            const OPed *p = &e->ped[j]; double *f = ped_f + (n * P + j) * 9; int32_t *q = ped_i + (n * P + j) * 9;
            f[0] = p->Vpx; f[1] = p->Vpy; f[2] = p->Spx; f[3] = p->Spy; f[4] = p->v0x; f[5] = p->v0y;
            f[6] = p->cross_stop; f[7] = p->delta; f[8] = p->worst_dl;
            q[0] = p->t0c; q[1] = p->waitc; q[2] = p->crossc; q[3] = p->time_stop; q[4] = p->line_pos;
            q[5] = p->dir; q[6] = p->gender; q[7] = p->age; q[8] = (int32_t)pack_flags(p);
        }
        env_f[n] = e->cross;
        env_i[n * 4 + 0] = e->step_idx; env_i[n * 4 + 1] = e->ped_traffic; env_i[n * 4 + 2] = e->car_traffic;
        env_i[n * 4 + 3] = e->ctr;
    }
}

void mho_set_state(void *hh, const double *car_f, const int32_t *car_i, const double *ped_f,
                   const int32_t *ped_i, const double *env_f, const int64_t *env_i) {
    OHandle *h = (OHandle *)hh; const int C = h->C, P = h->c.nb_ped;
    for (int64_t n = 0; n < h->N; ++n) {
        OEnv *e = &h->env[n];
        for (int i = 0; i < C; ++i) {
            OCar *c = &e->car[i]; const double *f = car_f + (n * C + i) * 7; const int32_t *q = car_i + (n * C + i) * 2;
            c->Ac = f[0]; c->Vc = f[1]; c->Sc = f[2]; c->light = f[3]; c->pa = f[4]; c->es = f[5]; c->Ts = f[6];
            c->line = q[0]; c->exist = q[1];
        }
        for (int j = 0; j < P; ++j) {
            OPed *p = &e->ped[j]; const double *f = ped_f + (n * P + j) * 9; const int32_t *q = ped_i + (n * P + j) * 9;
            p->Vpx = f[0]; p->Vpy = f[1]; p->Spx = f[2]; p->Spy = f[3]; p->v0x = f[4]; p->v0y = f[5];
            p->cross_stop = f[6]; p->delta = f[7]; p->worst_dl = f[8];
            p->t0c = q[0]; p->waitc = q[1]; p->crossc = q[2]; p->time_stop = q[3]; p->line_pos = q[4];
            p->dir = q[5]; p->gender = q[6]; p->age = q[7];
            unsigned fl = (unsigned)q[8];
            p->exist = (fl >> F_EXIST) & 1; p->is_crossing = (fl >> F_IS_CROSSING) & 1;
            p->decision = (fl >> F_DECISION) & 1; p->at_crossing = (fl >> F_AT_CROSSING) & 1;
            p->ped_left = (fl >> F_PED_LEFT) & 1; p->ped_in_cross = (fl >> F_PED_IN_CROSS) & 1;
            p->not_waiting = (fl >> F_NOT_WAITING) & 1; p->accident = (fl >> F_ACCIDENT) & 1;
            p->worst_acc = (fl >> F_WORST_ACC) & 1; p->follow_rule = (fl >> F_FOLLOW_RULE) & 1;
            p->stop = (fl >> F_STOP) & 1; p->need_to_stop = (fl >> F_NEED_TO_STOP) & 1;
            p->ratio_eps = (fl >> F_RATIO_EPS) & 1;
        }
        e->cross = env_f[n];
        e->step_idx = (int)env_i[n * 4 + 0]; e->ped_traffic = (int)env_i[n * 4 + 1];
        e->car_traffic = (int)env_i[n * 4 + 2]; e->ctr = (uint32_t)env_i[n * 4 + 3];
        store_round(h, e);
    }
}
