// stats.cuh -- episode statistics of the evaluation records on the device.
//
// Replaces the per-step Python loops of the reference's analysis cell `get_average` (PY:1550-1675) and the waiting-time
// histogram input (PY:1390-1403): for every (episode, env) of the dense record written by Env_rollout.iterations
// (obs [E][T][n_obs][N], the state BEFORE each step, component-major) one thread walks the T rows once and emits
//   per pedestrian slot : waiting time on the kerb (PY:1591-1596), time until it has crossed (PY:1597-1601)
//   per car slot        : time until it is 25 m past the pedestrians (PY:1622-1631), last light decision (PY:1620), light after
//                         the first step (PY:1613, 1621), "could have stopped" flag (PY:1613-1617), free-flow time to 25 m
//                         (PY:1612)
//   sums                : car-0 speed / acceleration and pedestrian-0 speed moments over the rows (PY:1552-1562), speed of the
//                         yielding cars (PY:1621)
// The means / standard deviations / scenario frequencies the reference prints are reductions of these arrays, done by the
// host mirror (Env_rollout.get_average) with torch on the device.  Time is counted in steps (the reference multiplies by
// 0.3 = dt).  Memory-bound: every record word is read once, coalesced over envs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mhppo {

struct StatsDims { int E, T, C, P, car_w, env_w, n_obs; int64_t N; };
struct StatsOut {
    float *ped;                                        // [2][E][P][N]: waiting time, crossing time (seconds)
    float *car;                                        // [8][E][C][N]: leave time, free-flow time, last light, light after step 1,
                                                       //   could-stop (-1 n/a, 0, 1), number / sum / sum of squares of the 25 m
                                                       //   passage times (PY:1624-1625 appends one per passage)
    double *sums;                                      // [E][N][8]: n_rows, sum v0, sum v0^2, sum |a0|, sum a0, sum a0^2, sum |vp0|, sum vp0^2
    double *yield_speed;                               // [E][N][3]: n, sum v, sum v^2 over rows of cars whose light1 == 1
};

__global__ void __launch_bounds__(128) k_episode_stats(StatsDims d, const float *__restrict__ obs, float dt, StatsOut o) {
    const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int e = blockIdx.y;
    if (n >= d.N) return;
    const int64_t row = (int64_t)d.n_obs * d.N;                       // one time step of one episode
    const float *base = obs + (int64_t)e * d.T * row + n;
    auto at = [&](int t, int k) { return base[(int64_t)t * row + (int64_t)k * d.N]; };
    const int env0 = d.car_w * d.C, ped0 = env0 + d.env_w;
    // rows of one episode share `cross` (ep_cross[t] == ep_cross[t+1], PY:1587): the walk is t = 0 .. T-2
    for (int p = 0; p < d.P; ++p) {
        float wait = 0.f, leave = 0.f;
        for (int t = 0; t + 1 < d.T; ++t) {
            const float y0 = at(t, ped0 + 9 * p + 3), y1 = at(t + 1, ped0 + 9 * p + 3);
            const float c0 = at(t, env0), c1 = at(t + 1, env0);
            const float d0 = at(t, ped0 + 9 * p + 8), d1 = at(t + 1, ped0 + 9 * p + 8);
            if (fabsf(y0) == c0 && fabsf(y1) == c1) wait += 1.f;                       // PY:1591
            if (fabsf(y0) == 0.f && fabsf(y1) == 0.f) wait += 1.f;                     // PY:1594
            if (y0 * d0 < c0 && y1 * d1 >= c1) leave = (float)t;                       // PY:1597-1598
        }
        if (leave == 0.f) leave = (float)(d.T - 1);                                    // PY:1600-1601
        const int64_t q = ((int64_t)e * d.P + p) * d.N + n, PS = (int64_t)d.E * d.P * d.N;
        o.ped[q] = wait * dt; o.ped[PS + q] = leave * dt;
    }
    double ys_n = 0.0, ys_s = 0.0, ys_ss = 0.0;
    for (int i = 0; i < d.C; ++i) {
        const float x0 = at(0, d.car_w * i + 3), v0 = at(0, d.car_w * i + 1), light1 = at(1, d.car_w * i + 4);
        float leave = 0.f, decision = 0.f, en = 0.f, es = 0.f, ess = 0.f;
        for (int t = 0; t + 1 < d.T; ++t) {
            decision = at(t, d.car_w * i + 4);                                         // PY:1620
            if (light1 == 1.f) { const double v = (double)at(t, d.car_w * i + 1); ys_n += 1.0; ys_s += v; ys_ss += v * v; }   // PY:1621
            float mx = at(t + 1, ped0 + 2);
            for (int p = 1; p < d.P; ++p) mx = fmaxf(mx, at(t + 1, ped0 + 9 * p + 2));
            const float a0 = at(t, d.car_w * i + 3) - mx - 25.f, a1 = at(t + 1, d.car_w * i + 3) - mx - 25.f;
            if (a0 < 0.f && a1 >= 0.f) { leave = (float)t; en += 1.f; es += (float)t * dt; ess += ((float)t * dt) * ((float)t * dt); }   // PY:1622-1625
        }
        if (leave == 0.f) leave = (float)(d.T - 1);                                    // PY:1630-1631
        const int64_t q = ((int64_t)e * d.C + i) * d.N + n, CS = (int64_t)d.E * d.C * d.N;
        o.car[q] = leave * dt; o.car[CS + q] = (25.f - x0) / v0;                       // PY:1612
        o.car[2 * CS + q] = decision; o.car[3 * CS + q] = light1;
        o.car[4 * CS + q] = (light1 < 0.f) ? ((-x0 - (v0 * v0 / 8.f + v0) > 0.f) ? 1.f : 0.f) : -1.f;   // PY:1613-1617
        o.car[5 * CS + q] = en; o.car[6 * CS + q] = es; o.car[7 * CS + q] = ess;
    }
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < d.T; ++t) {                                                    // PY:1552-1562: car 0 / pedestrian 0, every row
        const double v = (double)at(t, 1), a = (double)at(t, 0), vp = fabs((double)at(t, ped0 + 1));
        s[0] += 1.0; s[1] += v; s[2] += v * v; s[3] += fabs(a); s[4] += a; s[5] += a * a; s[6] += vp; s[7] += vp * vp;
    }
    double *sp = o.sums + ((int64_t)e * d.N + n) * 8;
    for (int k = 0; k < 8; ++k) sp[k] = s[k];
    double *yp = o.yield_speed + ((int64_t)e * d.N + n) * 3;
    yp[0] = ys_n; yp[1] = ys_s; yp[2] = ys_ss;
}

}  // namespace mhppo
