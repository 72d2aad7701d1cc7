// env_state.cuh -- HBM layout of the vectorised env state and its load/store into registers.
//
// Struct-of-arrays with the env index innermost and 16-byte groups, so a warp (lane = env) moves
// each group with one fully coalesced 512-byte float4 transaction:
//
//   car_a[slot][env] = (Vc, Sc, light, Ac)                         fp32
//   car_b[slot][env] = (possible_accident, error_scenario, Ts, bits{line:8, exist:1})
//   ped_a[slot][env] = (Sp_x, Sp_y, Vp_x, Vp_y)
//   ped_b[slot][env] = (v0x, v0y, cross_stop, delta)
//   ped_c[slot][env] = (worst_dl, counts{t0/dt:8, waiting/dt:8, crossing/dt:8, time_stop:8},
//                       bits{12 flags, gender:1, age:2, dir+1:2, line_pos+1:5, y_kerb:1, y_lane:1, ratio_eps:1}, spare)
//   env_e[env]       = (cross as fp64 in two words, {step_idx:8, ped_traffic:8, car_traffic:8}, philox_ctr)
//                      cross is the one geometry parameter every threshold derives from (W = L*cross,
//                      lane edges); it stays fp64 so those expressions round exactly as in the reference
//
// = 32*C + 48*P + 16 bytes per env.
//
// Symbolic positions: the reference snaps a pedestrian to exactly -dir*W/2 when it reaches the kerb
// (SC:326) and to a lane edge when it refuses a lane change (SC:397), and later compares Sp_y with
// those same expressions using <, <=, > (SC:320,331,470-476).  W/2 is not an fp32 number in general,
// so such a position is stored as a tag (y_kerb / y_lane) next to its fp32 rounding and rebuilt in
// fp64 from (cross, dir, line_pos) at load; the equality-sensitive flags then match the fp64
// reference.  The tag is derived at store time by an fp32 equality test, so it needs no tracking
// in the step logic and survives import/export of plain fp32 dumps.  Everything the reference keeps per object but never reads on
// the path (time_to_remove, remove, CG, worst_pos_p, time_before_crossing, finish_crossing, the
// deque of previous accelerations, ...) is not stored.
#pragma once
#include "env_core.cuh"

#ifndef __CUDACC__
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
#endif

namespace mhppo {

struct EnvArena {
    float4 *car_a, *car_b, *ped_a, *ped_b, *ped_c, *env_e;
    int64_t N;          // envs (row pitch of every array)
};

MH_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
MH_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

// exact snapped positions: kerb SC:326, lane edge SC:392-397 (dtc from SC:350-351)
MH_HD double sym_kerb(const Geo &g, const PedR &p) { return (double)(-p.dir) * g.W / 2.0; }
MH_HD double sym_lane(const Geo &g, const PedR &p) {
    double dtc = ((g.Lf - (double)p.lpos - 1.0) * g.cross) * (double)(p.dir > 0);
    dtc += ((double)p.lpos * g.cross) * (double)(p.dir < 0);
    return (double)p.dir * ((g.W - dtc) - g.W / 2.0);
}
enum : uint32_t { PB_Y_KERB = 1u << 22, PB_Y_LANE = 1u << 23 };

MH_HD double words_to_double(uint32_t lo, uint32_t hi) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)hi, (int)lo);
#else
    const uint64_t u = ((uint64_t)hi << 32) | lo; double d; memcpy(&d, &u, 8); return d;
#endif
}
MH_HD void double_to_words(double d, uint32_t &lo, uint32_t &hi) {
#ifdef __CUDA_ARCH__
    lo = (uint32_t)__double2loint(d); hi = (uint32_t)__double2hiint(d);
#else
    uint64_t u; memcpy(&u, &d, 8); lo = (uint32_t)u; hi = (uint32_t)(u >> 32);
#endif
}
MH_HD uint32_t pack_env_word(int step, int ped_traffic, int car_traffic) {
    return ((uint32_t)step & 255u) | (((uint32_t)ped_traffic & 255u) << 8) | (((uint32_t)car_traffic & 255u) << 16);
}

MH_HD uint32_t pack_ped_bits(const PedR &p) {
    return (p.fl & 0xFFFu) | ((uint32_t)(p.gender & 1) << 12) | ((uint32_t)(p.age & 3) << 13) |
           ((uint32_t)((p.dir + 1) & 3) << 15) | ((uint32_t)((p.lpos + 1) & 31) << 17) | (((p.fl >> 12) & 1u) << 24);   // bit 24: PF_RATIO_EPS
}
MH_HD void unpack_ped_bits(uint32_t b, PedR &p) {
    p.fl = (b & 0xFFFu) | (((b >> 24) & 1u) << 12); p.gender = (int)((b >> 12) & 1u); p.age = (int)((b >> 13) & 3u);
    p.dir = (int)((b >> 15) & 3u) - 1; p.lpos = (int)((b >> 17) & 31u) - 1;
}
MH_HD uint32_t pack_ped_counts(const PedR &p) {
    return ((uint32_t)p.t0c & 255u) | (((uint32_t)p.waitc & 255u) << 8) | (((uint32_t)p.crossc & 255u) << 16) |
           (((uint32_t)p.tstop & 255u) << 24);
}
MH_HD void unpack_ped_counts(uint32_t b, PedR &p) {
    p.t0c = (int)(b & 255u); p.waitc = (int)((b >> 8) & 255u); p.crossc = (int)((b >> 16) & 255u);
    p.tstop = (int)((b >> 24) & 255u);
}

template <int MC, int MP>
MH_HD void load_env(const EnvArena &a, const EnvConst &c, int64_t n, EnvR<MC, MP> &e) {
    const float4 ee = a.env_e[n];
    e.cross = words_to_double(f2u(ee.x), f2u(ee.y));
    const uint32_t tr = f2u(ee.z);
    e.step = (int)(tr & 255u); e.ped_traffic = (int)((tr >> 8) & 255u); e.car_traffic = (int)((tr >> 16) & 255u);
    e.rng.ctr = f2u(ee.w);
#pragma unroll
    for (int i = 0; i < MC; ++i) {
        if (i >= c.nC) continue;
        const float4 ka = a.car_a[(int64_t)i * a.N + n], kb = a.car_b[(int64_t)i * a.N + n];
        CarR &k = e.car[i];
        k.Vc = (double)ka.x; k.Sc = (double)ka.y; k.light = (double)ka.z; k.Ac = (double)ka.w;
        k.pa = (double)kb.x; k.es = (double)kb.y; k.Ts = (double)kb.z;
        const uint32_t b = f2u(kb.w);
        k.line = (int)(b & 255u); k.exist = (int)((b >> 8) & 1u);
    }
#pragma unroll
    for (int j = 0; j < MP; ++j) {
        if (j >= c.nP) continue;
        const float4 pa = a.ped_a[(int64_t)j * a.N + n], pb = a.ped_b[(int64_t)j * a.N + n],
                     pc = a.ped_c[(int64_t)j * a.N + n];
        PedR &p = e.ped[j];
        p.Spx = (double)pa.x; p.Spy = (double)pa.y; p.Vpx = (double)pa.z; p.Vpy = (double)pa.w;
        p.v0x = (double)pb.x; p.v0y = (double)pb.y; p.cstop = (double)pb.z; p.delta = (double)pb.w;
        p.wdl = (double)pc.x;
        unpack_ped_counts(f2u(pc.y), p);
        const uint32_t bits = f2u(pc.z);
        unpack_ped_bits(bits, p);
        if (bits & (PB_Y_KERB | PB_Y_LANE)) {
            const Geo g = make_geo(e.cross, c.L);
            p.Spy = (bits & PB_Y_KERB) ? sym_kerb(g, p) : sym_lane(g, p);
        }
    }
}

// tag of a pedestrian whose fp32 Sp_y equals the fp32 rounding of a snapped position
MH_HD uint32_t sym_tag(const Geo &g, const PedR &p, float spy32) {
    // both candidates are a handful of operations: selects, so that the lanes of a warp do not split here
    const bool kerb = spy32 == (float)sym_kerb(g, p), lane = spy32 == (float)sym_lane(g, p);
    return kerb ? PB_Y_KERB : (lane ? PB_Y_LANE : 0u);
}
template <int MC, int MP>
MH_HD void store_env(const EnvArena &a, const EnvConst &c, int64_t n, const EnvR<MC, MP> &e) {
    uint32_t clo, chi;
    double_to_words(e.cross, clo, chi);
    a.env_e[n] = make_float4(u2f(clo), u2f(chi), u2f(pack_env_word(e.step, e.ped_traffic, e.car_traffic)), u2f(e.rng.ctr));
#pragma unroll
    for (int i = 0; i < MC; ++i) {
        if (i >= c.nC) continue;
        const CarR &k = e.car[i];
        a.car_a[(int64_t)i * a.N + n] = make_float4((float)k.Vc, (float)k.Sc, (float)k.light, (float)k.Ac);
        a.car_b[(int64_t)i * a.N + n] = make_float4((float)k.pa, (float)k.es, (float)k.Ts,
                                                    u2f(((uint32_t)k.line & 255u) | (((uint32_t)k.exist & 1u) << 8)));
    }
    const Geo g = make_geo(e.cross, c.L);
#pragma unroll
    for (int j = 0; j < MP; ++j) {
        if (j >= c.nP) continue;
        const PedR &p = e.ped[j];
        a.ped_a[(int64_t)j * a.N + n] = make_float4((float)p.Spx, (float)p.Spy, (float)p.Vpx, (float)p.Vpy);
        a.ped_b[(int64_t)j * a.N + n] = make_float4((float)p.v0x, (float)p.v0y, (float)p.cstop, (float)p.delta);
        a.ped_c[(int64_t)j * a.N + n] = make_float4((float)p.wdl, u2f(pack_ped_counts(p)),
                                                    u2f(pack_ped_bits(p) | sym_tag(g, p, (float)p.Spy)), 0.f);
    }
}

}  // namespace mhppo

// ---------------------------------------------------------------------------------------------
// canonical state dump <-> arena (include/mhppo.h: mhppo_env_export_state / _import_state); one env
// per thread, runtime slot counts.  This is the vectorised form of the reference's get_state() /
// reset_pedestrian() / reset_cars() attribute access (SC:948-969).
namespace mhppo {

struct DumpPtrs {
    float *car_f; int32_t *car_i; float *ped_f; int32_t *ped_i; double *env_f; int64_t *env_i;
};

MH_HD void export_one(const EnvArena &a, const EnvConst &c, int64_t n, const DumpPtrs &d) {
    const float4 ee = a.env_e[n];
    d.env_f[n] = words_to_double(f2u(ee.x), f2u(ee.y));
    const uint32_t tr = f2u(ee.z);
    d.env_i[n * 4 + 0] = (int64_t)(tr & 255u); d.env_i[n * 4 + 1] = (int64_t)((tr >> 8) & 255u);
    d.env_i[n * 4 + 2] = (int64_t)((tr >> 16) & 255u); d.env_i[n * 4 + 3] = (int64_t)f2u(ee.w);
    for (int i = 0; i < c.nC; ++i) {
        const float4 ka = a.car_a[(int64_t)i * a.N + n], kb = a.car_b[(int64_t)i * a.N + n];
        float *f = d.car_f + (n * c.nC + i) * 7; int32_t *q = d.car_i + (n * c.nC + i) * 2;
        f[0] = ka.w; f[1] = ka.x; f[2] = ka.y; f[3] = ka.z; f[4] = kb.x; f[5] = kb.y; f[6] = kb.z;
        const uint32_t b = f2u(kb.w);
        q[0] = (int32_t)(b & 255u); q[1] = (int32_t)((b >> 8) & 1u);
    }
    for (int j = 0; j < c.nP; ++j) {
        const float4 pa = a.ped_a[(int64_t)j * a.N + n], pb = a.ped_b[(int64_t)j * a.N + n],
                     pc = a.ped_c[(int64_t)j * a.N + n];
        float *f = d.ped_f + (n * c.nP + j) * 9; int32_t *q = d.ped_i + (n * c.nP + j) * 9;
        f[0] = pa.z; f[1] = pa.w; f[2] = pa.x; f[3] = pa.y; f[4] = pb.x; f[5] = pb.y; f[6] = pb.z; f[7] = pb.w;
        f[8] = pc.x;
        PedR p; unpack_ped_counts(f2u(pc.y), p); unpack_ped_bits(f2u(pc.z), p);
        q[0] = p.t0c; q[1] = p.waitc; q[2] = p.crossc; q[3] = p.tstop; q[4] = p.lpos; q[5] = p.dir;
        q[6] = p.gender; q[7] = p.age; q[8] = (int32_t)p.fl;
    }
}

MH_HD void import_one(const EnvArena &a, const EnvConst &c, int64_t n, const DumpPtrs &d) {
    uint32_t clo, chi;
    double_to_words(d.env_f[n], clo, chi);
    a.env_e[n] = make_float4(u2f(clo), u2f(chi),
                             u2f(pack_env_word((int)d.env_i[n * 4 + 0], (int)d.env_i[n * 4 + 1], (int)d.env_i[n * 4 + 2])),
                             u2f((uint32_t)d.env_i[n * 4 + 3]));
    for (int i = 0; i < c.nC; ++i) {
        const float *f = d.car_f + (n * c.nC + i) * 7; const int32_t *q = d.car_i + (n * c.nC + i) * 2;
        a.car_a[(int64_t)i * a.N + n] = make_float4(f[1], f[2], f[3], f[0]);
        a.car_b[(int64_t)i * a.N + n] = make_float4(f[4], f[5], f[6], u2f(((uint32_t)q[0] & 255u) | (((uint32_t)q[1] & 1u) << 8)));
    }
    for (int j = 0; j < c.nP; ++j) {
        const float *f = d.ped_f + (n * c.nP + j) * 9; const int32_t *q = d.ped_i + (n * c.nP + j) * 9;
        PedR p; p.t0c = q[0]; p.waitc = q[1]; p.crossc = q[2]; p.tstop = q[3]; p.lpos = q[4]; p.dir = q[5];
        p.gender = q[6]; p.age = q[7]; p.fl = (uint32_t)q[8];
        a.ped_a[(int64_t)j * a.N + n] = make_float4(f[2], f[3], f[0], f[1]);
        a.ped_b[(int64_t)j * a.N + n] = make_float4(f[4], f[5], f[6], f[7]);
        const Geo g = make_geo(d.env_f[n], c.L);
        a.ped_c[(int64_t)j * a.N + n] = make_float4(f[8], u2f(pack_ped_counts(p)), u2f(pack_ped_bits(p) | sym_tag(g, p, f[3])), 0.f);
    }
}

}  // namespace mhppo
