// deepterrainrl_b200 -- device-side procedural terrain (one env per lane; runs only on the rare lanes that cross a
// segment boundary or reset, so divergence cost is amortised to ~1e-3 of a step).
//
// Reproduces, bit for bit, what the reference computes with libstdc++'s <random> on the host:
//   util/Rand.cpp:6-104            cRand = std::default_random_engine (= minstd_rand0) + uniform_real_distribution<double>
//                                  + uniform_int_distribution<int>(0, INT_MAX)
//   sim/TerrainGen2D.cpp:4,185-707 strip generators (float vertices, double accumulators, float 0.1f spacing)
//   sim/GroundVar2D.cpp:43-116,239-355,559-632   two ping-pong segments, streaming update, lerp sampling
// libstdc++ (GCC 13) algorithms restated here: generate_canonical<double,53> over minstd_rand0 (2 draws, range
// R = 2147483646: (x1-1 + (x2-1) R) / R^2) and uniform_int_distribution's "upscaling" path with its nested
// two-division rejection loop.  Every float/double operation that the host performs as separate IEEE operations is
// written with an explicit *_rn intrinsic so nvcc cannot contract it into an FMA.
#pragma once
#include "trl_types.h"

#ifndef TRL_NOINLINE_COLD
#define TRL_NOINLINE_COLD 0
#endif
#if TRL_NOINLINE_COLD && !defined(TRL_SIMT_EMU)
#define TRL_COLD_ATTR __noinline__
#else
#define TRL_COLD_ATTR
#endif

namespace trl {

struct TerrainRng {
    uint32_t x;
    __device__ __forceinline__ uint32_t next() {
        // x * 16807 mod (2^31 - 1) by Mersenne folding (exactly the minstd_rand0 recurrence)
        uint64_t p = (uint64_t)x * 16807ull;
        uint32_t r = (uint32_t)(p & 0x7fffffffull) + (uint32_t)(p >> 31);
        if (r >= 2147483647u) r -= 2147483647u;
        x = r;
        return x;
    }
    __device__ __forceinline__ static uint32_t seed_state(uint64_t s) {
        uint32_t v = (uint32_t)(s % 2147483647ull);
        return v == 0 ? 1u : v;
    }
    // std::generate_canonical<double, 53>(minstd_rand0)
    __device__ double canonical() {
        const double R = 2147483646.0;
        double x1 = (double)(next() - 1u);
        double x2 = (double)(next() - 1u);
        double sum = __dadd_rn(x1, __dmul_rn(x2, R));
        double ret = __ddiv_rn(sum, __dmul_rn(R, R));
        if (ret >= 1.0) ret = 0.99999999999999988897769753748;  // nextafter(1, 0)
        return ret;
    }
    __device__ double rand_double(double mn, double mx) {
        if (mn == mx) return mn;
        double r = canonical();
        return __dadd_rn(mn, __dmul_rn(r, __dsub_rn(mx, mn)));
    }
    // std::uniform_int_distribution<int>(0, INT_MAX)(minstd_rand0), libstdc++ bits/uniform_int_dist.h
    __device__ int rand_int_full() {
        const uint64_t urngrange = 2147483645ull, urange = 2147483647ull;
        uint64_t ret, tmp;
        do {
            // nested call for the range [0, urange / (urngrange + 1)] = [0, 1]: downscaling with two divisions
            const uint64_t scaling = urngrange / 2ull, past = 2ull * scaling;
            uint64_t inner;
            do { inner = (uint64_t)next() - 1ull; } while (inner >= past);
            inner /= scaling;
            tmp = (urngrange + 1ull) * inner;
            ret = tmp + ((uint64_t)next() - 1ull);
        } while (ret > urange || ret < tmp);
        return (int)ret;
    }
    __device__ int rand_int(int mn, int mx) {
        if (mn == mx) return mn;
        int delta = mx - mn;
        int r = rand_int_full();
        return mn + r % delta;
    }
    __device__ bool flip_coin() { return rand_double(0.0, 1.0) < 0.5; }
    __device__ int rand_sign() { return flip_coin() ? -1 : 1; }
};

enum TerrainType {
    kFlat, kGaps, kSteps, kWalls, kBumps, kMixed, kNarrowGaps, kSlopes, kSlopesGaps, kSlopesWalls, kSlopesSteps,
    kSlopesMixed, kSlopesNarrowGaps, kCliffs
};
enum TerrainParam {
    pGapSpacingMin, pGapSpacingMax, pGapWMin, pGapWMax, pGapDMin, pGapDMax,
    pWallSpacingMin, pWallSpacingMax, pWallWMin, pWallWMax, pWallHMin, pWallHMax,
    pStepSpacingMin, pStepSpacingMax, pStepH0Min, pStepH0Max, pStepH1Min, pStepH1Max,
    pBumpHMin, pBumpHMax,
    pNGapSpacingMin, pNGapSpacingMax, pNGapDistMin, pNGapDistMax, pNGapWMin, pNGapWMax, pNGapDMin, pNGapDMax,
    pNGapCountMin, pNGapCountMax,
    pCliffSpacingMin, pCliffSpacingMax, pCliffH0Min, pCliffH0Max, pCliffH1Min, pCliffH1Max, pCliffMiniCountMax,
    pSlopeDeltaRange, pSlopeDeltaMin, pSlopeDeltaMax
};

#define TRL_VERT_SPACING_F 0.1f
#define TRL_VERT_SPACING_D ((double)0.1f)

// Strip builder writing into one segment's float array (std::vector<float> push_back semantics with a hard cap).
struct StripBuilder {
    float* d;
    int n;
    TerrainRng* rng;
    const double* p;

    __device__ __forceinline__ void push(float v) {
        if (n < kTerrainCap) d[n] = v;
        ++n;
    }
    __device__ __forceinline__ float back() const { return d[(n < kTerrainCap ? n : kTerrainCap) - 1]; }
    __device__ static int calc_num_verts(double w) { return (int)ceil(__ddiv_rn(w, TRL_VERT_SPACING_D)) + 1; }
    __device__ static double width_of(int verts) { return (double)__fmul_rn((float)verts, TRL_VERT_SPACING_F); }

    __device__ double add_flat(double width) {
        int nv = calc_num_verts(width);
        int n0 = n;
        bool empty = n0 == 0;
        float base = 0.f;
        if (!empty) { --nv; base = back(); }
        for (int i = 0; i < nv; ++i) push(base);
        int added = n - n0;
        if (empty) --added;
        return width_of(added);
    }
    __device__ double add_box(double spacing, double width, double depth) {
        int nv = calc_num_verts(spacing);
        int n0 = n;
        bool empty = n0 == 0;
        float base = 0.f;
        if (!empty) { --nv; base = back(); }
        for (int i = 0; i < nv; ++i) push(base);
        nv = calc_num_verts(width) - 1;
        float gap_h = __double2float_rn(__dadd_rn((double)base, depth));
        for (int i = 0; i < nv; ++i) push(gap_h);
        push(base);
        int added = n - n0;
        if (empty) --added;
        return width_of(added);
    }
    __device__ double add_step(double width, double height) {
        int nv = calc_num_verts(width);
        int n0 = n;
        bool empty = n0 == 0;
        float base = 0.f;
        if (!empty) { --nv; base = back(); }
        for (int i = 0; i < nv; ++i) push(base);
        push(__double2float_rn(__dadd_rn((double)base, height)));
        int added = n - n0;
        if (empty) --added;
        return width_of(added);
    }
    __device__ void overlay_slopes(int beg, int end) {
        double range = fabs(p[pSlopeDeltaRange]), dmin = p[pSlopeDeltaMin], dmax = p[pSlopeDeltaMax];
        double slope = 0.0, dh = 0.0;
        double mean = __dmul_rn(0.5, __dadd_rn(dmin, dmax)), diff = __dmul_rn(0.5, __dsub_rn(dmax, dmin));
        for (int i = beg; i < end; ++i) {
            double delta = rng->rand_double(0.0, range);
            double sign_rand = rng->rand_double(-1.0, 1.0);
            double thr = __ddiv_rn(__dsub_rn(slope, mean), diff);
            if (sign_rand < thr) delta = -delta;
            slope = __dadd_rn(slope, delta);
            dh = __dadd_rn(dh, __dmul_rn(slope, TRL_VERT_SPACING_D));
            if (i < kTerrainCap) d[i] = __fadd_rn(d[i], __double2float_rn(dh));
        }
    }
    __device__ void overlay_bumps(int beg, int end) {
        for (int i = beg; i < end - 1; ++i) {
            int sgn = rng->rand_sign();
            double delta = __dmul_rn((double)sgn, rng->rand_double(p[pBumpHMin], p[pBumpHMax]));
            if (i < kTerrainCap) d[i] = __fadd_rn(d[i], __double2float_rn(delta));
        }
    }
    __device__ double build_gaps(double width) {
        double total = 0.0;
        while (total < width) {
            double spacing = rng->rand_double(p[pGapSpacingMin], p[pGapSpacingMax]);
            double w = rng->rand_double(p[pGapWMin], p[pGapWMax]);
            double dd = rng->rand_double(p[pGapDMin], p[pGapDMax]);
            total = __dadd_rn(total, add_box(spacing, w, dd));
        }
        return total;
    }
    __device__ void pick_h(double h0mn, double h0mx, double h1mn, double h1mx, double& mn, double& mx) {
        bool v0 = (h0mn != 0 || h0mx != 0), v1 = (h1mn != 0 || h1mx != 0);
        if (v0 && v1) { bool heads = rng->flip_coin(); mn = heads ? h0mn : h1mn; mx = heads ? h0mx : h1mx; }
        else if (v0) { mn = h0mn; mx = h0mx; }
        else { mn = h1mn; mx = h1mx; }
    }
    __device__ double build_steps(double width) {
        double total = 0.0;
        while (total < width) {
            double mn = 0, mx = 0;
            pick_h(p[pStepH0Min], p[pStepH0Max], p[pStepH1Min], p[pStepH1Max], mn, mx);
            double w = rng->rand_double(p[pStepSpacingMin], p[pStepSpacingMax]);
            double h = rng->rand_double(mn, mx);
            total = __dadd_rn(total, add_step(w, h));
        }
        return total;
    }
    __device__ double build_walls(double width) {
        double total = 0.0;
        while (total < width) {
            double spacing = rng->rand_double(p[pWallSpacingMin], p[pWallSpacingMax]);
            double w = rng->rand_double(p[pWallWMin], p[pWallWMax]);
            double h = rng->rand_double(p[pWallHMin], p[pWallHMax]);
            total = __dadd_rn(total, add_box(spacing, w, h));
        }
        return total;
    }
    __device__ double build_mixed(double width) {
        double total = 0.0;
        const double dummy_w = TRL_VERT_SPACING_D;
        while (total < width) {
            double cw = 0.0;
            int t = rng->rand_int(0, 3);
            if (t == 0) cw = build_gaps(dummy_w);
            else if (t == 1) cw = build_steps(dummy_w);
            else if (t == 2) cw = build_walls(dummy_w);
            total = __dadd_rn(total, cw);
        }
        return total;
    }
    __device__ double build_narrow_gaps(double width) {
        int cmin = max(1, (int)p[pNGapCountMin]), cmax = max(1, (int)p[pNGapCountMax]);
        double total = 0.0;
        while (total < width) {
            double spacing = rng->rand_double(p[pNGapSpacingMin], p[pNGapSpacingMax]);
            int count = rng->rand_int(cmin, cmax + 1);
            for (int i = 0; i < count; ++i) {
                double w = rng->rand_double(p[pNGapWMin], p[pNGapWMax]);
                double dd = rng->rand_double(p[pNGapDMin], p[pNGapDMax]);
                total = __dadd_rn(total, add_box(spacing, w, dd));
                spacing = rng->rand_double(p[pNGapDistMin], p[pNGapDistMax]);
            }
        }
        return total;
    }
    __device__ double build_cliffs(double width) {
        int mini_max = (int)p[pCliffMiniCountMax];
        int beg = n;
        double total = 0.0;
        while (total < width) {
            double mn = 0, mx = 0;
            pick_h(p[pCliffH0Min], p[pCliffH0Max], p[pCliffH1Min], p[pCliffH1Max], mn, mx);
            double w = rng->rand_double(p[pCliffSpacingMin], p[pCliffSpacingMax]);
            double h = rng->rand_double(mn, mx);
            double cw = 0.0, cur_dh = 0.0;
            int num_mini = rng->rand_int(0, mini_max + 1);
            for (int i = 0; i < num_mini + 1; ++i) {
                const double mini_w = (i == 0) ? w : 0.1;
                double mini_h = rng->rand_double(cur_dh, h);
                mini_h = (i == num_mini) ? h : mini_h;
                double dh = __dsub_rn(mini_h, cur_dh);
                cw = __dadd_rn(cw, add_step(mini_w, dh));
                cur_dh = mini_h;
            }
            total = __dadd_rn(total, cw);
        }
        int end = n;
        overlay_slopes(beg, end);
        overlay_bumps(beg, end);
        return total;
    }
    __device__ double build(int type, double width) {
        int beg = n;
        double total = 0.0;
        bool slopes = false;
        switch (type) {
            case kGaps: total = build_gaps(width); break;
            case kSteps: total = build_steps(width); break;
            case kWalls: total = build_walls(width); break;
            case kBumps: total = add_flat(width); overlay_bumps(beg, n); break;
            case kMixed: total = build_mixed(width); break;
            case kNarrowGaps: total = build_narrow_gaps(width); break;
            case kSlopes: total = add_flat(width); slopes = true; break;
            case kSlopesGaps: total = build_gaps(width); slopes = true; break;
            case kSlopesSteps: total = build_steps(width); slopes = true; break;
            case kSlopesWalls: total = build_walls(width); slopes = true; break;
            case kSlopesMixed: total = build_mixed(width); slopes = true; break;
            case kSlopesNarrowGaps: total = build_narrow_gaps(width); slopes = true; break;
            case kCliffs: total = build_cliffs(width); break;
            default: total = add_flat(width); break;
        }
        if (slopes) overlay_slopes(beg, n);
        return total;
    }
};

// Per-env view of the two-segment ground (cGroundVar2D).
struct GroundView {
    float* data;      // [2][kTerrainCap]
    int n[2];
    double min_x[2];
    int flip;
    uint32_t rng_state;

    // n[] / min_x[] are only ever indexed with constants (through these selects): a run-time index would force the whole view into
    // local memory, and every height sample of the contact pass would pay two local-memory loads for it
    __device__ __forceinline__ int nn(int id) const { return id ? n[1] : n[0]; }
    __device__ __forceinline__ double mnx(int id) const { return id ? min_x[1] : min_x[0]; }

    __device__ __forceinline__ int seg_id(int s) const { return flip ? (s == 0 ? 1 : 0) : s; }
    __device__ __forceinline__ double seg_max_x(int id) const {
        return nn(id) == 0 ? -INFINITY : __dadd_rn(mnx(id), __dmul_rn((double)(nn(id) - 1), TRL_VERT_SPACING_D));
    }
    __device__ __forceinline__ double seg_min_x(int id) const { return nn(id) == 0 ? INFINITY : mnx(id); }

    // tSegment::SampleHeight (clamped grid coordinate, lerp of the two neighbouring float vertices)
    __device__ double sample_seg(int id, double x, double* slope) const {
        const float* d = data + id * kTerrainCap;
        int w = nn(id) < kTerrainCap ? nn(id) : kTerrainCap;
        double coord = (x - mnx(id)) / TRL_VERT_SPACING_D;
        coord = fmin(fmax(coord, 0.0), (double)(w - 1));
        int i = (int)coord;
        int j = min(w - 1, i + 1);
        double lerp = coord - (double)i;
        double a = (double)d[i], b = (double)d[j];
        if (slope) *slope = (b - a) / TRL_VERT_SPACING_D;
        return (1.0 - lerp) * a + lerp * b;
    }
    __device__ double sample(double x, double* slope = nullptr) const {
        int ms = seg_id(0);
        int id = (x >= seg_max_x(ms)) ? seg_id(1) : ms;
        return sample_seg(id, x, slope);
    }

    // sample() for the contact pass: same vertices and lerp, with the two divisions by the vertex spacing replaced by
    // multiplications with its reciprocal (differs from sample() by an ulp or two; policy features and resets keep the
    // division so that they stay bit-identical to cGroundVar2D::SampleHeight)
    __device__ __forceinline__ double sample_fast(double x, double* slope) const {
        const double inv_sp = 1.0 / TRL_VERT_SPACING_D;
        const int ms = seg_id(0);
        const int id = (x >= seg_max_x(ms)) ? seg_id(1) : ms;
        const float* d = data + id * kTerrainCap;
        const int w = nn(id) < kTerrainCap ? nn(id) : kTerrainCap;
        double coord = (x - mnx(id)) * inv_sp;
        coord = fmin(fmax(coord, 0.0), (double)(w - 1));
        const int i = (int)coord;
        const int j = min(w - 1, i + 1);
        const double lerp = coord - (double)i;
        const double a = (double)d[i], b = (double)d[j];
        *slope = (b - a) * inv_sp;
        return a + lerp * (b - a);
    }

    // The slot-th terrain vertex with xmin <= x <= xmax (this project's contact model; oracle/terrain.h: Ground::for_vertices): the
    // min segment's vertices first (without its last one, which the max segment repeats at the seam), then the max segment's.
    // Returns the vertex (x, h) and its two neighbours' heights; the window's end vertices repeat their own height.
    __device__ bool vertex_slot(double xmin, double xmax, int slot, double* xv, double* hv, double* hp, double* hn) const {
        const double sp = TRL_VERT_SPACING_D;
        const int ia = seg_id(0), ib = seg_id(1);
        const int wa = nn(ia) < kTerrainCap ? nn(ia) : kTerrainCap, wb = nn(ib) < kTerrainCap ? nn(ib) : kTerrainCap;
        int a0 = 0, a1 = -1, b0 = 0, b1 = -1;
        if (wa > 0) {
            a0 = max((int)ceil((xmin - mnx(ia)) / sp - 1e-9), 0);
            a1 = min((int)floor((xmax - mnx(ia)) / sp + 1e-9), wa - 2);        // x < seam: the last vertex belongs to the max segment
        }
        if (wb > 0) {
            b0 = max((int)ceil((xmin - mnx(ib)) / sp - 1e-9), 0);
            b1 = min((int)floor((xmax - mnx(ib)) / sp + 1e-9), wb - 1);
        }
        const int na = max(a1 - a0 + 1, 0);
        if (slot < na) {
            const int kk = a0 + slot;
            const float* d = data + ia * kTerrainCap;
            *xv = mnx(ia) + kk * sp; *hv = (double)d[kk];
            *hp = kk > 0 ? (double)d[kk - 1] : (double)d[kk];
            *hn = (double)d[kk + 1];
            return true;
        }
        const int kk = b0 + (slot - na);
        if (kk > b1) return false;
        const float* d = data + ib * kTerrainCap;
        *xv = mnx(ib) + kk * sp; *hv = (double)d[kk];
        *hp = kk > 0 ? (double)d[kk - 1] : sample(*xv - sp);
        *hn = kk < wb - 1 ? (double)d[kk + 1] : (double)d[kk];
        return true;
    }

    // Upper bound of sample(x) over x in [x0, x1]: max of the vertices either segment can interpolate between for that
    // window (clamped sampling maps x outside a segment to its end vertex, which the clamped index range includes).
    // Warp-cooperative; every lane returns the same value.
    __device__ double window_max(double x0, double x1, int lane) const {
        float mx = -INFINITY;
#pragma unroll
        for (int id = 0; id < 2; ++id) {
            const int w = nn(id) < kTerrainCap ? nn(id) : kTerrainCap;
            if (w <= 0) continue;
            const float* d = data + id * kTerrainCap;
            const double c0 = fmin(fmax((x0 - mnx(id)) / TRL_VERT_SPACING_D, 0.0), (double)(w - 1));
            const double c1 = fmin(fmax((x1 - mnx(id)) / TRL_VERT_SPACING_D, 0.0), (double)(w - 1));
            const int i0 = (int)c0, i1 = min(w - 1, (int)c1 + 1);
            for (int i = i0 + lane; i <= i1; i += 32) mx = fmaxf(mx, d[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        return (double)mx;
    }

    // cGroundVar2D::BuildSegment + AddPadding
    __device__ void build_segment(int id, double bmin, double bmax, bool align_min, double fix_y, int type,
                                  const double* params, double seg_width) {
        TerrainRng rng{rng_state};
        StripBuilder sb{data + id * kTerrainCap, 0, &rng, params};
        bool contains_origin = (bmin <= 0.0) && (bmax >= 0.0);
        if (contains_origin) {
            double flat_w = fmin(__dsub_rn(bmax, bmin), __dsub_rn(1.0, bmin));
            sb.add_flat(flat_w);
        }
        sb.build(type, __dsub_rn(bmax, bmin));
        int nv = sb.n < kTerrainCap ? sb.n : kTerrainCap;
        float* d = data + id * kTerrainCap;
        float end_h = 0.f;
        if (nv > 0) end_h = align_min ? d[0] : d[nv - 1];
        float h_off = __double2float_rn(__dsub_rn(fix_y, (double)end_h));
        for (int i = 0; i < nv; ++i) d[i] = __fadd_rn(d[i], h_off);
        if (id) n[1] = nv; else n[0] = nv;
        const double mn_ = align_min ? bmin : __dsub_rn(bmax, __dmul_rn((double)(nv - 1), TRL_VERT_SPACING_D));
        if (id) min_x[1] = mn_; else min_x[0] = mn_;
        rng_state = rng.x;
    }
    __device__ void init_segments(double bmin, double bmax, int type, const double* params, double seg_width) {
        n[0] = n[1] = 0;
        flip = 0;
        double mid = __dmul_rn(0.5, __dadd_rn(bmax, bmin));
        for (int i = 0; i < 2; ++i) {
            bool align_max = (i == 0);
            double lo = __dadd_rn(align_max ? -seg_width : 0.0, mid), hi = __dadd_rn(align_max ? 0.0 : seg_width, mid);
            build_segment(i, lo, hi, !align_max, 0.0, type, params, seg_width);
        }
    }
    // cGroundVar2D::Update; returns true if anything was rebuilt.  -DTRL_NOINLINE_COLD=1 (experiment) keeps this rare, large
    // path (the generators are ~19 k instructions) out of line, so the env-step kernel's hot loop is not interleaved with it
    // in the instruction cache; same arithmetic, bit-identical results.
    __device__ TRL_COLD_ATTR bool update(double bmin, double bmax, int type, const double* params, double seg_width) {
        int smin = seg_id(0), smax = seg_id(1);
        double mn = seg_min_x(smin), mx = seg_max_x(smax);
        if (bmax < mx && bmin > mn) return false;
        if (bmax <= mn || bmin >= mx) { init_segments(bmin, bmax, type, params, seg_width); return true; }
        if (bmax >= mx) {
            const float* dm = data + smax * kTerrainCap;
            double end_h = (double)dm[nn(smax) - 1];
            build_segment(smin, mx, __dadd_rn(mx, seg_width), true, end_h, type, params, seg_width);
        } else {
            const float* dm = data + smin * kTerrainCap;
            double start_h = (double)dm[0];
            build_segment(smax, __dsub_rn(mn, seg_width), mn, false, start_h, type, params, seg_width);
        }
        flip = (seg_id(0) == 0) ? 1 : 0;
        return true;
    }
};

}  // namespace trl
