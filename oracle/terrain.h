// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/README.md). Never linked into the product path.
//
// Procedural terrain: RNG, strip generators, two-segment streaming ground and height sampling.
// Uses the *actual* libstdc++ <random> types the reference uses, so the product's hand-written device
// restatement of minstd_rand0 / generate_canonical / uniform_int_distribution is checked against the real thing.
//
// Pinned bit for bit against the reference's own compiled cRand / cTerrainGen2D (oracle/_ref, tests/test_ref_pinning_cpu.py).
//
// Follows (file:line under /root/reference):
//   util/Rand.cpp:6-104             cRand (default_random_engine + std distributions)
//   sim/TerrainGen2D.cpp:4,185-707  gVertSpacing (a float!), Build*, Add*, Overlay*
//   sim/GroundVar2D.cpp:43-116      streaming Update, segment pick for SampleHeight
//   sim/GroundVar2D.cpp:239-355     InitSegments / BuildSegment / AddPadding
//   sim/GroundVar2D.cpp:559-632     per-segment SampleHeight / CalcGridCoord (Bullet float round-trips of the
//                                   segment origin are not reproducible without Bullet; the arithmetic below
//                                   is the same lerp on the same float vertex data in double)
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <random>
#include <vector>

namespace orc {

struct Rand {
    std::default_random_engine gen;
    std::uniform_real_distribution<double> dd{0, 1};
    std::normal_distribution<double> nd{0, 1};
    std::uniform_int_distribution<int> id{0, std::numeric_limits<int>::max()};
    void seed(unsigned long s) { gen.seed(s); }
    double rand_double() { return dd(gen); }
    double rand_double(double mn, double mx) {
        if (mn == mx) return mn;
        double r = dd(gen);
        return mn + (r * (mx - mn));
    }
    int rand_int() { return id(gen); }
    int rand_int(int mn, int mx) {
        if (mn == mx) return mn;
        int delta = mx - mn;
        int r = id(gen);
        return mn + r % delta;
    }
    bool flip_coin(double p = 0.5) { return rand_double(0, 1) < p; }
    int rand_sign() { return flip_coin() ? -1 : 1; }
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
    pSlopeDeltaRange, pSlopeDeltaMin, pSlopeDeltaMax, pTerrainParamMax
};

struct TerrainGen {
    static constexpr float kVertSpacing = 0.1f;  // sim/TerrainGen2D.cpp:4
    typedef std::vector<float> Data;

    static int calc_num_verts(double w) { return static_cast<int>(std::ceil(w / kVertSpacing)) + 1; }

    static double add_flat(double width, Data& d) {
        int nv = calc_num_verts(width);
        int sz0 = (int)d.size();
        bool empty = sz0 == 0;
        float base = 0;
        if (!empty) { --nv; base = d[sz0 - 1]; }
        for (int i = 0; i < nv; ++i) d.push_back(base);
        int added = (int)d.size() - sz0;
        if (empty) --added;
        return added * kVertSpacing;
    }
    static double add_box(double spacing, double width, double depth, Data& d) {
        int nv = calc_num_verts(spacing);
        int sz0 = (int)d.size();
        bool empty = sz0 == 0;
        float base = 0;
        if (!empty) { --nv; base = d[sz0 - 1]; }
        for (int i = 0; i < nv; ++i) d.push_back(base);
        nv = calc_num_verts(width) - 1;
        float gap_h = static_cast<float>(base + depth);
        for (int i = 0; i < nv; ++i) d.push_back(gap_h);
        d.push_back(base);
        int added = (int)d.size() - sz0;
        if (empty) --added;
        return added * kVertSpacing;
    }
    static double add_step(double width, double height, Data& d) {
        int nv = calc_num_verts(width);
        int sz0 = (int)d.size();
        bool empty = sz0 == 0;
        float base = 0;
        if (!empty) { --nv; base = d[sz0 - 1]; }
        for (int i = 0; i < nv; ++i) d.push_back(base);
        d.push_back(static_cast<float>(base + height));
        int added = (int)d.size() - sz0;
        if (empty) --added;
        return added * kVertSpacing;
    }
    static void overlay_slopes(double range, double dmin, double dmax, double init_slope, int beg, int end, Rand& r, Data& d) {
        double slope = init_slope, dh = 0;
        double mean = 0.5 * (dmin + dmax), diff = 0.5 * (dmax - dmin);
        for (int i = beg; i < end; ++i) {
            double delta = r.rand_double(0, range);
            double sign_rand = r.rand_double(-1, 1);
            double thr = (slope - mean) / diff;
            bool neg = sign_rand < thr;
            delta = neg ? -delta : delta;
            slope += delta;
            dh += slope * kVertSpacing;
            d[i] += static_cast<float>(dh);
        }
    }
    static void overlay_bumps(double mn, double mx, int beg, int end, Rand& r, Data& d) {
        for (int i = beg; i < end - 1; ++i) {
            double delta = r.rand_sign() * r.rand_double(mn, mx);
            d[i] += static_cast<float>(delta);
        }
    }

    static double build_flat(double w, const double*, Rand&, Data& d) { return add_flat(w, d); }
    static double build_gaps(double width, const double* p, Rand& r, Data& d) {
        double total = 0;
        while (total < width) {
            double spacing = r.rand_double(p[pGapSpacingMin], p[pGapSpacingMax]);
            double w = r.rand_double(p[pGapWMin], p[pGapWMax]);
            double dd = r.rand_double(p[pGapDMin], p[pGapDMax]);
            total += add_box(spacing, w, dd, d);
        }
        return total;
    }
    static void pick_h(double h0mn, double h0mx, double h1mn, double h1mx, Rand& r, double& mn, double& mx) {
        bool v0 = (h0mn != 0 || h0mx != 0), v1 = (h1mn != 0 || h1mx != 0);
        if (v0 && v1) { bool heads = r.flip_coin(); mn = heads ? h0mn : h1mn; mx = heads ? h0mx : h1mx; }
        else if (v0) { mn = h0mn; mx = h0mx; }
        else { mn = h1mn; mx = h1mx; }
    }
    static double build_steps(double width, const double* p, Rand& r, Data& d) {
        double total = 0;
        while (total < width) {
            double mn = 0, mx = 0;
            pick_h(p[pStepH0Min], p[pStepH0Max], p[pStepH1Min], p[pStepH1Max], r, mn, mx);
            double w = r.rand_double(p[pStepSpacingMin], p[pStepSpacingMax]);
            double h = r.rand_double(mn, mx);
            total += add_step(w, h, d);
        }
        return total;
    }
    static double build_walls(double width, const double* p, Rand& r, Data& d) {
        double total = 0;
        while (total < width) {
            double spacing = r.rand_double(p[pWallSpacingMin], p[pWallSpacingMax]);
            double w = r.rand_double(p[pWallWMin], p[pWallWMax]);
            double h = r.rand_double(p[pWallHMin], p[pWallHMax]);
            total += add_box(spacing, w, h, d);
        }
        return total;
    }
    static double build_mixed(double width, const double* p, Rand& r, Data& d) {
        double total = 0;
        const double dummy_w = kVertSpacing;
        while (total < width) {
            double cw = 0;
            int t = r.rand_int(0, 3);
            if (t == 0) cw = build_gaps(dummy_w, p, r, d);
            else if (t == 1) cw = build_steps(dummy_w, p, r, d);
            else if (t == 2) cw = build_walls(dummy_w, p, r, d);
            total += cw;
        }
        return total;
    }
    static double build_narrow_gaps(double width, const double* p, Rand& r, Data& d) {
        int cmin = std::max(1, (int)p[pNGapCountMin]), cmax = std::max(1, (int)p[pNGapCountMax]);
        double total = 0;
        while (total < width) {
            double spacing = r.rand_double(p[pNGapSpacingMin], p[pNGapSpacingMax]);
            int count = r.rand_int(cmin, cmax + 1);
            for (int i = 0; i < count; ++i) {
                double w = r.rand_double(p[pNGapWMin], p[pNGapWMax]);
                double dd = r.rand_double(p[pNGapDMin], p[pNGapDMax]);
                total += add_box(spacing, w, dd, d);
                spacing = r.rand_double(p[pNGapDistMin], p[pNGapDistMax]);
            }
        }
        return total;
    }
    static double build_cliffs(double width, const double* p, Rand& r, Data& d) {
        int mini_max = (int)p[pCliffMiniCountMax];
        int beg = (int)d.size();
        double total = 0;
        while (total < width) {
            double mn = 0, mx = 0;
            pick_h(p[pCliffH0Min], p[pCliffH0Max], p[pCliffH1Min], p[pCliffH1Max], r, mn, mx);
            double w = r.rand_double(p[pCliffSpacingMin], p[pCliffSpacingMax]);
            double h = r.rand_double(mn, mx);
            double cw = 0, cur_dh = 0;
            int num_mini = r.rand_int(0, mini_max + 1);
            for (int i = 0; i < num_mini + 1; ++i) {
                const double mini_w = (i == 0) ? w : 0.1;
                double mini_h = r.rand_double(cur_dh, h);
                mini_h = (i == num_mini) ? h : mini_h;
                double dh = mini_h - cur_dh;
                cw += add_step(mini_w, dh, d);
                cur_dh = mini_h;
            }
            total += cw;
        }
        int end = (int)d.size();
        overlay_slopes(std::abs(p[pSlopeDeltaRange]), p[pSlopeDeltaMin], p[pSlopeDeltaMax], 0, beg, end, r, d);
        overlay_bumps(p[pBumpHMin], p[pBumpHMax], beg, end, r, d);
        return total;
    }

    // cTerrainGen2D::GetTerrainFunc + the Build* wrappers (sim/TerrainGen2D.cpp:146-446)
    static double build(int type, double width, const double* p, Rand& r, Data& d) {
        int beg = (int)d.size();
        double total = 0;
        bool slopes = false;
        switch (type) {
            case kGaps: total = build_gaps(width, p, r, d); break;
            case kSteps: total = build_steps(width, p, r, d); break;
            case kWalls: total = build_walls(width, p, r, d); break;
            case kBumps:
                total = build_flat(width, p, r, d);
                overlay_bumps(p[pBumpHMin], p[pBumpHMax], beg, (int)d.size(), r, d);
                break;
            case kMixed: total = build_mixed(width, p, r, d); break;
            case kNarrowGaps: total = build_narrow_gaps(width, p, r, d); break;
            case kSlopes: total = build_flat(width, p, r, d); slopes = true; break;
            case kSlopesGaps: total = build_gaps(width, p, r, d); slopes = true; break;
            case kSlopesSteps: total = build_steps(width, p, r, d); slopes = true; break;
            case kSlopesWalls: total = build_walls(width, p, r, d); slopes = true; break;
            case kSlopesMixed: total = build_mixed(width, p, r, d); slopes = true; break;
            case kSlopesNarrowGaps: total = build_narrow_gaps(width, p, r, d); slopes = true; break;
            case kCliffs: total = build_cliffs(width, p, r, d); break;
            default: total = build_flat(width, p, r, d); break;
        }
        if (slopes)
            overlay_slopes(std::abs(p[pSlopeDeltaRange]), p[pSlopeDeltaMin], p[pSlopeDeltaMax], 0, beg, (int)d.size(), r, d);
        return total;
    }
};

// cGroundVar2D with its two ping-pong segments
struct Ground {
    static constexpr int kNumSeg = 2;
    struct Seg {
        std::vector<float> data;
        double min_x = 0;
        bool empty() const { return data.empty(); }
        double spacing() const { return (double)TerrainGen::kVertSpacing; }
        double get_min_x() const { return empty() ? std::numeric_limits<double>::infinity() : min_x; }
        double get_max_x() const { return empty() ? -std::numeric_limits<double>::infinity() : min_x + ((double)data.size() - 1) * spacing(); }
        double start_h() const { return data.front(); }
        double end_h() const { return data.back(); }
        // tSegment::SampleHeight / CalcGridCoord / ClampCoord (sim/GroundVar2D.cpp:559-632)
        double sample(double x, double* slope = nullptr) const {
            int w = (int)data.size();
            double coord = (x - min_x) / spacing();
            coord = std::min(std::max(coord, 0.0), w - 1.0);
            int i = (int)coord;
            int j = std::min(w - 1, i + 1);
            double lerp = coord - i;
            double a = data[i], b = data[j];
            if (slope) *slope = (b - a) / spacing();
            return (1 - lerp) * a + lerp * b;
        }
    };
    Seg seg[kNumSeg];
    bool flip = false;
    Rand rand;
    int type = kFlat;
    double params[pTerrainParamMax];
    double seg_width = 20;  // 2 * char_view_dist (scenarios/ScenarioSimChar.cpp:350-352)

    int seg_id(int s) const { return flip ? (s == 0 ? 1 : 0) : s; }
    const Seg& min_seg() const { return seg[seg_id(0)]; }
    const Seg& max_seg() const { return seg[seg_id(1)]; }
    double min_x() const { return min_seg().get_min_x(); }
    double max_x() const { return max_seg().get_max_x(); }

    void clear() { for (auto& s : seg) s.data.clear(); flip = false; }

    // cGroundVar2D::BuildSegment + AddPadding (sim/GroundVar2D.cpp:312-355)
    void build_segment(int id, double bmin, double bmax, bool align_min, double fix_y) {
        Seg& s = seg[id];
        s.data.clear();
        bool contains_origin = (bmin <= 0) && (bmax >= 0);
        if (contains_origin) {
            double flat_w = std::min(bmax - bmin, 1 - bmin);
            TerrainGen::build_flat(flat_w, params, rand, s.data);
        }
        TerrainGen::build(type, bmax - bmin, params, rand, s.data);
        int nv = (int)s.data.size();
        float end_h = 0;
        if (nv > 0) end_h = align_min ? s.data[0] : s.data[nv - 1];
        float h_off = static_cast<float>(fix_y - end_h);
        for (int i = 0; i < nv; ++i) s.data[i] += h_off;
        s.min_x = align_min ? bmin : (bmax - (nv - 1) * (double)TerrainGen::kVertSpacing);
    }
    // cGroundVar2D::InitSegments (sim/GroundVar2D.cpp:239-259)
    void init_segments(double bmin, double bmax) {
        clear();
        double mid = 0.5 * (bmax + bmin);
        for (int i = 0; i < kNumSeg; ++i) {
            int id = seg_id(i);
            bool align_max = (i == 0);  // GetSegAlignMode with mFlipSeg == false
            double w = seg_width;
            double lo = (align_max ? -w : 0) + mid, hi = (align_max ? 0 : w) + mid;
            build_segment(id, lo, hi, !align_max, 0.0);
        }
    }
    // cGroundVar2D::Update (sim/GroundVar2D.cpp:43-91)
    void update(double bmin, double bmax) {
        double mn = min_x(), mx = max_x();
        if (bmax < mx && bmin > mn) return;
        if (bmax <= mn || bmin >= mx) { init_segments(bmin, bmax); return; }
        if (bmax >= mx) {
            int id = seg_id(0);
            build_segment(id, mx, mx + seg_width, true, max_seg().end_h());
        } else {
            int id = seg_id(1);
            build_segment(id, mn - seg_width, mn, false, min_seg().start_h());
        }
        flip = (seg_id(0) == 0);
    }
    // (this project's contact model, not a restatement) the terrain vertices with xmin <= x <= xmax, each once: the vertex the two
    // segments share at their seam is reported from the max segment, as sample() resolves it.  fn(x, h, h_prev, h_next); the
    // neighbours of a window's end vertices repeat the end height (sample()'s clamp).
    template <typename F>
    void for_vertices(double xmin, double xmax, F&& fn) const {
        const double sp = (double)TerrainGen::kVertSpacing;
        const double seam = min_seg().get_max_x();
        for (int si = 0; si < kNumSeg; ++si) {
            const Seg& sg = seg[seg_id(si)];
            const int w = (int)sg.data.size();
            if (w == 0) continue;
            int k0 = (int)std::ceil((xmin - sg.min_x) / sp - 1e-9), k1 = (int)std::floor((xmax - sg.min_x) / sp + 1e-9);
            k0 = std::max(k0, 0); k1 = std::min(k1, w - 1);
            for (int k = k0; k <= k1; ++k) {
                const double x = sg.min_x + k * sp;
                if (si == 0 && x >= seam) continue;
                if (si == 1 && x < seam) continue;
                const double hp = k > 0 ? (double)sg.data[k - 1] : (si == 1 ? sample(x - sp) : (double)sg.data[k]);
                const double hn = k < w - 1 ? (double)sg.data[k + 1] : (si == 0 ? sample(x + sp) : (double)sg.data[k]);
                fn(x, (double)sg.data[k], hp, hn);
            }
        }
    }
    // cGroundVar2D::SampleHeight (sim/GroundVar2D.cpp:103-116)
    double sample(double x, double* slope = nullptr) const {
        const Seg& ms = min_seg();
        int idx = (x >= ms.get_max_x()) ? 1 : 0;
        return seg[seg_id(idx)].sample(x, slope);
    }
};

}  // namespace orc
