// TEST INFRASTRUCTURE ONLY -- CPU oracle for deepterrainrl_b200. Nothing under oracle/ is linked, imported or
// executed by the product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
//
// Reader for the `.trlpack` scene files written by tools/pack_scene.py (format documented there).
// Independent of the product's reader (deepterrainrl_b200/csrc/host/scene_pack.h) on purpose.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

struct PackArray {
    int dtype = 0;  // 0 f64, 1 i32
    std::vector<double> f;
    std::vector<int32_t> i;
};

struct Pack {
    std::map<std::string, PackArray> rec;

    bool has(const std::string& k) const { return rec.count(k) != 0; }
    const std::vector<double>& f64(const std::string& k) const {
        auto it = rec.find(k);
        if (it == rec.end() || it->second.dtype != 0) throw std::runtime_error("pack: missing f64 record " + k);
        return it->second.f;
    }
    const std::vector<int32_t>& i32(const std::string& k) const {
        auto it = rec.find(k);
        if (it == rec.end() || it->second.dtype != 1) throw std::runtime_error("pack: missing i32 record " + k);
        return it->second.i;
    }

    static Pack load(const std::string& path) {
        FILE* fp = std::fopen(path.c_str(), "rb");
        if (!fp) throw std::runtime_error("pack: cannot open " + path);
        auto rd = [&](void* p, size_t n) {
            if (std::fread(p, 1, n, fp) != n) { std::fclose(fp); throw std::runtime_error("pack: truncated " + path); }
        };
        char magic[8];
        rd(magic, 8);
        if (std::memcmp(magic, "TRLPACK1", 8) != 0) { std::fclose(fp); throw std::runtime_error("pack: bad magic"); }
        uint32_t n;
        rd(&n, 4);
        Pack p;
        for (uint32_t r = 0; r < n; ++r) {
            uint32_t ln, dt;
            uint64_t cnt;
            rd(&ln, 4);
            std::string name(ln, '\0');
            rd(&name[0], ln);
            rd(&dt, 4);
            rd(&cnt, 8);
            PackArray a;
            a.dtype = (int)dt;
            if (dt == 1) { a.i.resize(cnt); if (cnt) rd(a.i.data(), 4 * cnt); }
            else { a.f.resize(cnt); if (cnt) rd(a.f.data(), 8 * cnt); }
            p.rec[name] = std::move(a);
        }
        std::fclose(fp);
        return p;
    }
};

}  // namespace orc
