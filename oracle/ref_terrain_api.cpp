// TEST INFRASTRUCTURE ONLY -- C entry points around the REFERENCE's own sources, compiled where they lie under /root/reference
// into oracle/_ref/libref_terrain.so (recipe: oracle/Makefile, target ref).  Used by tests/test_ref_pinning_cpu.py to pin the
// oracle's restatement of cRand / cTerrainGen2D / cArgParser (oracle/terrain.h, csrc/ref_loader.h) bit for bit.
#include <cstring>
#include <string>
#include <vector>

#include "sim/TerrainGen2D.h"
#include "util/ArgParser.h"

extern "C" {

// cTerrainGen2D::GetTerrainFunc(type)(width, params, rand, out) with a cRand seeded like cGroundVar2D does
int ref_terrain_build(int type, const double* params40, unsigned long seed, double width, float* out, int cap, double* total_w) {
    cTerrainGen2D::tParams params;
    for (int i = 0; i < cTerrainGen2D::eParamsMax; ++i) params[i] = params40[i];
    cRand rand;
    rand.Seed(seed);
    std::vector<float> data;
    double w = cTerrainGen2D::GetTerrainFunc(static_cast<cTerrainGen2D::eType>(type))(width, params, rand, data);
    if (total_w) *total_w = w;
    int n = (int)data.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = data[i];
    return n;
}
// consecutive strips appended to the same vertex vector with the same generator (what BuildSegment does when the first strip
// is the flat start): exercises the non-empty branches of Add*
int ref_terrain_build_after_flat(int type, const double* params40, unsigned long seed, double flat_w, double width, float* out, int cap) {
    cTerrainGen2D::tParams params;
    for (int i = 0; i < cTerrainGen2D::eParamsMax; ++i) params[i] = params40[i];
    cRand rand;
    rand.Seed(seed);
    std::vector<float> data;
    cTerrainGen2D::AddFlat(flat_w, data);
    cTerrainGen2D::GetTerrainFunc(static_cast<cTerrainGen2D::eType>(type))(width, params, rand, data);
    int n = (int)data.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = data[i];
    return n;
}
void ref_terrain_default_params(double* out40) {
    cTerrainGen2D::tParams p = cTerrainGen2D::GetDefaultParams();
    for (int i = 0; i < cTerrainGen2D::eParamsMax; ++i) out40[i] = p[i];
}
int ref_terrain_parse_type(const char* name) {
    cTerrainGen2D::eType t;
    cTerrainGen2D::ParseType(name, t);
    return (int)t;
}
// cRand stream: n draws of the given kind after Seed(seed). kind 0 RandDouble(), 1 RandDouble(a,b), 2 RandInt(), 3 RandInt(a,b),
// 4 FlipCoin(), 5 RandSign()
void ref_rand_stream(unsigned long seed, int kind, double a, double b, int n, double* out) {
    cRand r;
    r.Seed(seed);
    for (int i = 0; i < n; ++i) {
        switch (kind) {
            case 0: out[i] = r.RandDouble(); break;
            case 1: out[i] = r.RandDouble(a, b); break;
            case 2: out[i] = r.RandInt(); break;
            case 3: out[i] = r.RandInt((int)a, (int)b); break;
            case 4: out[i] = r.FlipCoin() ? 1 : 0; break;
            default: out[i] = r.RandSign(); break;
        }
    }
}
// cArgParser(file) then CLI tokens are NOT mixed here: one parser over one arg file; returns 1 and the value if the key parses
int ref_args_string(const char* file, const char* key, char* out, int cap) {
    cArgParser p{std::string(file)};
    std::string v;
    if (!p.ParseString(key, v)) return 0;
    std::strncpy(out, v.c_str(), cap - 1);
    out[cap - 1] = 0;
    return 1;
}
int ref_args_double(const char* file, const char* key, double* out) {
    cArgParser p{std::string(file)};
    return p.ParseDouble(key, *out) ? 1 : 0;
}
int ref_args_count(const char* file) {
    cArgParser p{std::string(file)};
    return p.GetNumArgs();
}

}  // extern "C"
