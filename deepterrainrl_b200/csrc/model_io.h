// deepterrainrl_b200 -- Caffe `ToHDF5` model files written natively (cNeuralNet::OutputModel, learning/NeuralNet.cpp:571-587,
// 1182-1205): HDF5 superblock v0, old-style groups /data/<layer>/<blob>, contiguous IEEE f64 datasets, one (possibly empty) group
// per layer of the deploy net, plus `<stem>_scale.txt`.  Byte-for-byte the same output as deepterrainrl_b200/model_io.py
// (tests/test_train_host_cpu.py compares the two), whose object layout follows the files the reference ships.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace trl {

class H5Writer {
public:
    struct Blob { std::vector<uint64_t> dims; const double* data; };

    // layers: group names in net order; blobs: layer -> datasets "0", "1", ...
    static std::vector<uint8_t> build(const std::vector<std::string>& layers, const std::map<std::string, std::vector<Blob>>& blobs,
                                      uint32_t mtime) {
        H5Writer w;
        w.alloc(96);
        std::map<std::string, Child> groups;
        for (const std::string& name : layers) {
            std::map<std::string, Child> kids;
            auto it = blobs.find(name);
            if (it != blobs.end())
                for (size_t i = 0; i < it->second.size(); ++i) kids[std::to_string(i)] = Child{w.dataset(it->second[i], mtime), false, 0, 0};
            groups[name] = w.group(kids);
        }
        for (auto& kv : blobs)
            if (!groups.count(kv.first)) throw std::runtime_error("model_io: layer without a group: " + kv.first);
        Child data = w.group(groups);
        std::map<std::string, Child> rootkids;
        rootkids["data"] = data;
        Child root = w.group(rootkids);
        while (w.b_.size() % 8) w.b_.push_back(0);
        const uint64_t eof = w.b_.size(), undef = ~0ull;
        std::vector<uint8_t> sb = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n', 0, 0, 0, 0, 0, 8, 8, 0};
        put16(sb, kLeafK); put16(sb, kIntK); put32(sb, 0);
        put64(sb, 0); put64(sb, undef); put64(sb, eof); put64(sb, undef);
        put64(sb, 0); put64(sb, root.ohdr); put32(sb, 1); put32(sb, 0); put64(sb, root.btree); put64(sb, root.heap);
        std::memcpy(w.b_.data(), sb.data(), 96);
        return w.b_;
    }

private:
    static constexpr int kLeafK = 4, kIntK = 16;
    struct Child { uint64_t ohdr = 0; bool is_group = false; uint64_t btree = 0, heap = 0; };
    std::vector<uint8_t> b_;

    static void put16(std::vector<uint8_t>& v, uint16_t x) { for (int i = 0; i < 2; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
    static void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
    static void put64(std::vector<uint8_t>& v, uint64_t x) { for (int i = 0; i < 8; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
    uint64_t alloc(size_t n) {
        while (b_.size() % 8) b_.push_back(0);
        uint64_t a = b_.size();
        b_.resize(b_.size() + n, 0);
        return a;
    }
    void put(uint64_t addr, const std::vector<uint8_t>& d) { std::memcpy(b_.data() + addr, d.data(), d.size()); }
    static std::vector<uint8_t> msg(uint16_t type, const std::vector<uint8_t>& body, uint8_t flags = 0) {
        std::vector<uint8_t> m;
        put16(m, type); put16(m, (uint16_t)body.size()); m.push_back(flags); m.push_back(0); m.push_back(0); m.push_back(0);
        m.insert(m.end(), body.begin(), body.end());
        return m;
    }
    uint64_t object_header(const std::vector<std::vector<uint8_t>>& msgs, int total) {
        std::vector<uint8_t> body;
        for (auto& m : msgs) body.insert(body.end(), m.begin(), m.end());
        uint16_t nmsg = (uint16_t)msgs.size();
        if (total > 0) {
            std::vector<uint8_t> nil = msg(0x0, std::vector<uint8_t>((size_t)total - body.size() - 8, 0));
            body.insert(body.end(), nil.begin(), nil.end());
            ++nmsg;
        }
        std::vector<uint8_t> hdr = {1, 0};
        put16(hdr, nmsg); put32(hdr, 1); put32(hdr, (uint32_t)body.size()); put32(hdr, 0);
        hdr.insert(hdr.end(), body.begin(), body.end());
        uint64_t a = alloc(hdr.size());
        put(a, hdr);
        return a;
    }
    uint64_t dataset(const Blob& bl, uint32_t mtime) {
        uint64_t count = 1;
        for (uint64_t dsz : bl.dims) count *= dsz;
        std::vector<uint8_t> space = {1, (uint8_t)bl.dims.size(), 1, 0, 0, 0, 0, 0};
        for (int rep = 0; rep < 2; ++rep)
            for (uint64_t dsz : bl.dims) put64(space, dsz);
        const uint64_t data_addr = alloc(count * 8);
        std::memcpy(b_.data() + data_addr, bl.data, count * 8);
        static const uint8_t f64[24] = {0x11, 0x20, 0x3f, 0x00, 0x08, 0x00, 0x00, 0x00, 0x00, 0x00, 0x40, 0x00, 0x34, 0x0b, 0x00, 0x34,
                                        0xff, 0x03, 0x00, 0x00, 0, 0, 0, 0};
        std::vector<uint8_t> layout = {3, 1};
        put64(layout, data_addr); put64(layout, count * 8);
        layout.resize(24, 0);
        std::vector<uint8_t> mt = {1, 0, 0, 0};
        put32(mt, mtime);
        return object_header({msg(0x1, space), msg(0x3, std::vector<uint8_t>(f64, f64 + 24), 1),
                              msg(0x5, {2, 2, 2, 1, 0, 0, 0, 0}, 1), msg(0x8, layout, 1), msg(0x12, mt)}, 256);
    }
    Child group(const std::map<std::string, Child>& children) {      // std::map iterates in strcmp order, like the symbol nodes
        std::vector<uint8_t> heap(8, 0);
        std::map<std::string, uint64_t> name_off;
        std::vector<std::string> names;
        for (auto& kv : children) {
            names.push_back(kv.first);
            name_off[kv.first] = heap.size();
            heap.insert(heap.end(), kv.first.begin(), kv.first.end());
            heap.push_back(0);
            while (heap.size() % 8) heap.push_back(0);
        }
        const uint64_t used = heap.size();
        uint64_t size = used + 16;
        size += (8 - size % 8) % 8;
        size = std::max<uint64_t>(88, size);
        put64(heap, 1); put64(heap, size - used);
        heap.resize(size, 0);
        const size_t cap = 2 * kLeafK;
        std::vector<uint64_t> snods;
        std::vector<std::string> last_names;
        for (size_t i = 0; i < std::max<size_t>(names.size(), 1); i += cap) {
            const size_t n = names.empty() ? 0 : std::min(cap, names.size() - i);
            std::vector<uint8_t> node = {'S', 'N', 'O', 'D', 1, 0};
            put16(node, (uint16_t)n);
            for (size_t k = 0; k < n; ++k) {
                const Child& c = children.at(names[i + k]);
                put64(node, name_off[names[i + k]]); put64(node, c.ohdr); put32(node, c.is_group ? 1 : 0); put32(node, 0);
                put64(node, c.is_group ? c.btree : 0); put64(node, c.is_group ? c.heap : 0);
            }
            uint64_t a = alloc(8 + 40 * cap);
            put(a, node);
            snods.push_back(a);
            last_names.push_back(n ? names[i + n - 1] : std::string());
        }
        const uint64_t bt = alloc(24 + 2 * kIntK * 16 + 8);
        std::vector<uint8_t> tree = {'T', 'R', 'E', 'E', 0, 0};
        put16(tree, (uint16_t)(names.empty() ? 0 : snods.size()));
        put64(tree, ~0ull); put64(tree, ~0ull); put64(tree, 0);
        if (!names.empty())
            for (size_t k = 0; k < snods.size(); ++k) { put64(tree, snods[k]); put64(tree, name_off[last_names[k]]); }
        put(bt, tree);
        const uint64_t hp = alloc(32);
        const uint64_t hd = alloc(size);
        put(hd, heap);
        std::vector<uint8_t> hh = {'H', 'E', 'A', 'P', 0, 0, 0, 0};
        put64(hh, size); put64(hh, used); put64(hh, hd);
        put(hp, hh);
        std::vector<uint8_t> st;
        put64(st, bt); put64(st, hp);
        Child c;
        c.ohdr = object_header({msg(0x11, st)}, 0);
        c.is_group = true; c.btree = bt; c.heap = hp;
        return c;
    }
};

inline void write_file(const std::string& path, const void* data, size_t n) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    std::fwrite(data, 1, n, f);
    std::fclose(f);
}
// cNeuralNet::WriteOffsetScale (learning/NeuralNet.cpp:1182-1205): same keys and order, %.17g numbers
inline void write_scale_file(const std::string& path, const double* in_off, const double* in_scale, int n_in, const double* out_off,
                             const double* out_scale, int n_out) {
    auto vec = [](const double* v, int n) {
        std::string s = "[";
        char buf[40];
        for (int i = 0; i < n; ++i) { std::snprintf(buf, sizeof(buf), "%.17g", v[i]); s += buf; if (i + 1 < n) s += ", "; }
        return s + "]";
    };
    std::string t = "{\n\"InputOffset\": " + vec(in_off, n_in) + ",\n\"InputScale\": " + vec(in_scale, n_in) + ",\n\"OutputOffset\": " +
                    vec(out_off, n_out) + ",\n\"OutputScale\": " + vec(out_scale, n_out) + "\n}";
    write_file(path, t.data(), t.size());
}

}  // namespace trl
