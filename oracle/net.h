// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/README.md). Never linked into the product path.
//
// MACE actor-critic forward pass in double precision, restating Caffe's layer semantics for the topology in
// data/policies/dog/nets/dog_mace3_deploy.prototxt (Convolution = cross-correlation, NCHW with H = 1;
// InnerProduct flattens C x W channel-major), wrapped by cNeuralNet::Eval's offset/scale normalisation
// (learning/NeuralNet.cpp:352-375, 977-986, 1027-1036).  Caffe itself is an absent third-party dependency
// (niuzhiheng/caffe @ 7b3e6f2, README.md:14-15): parity of this forward pass is pinned only by the shipped
// weights' behaviour, not by a numeric reference vector ("parity unpinned").
#pragma once
#include <map>
#include <string>
#include <vector>

#include "pack_reader.h"

namespace orc {

struct Net {
    bool valid = false;
    int n_in = 0, n_char = 0, n_out = 0, n_frags = 0, frag = 0;
    std::vector<double> conv0_w, conv0_b, conv1_w, conv1_b, conv2_w, conv2_b, tip0_w, tip0_b, ip0_w, ip0_b;
    std::vector<double> head0_w[4], head0_b[4], head1_w[4], head1_b[4];  // val, a0, a1, a2
    std::vector<double> in_off, in_scale, out_off, out_scale;

    void load(const Pack& p) {
        if (!p.has("net_dims")) return;
        const auto& d = p.i32("net_dims");
        n_in = d[0]; n_char = d[1]; n_out = d[2]; n_frags = d[3]; frag = d[4];
        conv0_w = p.f64("net_terr_conv0_w"); conv0_b = p.f64("net_terr_conv0_b");
        conv1_w = p.f64("net_terr_conv1_w"); conv1_b = p.f64("net_terr_conv1_b");
        conv2_w = p.f64("net_terr_conv2_w"); conv2_b = p.f64("net_terr_conv2_b");
        tip0_w = p.f64("net_terr_ip0_w"); tip0_b = p.f64("net_terr_ip0_b");
        ip0_w = p.f64("net_ip0_w"); ip0_b = p.f64("net_ip0_b");
        const char* heads[4] = {"val", "a0", "a1", "a2"};
        for (int h = 0; h < 4; ++h) {
            std::string n = std::string("net_") + heads[h];
            head0_w[h] = p.f64(n + "_ip0_w"); head0_b[h] = p.f64(n + "_ip0_b");
            head1_w[h] = p.f64(n + "_ip1_w"); head1_b[h] = p.f64(n + "_ip1_b");
        }
        in_off = p.f64("net_in_offset"); in_scale = p.f64("net_in_scale");
        out_off = p.f64("net_out_offset"); out_scale = p.f64("net_out_scale");
        valid = true;
    }

    static void conv1d_relu(const std::vector<double>& x, int cin, int win, const std::vector<double>& w,
                            const std::vector<double>& b, int cout, int k, std::vector<double>& y) {
        int wout = win - k + 1;
        y.assign((size_t)cout * wout, 0.0);
        for (int o = 0; o < cout; ++o)
            for (int t = 0; t < wout; ++t) {
                double acc = b[o];
                for (int c = 0; c < cin; ++c)
                    for (int kk = 0; kk < k; ++kk) acc += w[((size_t)o * cin + c) * k + kk] * x[(size_t)c * win + t + kk];
                y[(size_t)o * wout + t] = acc > 0 ? acc : 0;
            }
    }
    static void fc(const std::vector<double>& x, const std::vector<double>& w, const std::vector<double>& b, int nout,
                   bool relu, std::vector<double>& y) {
        size_t nin = x.size();
        y.assign(nout, 0.0);
        for (int o = 0; o < nout; ++o) {
            double acc = b[o];
            const double* wr = &w[(size_t)o * nin];
            for (size_t i = 0; i < nin; ++i) acc += wr[i] * x[i];
            y[o] = (relu && acc < 0) ? 0 : acc;
        }
    }

    // cNeuralNet::Eval: y = net((x + off_in) * scale_in) / scale_out - off_out
    void eval(const double* x, double* y) const {
        std::vector<double> terr(200), chr(n_char), a, b, c, t, cat, h;
        for (int i = 0; i < 200; ++i) terr[i] = (x[i] + in_off[i]) * in_scale[i];
        for (int i = 0; i < n_char; ++i) chr[i] = (x[200 + i] + in_off[200 + i]) * in_scale[200 + i];
        conv1d_relu(terr, 1, 200, conv0_w, conv0_b, 16, 8, a);   // -> 16 x 193
        conv1d_relu(a, 16, 193, conv1_w, conv1_b, 32, 4, b);     // -> 32 x 190
        conv1d_relu(b, 32, 190, conv2_w, conv2_b, 32, 4, c);     // -> 32 x 187
        fc(c, tip0_w, tip0_b, 64, true, t);
        cat = t;
        cat.insert(cat.end(), chr.begin(), chr.end());           // concat0: [terr_relu3 | char]
        fc(cat, ip0_w, ip0_b, 256, true, h);
        std::vector<double> raw;
        for (int k = 0; k < 4; ++k) {
            std::vector<double> h0, h1;
            fc(h, head0_w[k], head0_b[k], 128, true, h0);
            fc(h0, head1_w[k], head1_b[k], k == 0 ? n_frags : frag, false, h1);
            raw.insert(raw.end(), h1.begin(), h1.end());
        }
        for (int i = 0; i < n_out; ++i) y[i] = raw[i] / out_scale[i] - out_off[i];
    }

    // cNeuralNet::GetLayerState (learning/NeuralNet.cpp:814-833): the named blob of the deploy net after a forward pass on x.  Every
    // layer of data/policies/dog/nets/dog_mace3_deploy.prototxt writes its own top blob (the ReLUs are not in place), so the
    // pre-activation values are blobs of their own.  Returns false for an unknown name.
    bool layer_state(const double* x, const std::string& name, std::vector<double>& out) const {
        auto conv = [](const std::vector<double>& in, int cin, int win, const std::vector<double>& w, const std::vector<double>& b, int cout, int k,
                       std::vector<double>& pre) {
            const int wout = win - k + 1;
            pre.assign((size_t)cout * wout, 0.0);
            for (int o = 0; o < cout; ++o)
                for (int t = 0; t < wout; ++t) {
                    double acc = b[o];
                    for (int c = 0; c < cin; ++c)
                        for (int kk = 0; kk < k; ++kk) acc += w[((size_t)o * cin + c) * k + kk] * in[(size_t)c * win + t + kk];
                    pre[(size_t)o * wout + t] = acc;
                }
        };
        auto relu = [](const std::vector<double>& v) { std::vector<double> r(v); for (double& e : r) e = e > 0 ? e : 0; return r; };
        std::map<std::string, std::vector<double>> blob;
        std::vector<double> data(n_in);
        for (int i = 0; i < n_in; ++i) data[i] = (x[i] + in_off[i]) * in_scale[i];
        blob["data"] = data;
        blob["data_terrain"].assign(data.begin(), data.begin() + 200);
        blob["data_char"].assign(data.begin() + 200, data.end());
        blob["char_flatten0"] = blob["data_char"];
        conv(blob["data_terrain"], 1, 200, conv0_w, conv0_b, 16, 8, blob["terr_conv0"]);
        blob["terr_relu0"] = relu(blob["terr_conv0"]);
        conv(blob["terr_relu0"], 16, 193, conv1_w, conv1_b, 32, 4, blob["terr_conv1"]);
        blob["terr_relu1"] = relu(blob["terr_conv1"]);
        conv(blob["terr_relu1"], 32, 190, conv2_w, conv2_b, 32, 4, blob["terr_conv2"]);
        blob["terr_relu2"] = relu(blob["terr_conv2"]);
        fc(blob["terr_relu2"], tip0_w, tip0_b, 64, false, blob["terr_ip0"]);
        blob["terr_relu3"] = relu(blob["terr_ip0"]);
        std::vector<double> cat = blob["terr_relu3"];
        cat.insert(cat.end(), blob["data_char"].begin(), blob["data_char"].end());
        blob["concat0"] = cat;
        fc(cat, ip0_w, ip0_b, 256, false, blob["ip0"]);
        blob["relu0"] = relu(blob["ip0"]);
        const char* heads[4] = {"val", "a0", "a1", "a2"};
        std::vector<double> outv;
        for (int k = 0; k < 4; ++k) {
            const std::string h = heads[k];
            fc(blob["relu0"], head0_w[k], head0_b[k], 128, false, blob[h + "_ip0"]);
            blob[h + "_relu0"] = relu(blob[h + "_ip0"]);
            fc(blob[h + "_relu0"], head1_w[k], head1_b[k], k == 0 ? n_frags : frag, false, blob[h + "_ip1"]);
            outv.insert(outv.end(), blob[h + "_ip1"].begin(), blob[h + "_ip1"].end());
        }
        blob["output"] = outv;
        auto it = blob.find(name);
        if (it == blob.end()) return false;
        out = it->second;
        return true;
    }
};

}  // namespace orc
