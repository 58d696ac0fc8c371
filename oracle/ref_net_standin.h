// TEST INFRASTRUCTURE ONLY -- cNeuralNet (Caffe in the reference, learning/NeuralNet.cpp) for the libraries of reference sources
// compiled into oracle/_ref: included by exactly one translation unit per library (ref_ctrl_api.cpp, ref_train_api.cpp).
// The reference's controllers, scenarios and trainers reach the network only through this class; here every operation is handed to
// the test through callbacks (NetHooks), or -- for the controller tests that install one output vector per decision -- answered
// from that installed vector.  Instances are numbered in construction order so that the test can tell the controller's network,
// the trainer's network and its target apart.
#pragma once
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "learning/NeuralNet.h"

extern "C" void ref_abort_stub();

typedef void (*eval_fn)(int net, const double* X, int B, double* Y, void* user);
typedef void (*train_fn)(int net, const double* X, const double* Y, int B, void* user);
typedef void (*copy_fn)(int dst, int src, void* user);
typedef void (*calc_os_fn)(const double* X, int n, double* off, double* scale, void* user);
typedef void (*set_os_fn)(int net, const double* off, const double* scale, void* user);
struct NetHooks {
    int n_in = 0, n_out = 0, batch = 32;
    eval_fn eval = nullptr;
    train_fn train = nullptr;
    copy_fn copy = nullptr;
    calc_os_fn calc_os = nullptr;
    set_os_fn set_os = nullptr;
    void* user = nullptr;
};
static NetHooks g_hooks;
static std::map<const cNeuralNet*, int> g_net_id;
static int g_next_net = 0;
static int id_of(const cNeuralNet* n) { return g_net_id.at(n); }

// the installed-output mode (ref_ctrl_set_net_output) and the single-network callback of the scenario tests
typedef void (*net_fn)(const double* x, int n_in, double* y, int n_out, void* user);
static Eigen::VectorXd g_net_output;
static int g_net_in = 0, g_net_out = 0;
static Eigen::VectorXd g_out_scale;
static net_fn g_net_cb = nullptr;
static void* g_net_user = nullptr;
// cNeuralNet::GetLayerState (learning/NeuralNet.cpp:814-833): the test answers with the named blob of the network it stands in for,
// for the input of the most recent Eval (Caffe's blobs hold the last forward pass); returns the blob's size, 0 = unknown layer
typedef int (*layer_fn)(const char* name, const double* x_last, int n_in, double* out, int cap, void* user);
static layer_fn g_layer_cb = nullptr;
static std::vector<double> g_last_x;

static std::vector<double> flat(const Eigen::MatrixXd& M) {
    std::vector<double> v((size_t)M.rows() * M.cols());
    for (int i = 0; i < (int)M.rows(); ++i)
        for (int j = 0; j < (int)M.cols(); ++j) v[(size_t)i * M.cols() + j] = M(i, j);
    return v;
}

std::mutex cNeuralNet::gOutputLock;
cNeuralNet::tProblem::tProblem() : mPassesPerStep(1) {}
bool cNeuralNet::tProblem::HasData() const { return mX.size() > 0; }
cNeuralNet::cNeuralNet() : mValidModel(true), mAsync(false) { g_net_id[this] = g_next_net++; }
cNeuralNet::~cNeuralNet() { g_net_id.erase(this); }
void cNeuralNet::LoadNet(const std::string&) {}
void cNeuralNet::LoadModel(const std::string&) {}
void cNeuralNet::LoadSolver(const std::string&, bool) {}
void cNeuralNet::LoadScale(const std::string&) {}
void cNeuralNet::Clear() {}
void cNeuralNet::ResetSolver() {}
void cNeuralNet::OutputModel(const std::string&) const {}
bool cNeuralNet::HasNet() const { return g_hooks.n_out > 0 || g_net_out > 0; }
bool cNeuralNet::HasSolver() const { return true; }
bool cNeuralNet::HasLayer(const std::string) const { return false; }
bool cNeuralNet::HasValidModel() const { return true; }
int cNeuralNet::GetInputSize() const { return g_hooks.n_in ? g_hooks.n_in : g_net_in; }
int cNeuralNet::GetOutputSize() const { return g_hooks.n_out ? g_hooks.n_out : g_net_out; }
int cNeuralNet::GetBatchSize() const { return g_hooks.batch; }
const Eigen::VectorXd& cNeuralNet::GetOutputScale() const { return g_out_scale; }
void cNeuralNet::Train(const tProblem& prob) {
    const std::vector<double> x = flat(prob.mX), y = flat(prob.mY);
    g_hooks.train(id_of(this), x.data(), y.data(), (int)prob.mX.rows(), g_hooks.user);
}
void cNeuralNet::EvalBatch(const Eigen::MatrixXd& X, Eigen::MatrixXd& out_Y) const {
    const int B = (int)X.rows(), no = GetOutputSize();
    const std::vector<double> x = flat(X);
    std::vector<double> y((size_t)B * no);
    g_hooks.eval(id_of(this), x.data(), B, y.data(), g_hooks.user);
    out_Y.resize(B, no);
    for (int i = 0; i < B; ++i)
        for (int j = 0; j < no; ++j) out_Y(i, j) = y[(size_t)i * no + j];
}
void cNeuralNet::Eval(const Eigen::VectorXd& x, Eigen::VectorXd& out_y) const {
    if (!g_hooks.eval && !g_net_cb) { out_y = g_net_output; return; }
    const int no = GetOutputSize();
    std::vector<double> xi(x.size()), y(no);
    for (int i = 0; i < (int)x.size(); ++i) xi[i] = x[i];
    g_last_x = xi;
    if (g_hooks.eval) g_hooks.eval(id_of(this), xi.data(), 1, y.data(), g_hooks.user);
    else g_net_cb(xi.data(), (int)xi.size(), y.data(), no, g_net_user);
    out_y.resize(no);
    for (int j = 0; j < no; ++j) out_y[j] = y[j];
}
void cNeuralNet::GetLayerState(const std::string& layer_name, Eigen::VectorXd& out_state) const {
    std::vector<double> buf(16384);
    const int n = g_layer_cb ? g_layer_cb(layer_name.c_str(), g_last_x.data(), (int)g_last_x.size(), buf.data(), (int)buf.size(), g_net_user) : 0;
    if (n <= 0) { printf("Can't find layer named %s\n", layer_name.c_str()); out_state.resize(0); return; }
    out_state.resize(n);
    for (int i = 0; i < n; ++i) out_state[i] = buf[i];
}
void cNeuralNet::CopyModel(const cNeuralNet& other) { if (g_hooks.copy) g_hooks.copy(id_of(this), id_of(&other), g_hooks.user); }
void cNeuralNet::CalcOffsetScale(const Eigen::MatrixXd& X, Eigen::VectorXd& out_offset, Eigen::VectorXd& out_scale) const {
    const std::vector<double> x = flat(X);
    std::vector<double> off(X.cols()), sc(X.cols());
    g_hooks.calc_os(x.data(), (int)X.rows(), off.data(), sc.data(), g_hooks.user);
    out_offset.resize(X.cols()); out_scale.resize(X.cols());
    for (int j = 0; j < (int)X.cols(); ++j) { out_offset[j] = off[j]; out_scale[j] = sc[j]; }
}
void cNeuralNet::SetInputOffsetScale(const Eigen::VectorXd& offset, const Eigen::VectorXd& scale) {
    std::vector<double> off(offset.size()), sc(scale.size());
    for (int j = 0; j < (int)offset.size(); ++j) { off[j] = offset[j]; sc[j] = scale[j]; }
    g_hooks.set_os(id_of(this), off.data(), sc.data(), g_hooks.user);
}
void cNeuralNet::ForwardInjectNoisePrefilled(double, double, const std::string&, Eigen::VectorXd&) const { ref_abort_stub(); }
