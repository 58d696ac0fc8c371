// TEST INFRASTRUCTURE ONLY -- Caffe is not in this image.  Empty stand-ins for the Caffe / boost types learning/NeuralNet.h
// mentions; cNeuralNet itself is backed by oracle/ref_fake_sim.cpp (its Eval returns what the test installs).
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
template <typename T, typename U> shared_ptr<T> static_pointer_cast(const shared_ptr<U>& p) { return std::static_pointer_cast<T>(p); }
}  // namespace boost

namespace caffe {
enum Phase { TRAIN = 0, TEST = 1 };
class NetParameter {};
class SolverParameter {};
template <typename T> class Blob {};
template <typename T> class Layer {};
template <typename T> class Net {};
template <typename T> class Solver {};
template <typename T> class MemoryDataLayer : public Layer<T> {};
}  // namespace caffe
