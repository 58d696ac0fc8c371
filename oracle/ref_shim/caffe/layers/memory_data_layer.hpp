#pragma once
#include "caffe/caffe.hpp"
