// tests only: stands in for the CUDA runtime header when the kernel sources are compiled for the SIMT emulator
#pragma once
#include "../simt_runtime.h"
