// oracle/ref_shim/cuda/helpers.h -- TEST INFRASTRUCTURE. Stand-in for the header of the same name that
// /root/reference/optixSphere.cu includes (OptiX SDK 8.0.0 / CUDA toolkit); everything lives in shim_all.h.
#pragma once
#include "shim_all.h"
