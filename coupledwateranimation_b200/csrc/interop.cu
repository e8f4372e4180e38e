// interop.cu -- placeholder until the CUDA-GL interop entry points land
#include "internal.cuh"
