#include "shim.h"
#include "/root/reference/cuda_imp/score_cuda/src/score_computation_kernel.cu"
