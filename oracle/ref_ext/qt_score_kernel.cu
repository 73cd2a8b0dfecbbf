#include "shim.h"
#include "/root/reference/cuda_imp/QuadTreeAttention/QuadtreeAttention/src/score_computation_kernal.cu"
