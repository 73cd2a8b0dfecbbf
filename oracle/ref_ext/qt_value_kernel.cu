#include "shim.h"
#include "/root/reference/cuda_imp/QuadTreeAttention/QuadtreeAttention/src/value_aggregation_kernel.cu"
