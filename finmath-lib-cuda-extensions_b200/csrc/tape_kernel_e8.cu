// tape_kernel_e8.cu — the op-tape interpreter for chunks of 256 paths (8 elements per lane); see tape_interp.cuh
#define TE 8
#include "tape_interp.cuh"
