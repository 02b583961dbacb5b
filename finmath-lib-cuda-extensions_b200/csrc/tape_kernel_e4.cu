// tape_kernel_e4.cu — the op-tape interpreter for chunks of 128 paths (4 elements per lane); see tape_interp.cuh
#define TE 4
#include "tape_interp.cuh"
