// tape_kernel_e16.cu — the op-tape interpreter for chunks of 512 paths (16 elements per lane); see tape_interp.cuh
#define TE 16
#include "tape_interp.cuh"
