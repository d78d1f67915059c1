// slb_program.cu -- the step-program interpreter kernel (slb_program.cuh) and its launcher.
#define SLB_PROGRAM_IMPL
#include "slb_program.cuh"
