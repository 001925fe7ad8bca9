// Library identification + device query helpers of the C ABI.
#include "common.cuh"

extern "C" const char* ssd3d_version(void) { return "ssd3d_b200 sm_100a 0.1"; }
