#include "ccd_common.cuh"
extern "C" int ccd_abi_version(void) { return 2; }
