// ge_generate.cu -- device-side instance generation.  Filled in below.
#include "ge_common.cuh"
extern "C" int ge_generate(const ge_batch *, uint64_t, int32_t *, int32_t *, double *, float *, void *) { return GE_ERR_UNSUPPORTED; }
