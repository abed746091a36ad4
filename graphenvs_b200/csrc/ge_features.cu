// ge_features.cu -- structural node features (feature_extraction.py:6-37).  Filled in below.
#include "ge_common.cuh"
extern "C" int ge_features(const ge_batch *, void *) { return GE_ERR_UNSUPPORTED; }
