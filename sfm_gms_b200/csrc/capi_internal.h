// capi_internal.h — C++ entry points shared between capi.cu and multi.cu (not part of the C ABI).
#pragma once
#include <stdint.h>

struct sfmgms_ctx;

namespace sfmgms {

// One shard of a multi-GPU compact run: host outputs only; chunks are appended to the caller's ONE buffer through
// *shared_cursor (atomic reservation per chunk), inlier_begin[p] is the absolute first row of pair p.
int match_pairs_compact_shared(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                               double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int64_t* inlier_begin,
                               void* matches, float* pts1, float* pts2, int64_t capacity, int64_t* shared_cursor);

}  // namespace sfmgms
