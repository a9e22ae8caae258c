// Launch accounting for bench.py: every kernel launch of the library is counted per family, and (when enabled)
// bracketed with CUDA events on its own stream so the per-family device time of a step can be measured live.
#pragma once
#include <cuda_runtime.h>

namespace spe {

enum Family : int { kFamGemm = 0, kFamAttention, kFamElementwise, kFamHeads, kFamCrop, kFamPnp, kNumFamilies };

struct ProfScope {
  ProfScope(Family f, cudaStream_t s);
  ~ProfScope();
  Family fam;
  cudaStream_t stream;
  cudaEvent_t stop = nullptr;
};

void profile_enable(bool on);
bool profile_timing_enabled();
void profile_peek_launches(long long* launches_by_family);
// account for kernels launched through a replayed CUDA graph (sign = +1) or undo capture-time counting (-1)
void profile_add_launches(const long long* launches_by_family, int sign);
// sums event-timed milliseconds and launch counts per family since the last collect; synchronises the events
void profile_collect(double* ms_by_family, long long* launches_by_family);

}  // namespace spe
