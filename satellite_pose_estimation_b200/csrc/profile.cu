#include "profile.h"

#include <atomic>
#include <mutex>
#include <vector>

namespace spe {

namespace {
struct Rec { Family fam; cudaEvent_t start, stop; };
std::mutex g_mu;
std::vector<Rec> g_recs;
std::atomic<bool> g_timing{false};
std::atomic<long long> g_launches[kNumFamilies];
}  // namespace

ProfScope::ProfScope(Family f, cudaStream_t s) : fam(f), stream(s) {
  g_launches[f].fetch_add(1, std::memory_order_relaxed);
  if (!g_timing.load(std::memory_order_relaxed)) return;
  cudaEvent_t start;
  if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) { stop = nullptr; return; }
  cudaEventRecord(start, stream);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back(Rec{f, start, stop});
}

ProfScope::~ProfScope() {
  if (stop) cudaEventRecord(stop, stream);
}

void profile_enable(bool on) { g_timing.store(on); }
bool profile_timing_enabled() { return g_timing.load(); }
void profile_peek_launches(long long* out) {
  for (int i = 0; i < kNumFamilies; ++i) out[i] = g_launches[i].load();
}
void profile_add_launches(const long long* n, int sign) {
  for (int i = 0; i < kNumFamilies; ++i) g_launches[i].fetch_add(sign * n[i]);
}

void profile_collect(double* ms, long long* launches) {
  for (int i = 0; i < kNumFamilies; ++i) {
    ms[i] = 0.0;
    launches[i] = g_launches[i].exchange(0);
  }
  std::lock_guard<std::mutex> lk(g_mu);
  for (const Rec& r : g_recs) {
    float t = 0.f;
    if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&t, r.start, r.stop) == cudaSuccess)
      ms[r.fam] += t;
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  g_recs.clear();
}

}  // namespace spe
