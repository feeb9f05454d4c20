// Host-side launch counter shared by all translation units (bench.py reports it
// as "gpu_launches").
#pragma once
#include <atomic>
namespace upk {
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace upk
