// ABI bookkeeping entry points.
#include "launch_count.h"
#include "../../include/unopose_b200.h"

namespace upk {
std::atomic<unsigned long long> g_launches{0};
}

extern "C" {
int upk_abi_version(void) { return 1; }
int upk_built_sm(void) { return 100; }
unsigned long long upk_launch_count(void) { return upk::g_launches.load(); }
}
