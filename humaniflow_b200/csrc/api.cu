// Library bookkeeping entry points.
#include "common.cuh"

namespace hf {
thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};
}  // namespace hf

extern "C" int hf_version(void) { return 100; }
extern "C" const char* hf_last_error(void) { return hf::g_err; }
extern "C" long long hf_launch_count(void) { return hf::g_launches.load(); }
