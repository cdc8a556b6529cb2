// Library-wide state: the thread-local error string and the launch counter.
#include "common.cuh"

namespace aread {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> counter{0};
  return counter;
}

}  // namespace aread

extern "C" {

const char* aread_last_error(void) { return aread::last_error_buf(); }
int aread_abi_version(void) { return 12; }
uint64_t aread_launch_count(void) { return aread::launch_counter().load(std::memory_order_relaxed); }
void aread_launch_count_add(uint64_t n) { aread::launch_counter().fetch_add(n, std::memory_order_relaxed); }

}  // extern "C"
