// Library-wide state of the C ABI: error text, launch counter, ABI version.
#include <atomic>
#include <mutex>
#include <string>

#include "common.cuh"

namespace pnb {
static std::mutex g_mu;
static std::string g_err = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* where, cudaError_t e) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_err = std::string(where) + ": " + cudaGetErrorName(e) + " - " + cudaGetErrorString(e);
}
void set_error_msg(const char* msg) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_err = msg;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace pnb

extern "C" int pnb_abi_version(void) { return PNB_ABI_VERSION; }
extern "C" const char* pnb_last_error(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(pnb::g_mu);
  copy = pnb::g_err;
  return copy.c_str();
}
extern "C" long long pnb_launch_count(void) { return pnb::g_launches.load(std::memory_order_relaxed); }
