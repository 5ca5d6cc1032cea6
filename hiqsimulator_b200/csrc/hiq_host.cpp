#include "hiq_host.hpp"

#include <atomic>

namespace hiq {

static thread_local std::string g_last_error;
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const std::string& msg)
{
     g_last_error = msg;
     return code;
}

int check_cuda(cudaError_t e, const char* what)
{
     if (e == cudaSuccess) return HIQ_OK;
     return set_error(HIQ_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int check_launch(const char* what) { return check_cuda(cudaGetLastError(), what); }

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms()
{
     static std::atomic<int> cache[64];
     int dev = 0;
     if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
          cudaGetLastError();
          return 148;
     }
     int n = cache[dev].load(std::memory_order_relaxed);
     if (n == 0) {
          if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
               cudaGetLastError();
               n = 148;
          }
          cache[dev].store(n, std::memory_order_relaxed);
     }
     return n;
}

static std::atomic<uint64_t> g_max_grid{0};
uint64_t grid_cap(uint64_t natural)
{
     const uint64_t m = g_max_grid.load(std::memory_order_relaxed);
     return (m != 0 && m < natural) ? m : natural;
}

}  // namespace hiq

extern "C" const char* hiq_last_error(void) { return hiq::g_last_error.c_str(); }
extern "C" const char* hiq_version(void) { return "hiq_b200 0.1.0 sm_100a"; }
extern "C" uint64_t hiqk_launch_count(void) { return hiq::g_launches.load(); }
extern "C" int hiq_device_count(void)
{
     int n = 0;
     if (cudaGetDeviceCount(&n) != cudaSuccess) {
          cudaGetLastError();
          return -1;
     }
     return n;
}

extern "C" int hiqk_debug_set_max_grid(int max_ctas)
{
     hiq::g_max_grid.store(max_ctas > 0 ? static_cast<uint64_t>(max_ctas) : 0, std::memory_order_relaxed);
     return HIQ_OK;
}
