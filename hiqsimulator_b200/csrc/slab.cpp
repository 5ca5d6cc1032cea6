#include "slab.hpp"

#include <algorithm>
#include <string>

#include "hiq_host.hpp"

namespace hiq {

struct DriverApi;
static int check_cu(CUresult r, const char* what);

// Driver entry points are resolved through the (statically linked) runtime so that the library has
// no link-time dependency on libcuda.so.1 and still loads on a machine without a driver.
template <class Fn>
static int driver_fn(const char* name, Fn& fn)
{
     void* p = nullptr;
     cudaDriverEntryPointQueryResult q;
     cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
     if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
          cudaGetLastError();
          return set_error(HIQ_ERR_CUDA, std::string("CUDA driver entry point not available: ") + name);
     }
     fn = reinterpret_cast<Fn>(p);
     return HIQ_OK;
}

struct DriverApi {
     decltype(&cuGetErrorString) GetErrorString = nullptr;
     decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
     decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
     decltype(&cuMemAddressFree) MemAddressFree = nullptr;
     decltype(&cuMemCreate) MemCreate = nullptr;
     decltype(&cuMemRelease) MemRelease = nullptr;
     decltype(&cuMemMap) MemMap = nullptr;
     decltype(&cuMemUnmap) MemUnmap = nullptr;
     decltype(&cuMemSetAccess) MemSetAccess = nullptr;
     bool ready = false;
     int load()
     {
          if (ready) return HIQ_OK;
          int rc;
          if ((rc = driver_fn("cuGetErrorString", GetErrorString))) return rc;
          if ((rc = driver_fn("cuMemGetAllocationGranularity", MemGetAllocationGranularity))) return rc;
          if ((rc = driver_fn("cuMemAddressReserve", MemAddressReserve))) return rc;
          if ((rc = driver_fn("cuMemAddressFree", MemAddressFree))) return rc;
          if ((rc = driver_fn("cuMemCreate", MemCreate))) return rc;
          if ((rc = driver_fn("cuMemRelease", MemRelease))) return rc;
          if ((rc = driver_fn("cuMemMap", MemMap))) return rc;
          if ((rc = driver_fn("cuMemUnmap", MemUnmap))) return rc;
          if ((rc = driver_fn("cuMemSetAccess", MemSetAccess))) return rc;
          ready = true;
          return HIQ_OK;
     }
};
static DriverApi g_drv;

static int check_cu(CUresult r, const char* what)
{
     if (r == CUDA_SUCCESS) return HIQ_OK;
     const char* s = nullptr;
     if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
     return set_error(HIQ_ERR_CUDA, std::string(what) + ": " + (s ? s : "unknown CUDA driver error"));
}

#define HIQ_CU(call)                                   \
     do {                                              \
          int _rc = check_cu((call), #call);           \
          if (_rc != HIQ_OK) return _rc;               \
     } while (0)

int Slab::init(int device, uint64_t max_amps)
{
     device_ = device;
     HIQ_CUDA(cudaSetDevice(device));
     HIQ_CUDA(cudaFree(nullptr));  // make sure the primary context exists
     {
          int rc = g_drv.load();
          if (rc != HIQ_OK) return rc;
     }
     CUmemAllocationProp prop = {};
     prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
     prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
     prop.location.id = device;
     HIQ_CU(g_drv.MemGetAllocationGranularity(&gran_, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
     size_t total = 0, free_b = 0;
     HIQ_CUDA(cudaMemGetInfo(&free_b, &total));
     size_t want = static_cast<size_t>(max_amps) * sizeof(double2);
     want = std::min(want, total);  // more than the device holds can never be mapped
     want = std::max(want, gran_);
     reserved_ = (want + gran_ - 1) / gran_ * gran_;
     HIQ_CU(g_drv.MemAddressReserve(&base_, reserved_, 0, 0, 0));
     return HIQ_OK;
}

int Slab::ensure(uint64_t amps)
{
     const size_t need = static_cast<size_t>(amps) * sizeof(double2);
     if (need <= mapped_) return HIQ_OK;
     if (need > reserved_)
          return set_error(HIQ_ERR_RUNTIME, "state vector of " + std::to_string(need >> 20) +
                                                " MiB exceeds the reserved slab (max_local / device memory)");
     const size_t grow = (need - mapped_ + gran_ - 1) / gran_ * gran_;
     CUmemAllocationProp prop = {};
     prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
     prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
     prop.location.id = device_;
     CUmemGenericAllocationHandle h;
     HIQ_CU(g_drv.MemCreate(&h, grow, &prop, 0));
     CUresult r = g_drv.MemMap(base_ + mapped_, grow, 0, h, 0);
     if (r != CUDA_SUCCESS) {
          g_drv.MemRelease(h);
          return check_cu(r, "cuMemMap");
     }
     CUmemAccessDesc acc = {};
     acc.location = prop.location;
     acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
     r = g_drv.MemSetAccess(base_ + mapped_, grow, &acc, 1);
     if (r != CUDA_SUCCESS) {
          g_drv.MemUnmap(base_ + mapped_, grow);
          g_drv.MemRelease(h);
          return check_cu(r, "cuMemSetAccess");
     }
     chunks_.emplace_back(h, grow);
     mapped_ += grow;
     return HIQ_OK;
}

void Slab::release()
{
     if (!base_) return;
     cudaDeviceSynchronize();
     size_t off = 0;
     for (auto& c: chunks_) {
          g_drv.MemUnmap(base_ + off, c.second);
          g_drv.MemRelease(c.first);
          off += c.second;
     }
     chunks_.clear();
     g_drv.MemAddressFree(base_, reserved_);
     base_ = 0;
     mapped_ = reserved_ = 0;
}

}  // namespace hiq
