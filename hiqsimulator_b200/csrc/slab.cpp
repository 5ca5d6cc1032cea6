#include "slab.hpp"

#include <unistd.h>

#include <algorithm>
#include <string>
#include <cstdlib>

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
     decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
     decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
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
          if ((rc = driver_fn("cuMemExportToShareableHandle", MemExportToShareableHandle))) return rc;
          if ((rc = driver_fn("cuMemImportFromShareableHandle", MemImportFromShareableHandle))) return rc;
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

int Slab::init(int device, uint64_t max_amps, bool shareable)
{
     device_ = device;
     shareable_ = shareable;
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
     if (shareable_) prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
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
     if (shareable_) prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
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

int Slab::export_chunk(size_t i, int* fd) const
{
     if (!shareable_) return set_error(HIQ_ERR_RUNTIME, "slab was not created shareable");
     if (i >= chunks_.size()) return set_error(HIQ_ERR_ARG, "export_chunk: index out of range");
     int out = -1;
     HIQ_CU(g_drv.MemExportToShareableHandle(&out, chunks_[i].first, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
     *fd = out;
     return HIQ_OK;
}

// ---------------------------------------------------------------------------------------------
PeerSlab& PeerSlab::operator=(PeerSlab&& o) noexcept
{
     if (this != &o) {
          release();
          device_ = o.device_;
          base_ = o.base_;
          reserved_ = o.reserved_;
          mapped_ = o.mapped_;
          chunks_ = std::move(o.chunks_);
          epoch = o.epoch;
          o.base_ = 0;
          o.reserved_ = o.mapped_ = 0;
          o.chunks_.clear();
     }
     return *this;
}

int PeerSlab::init(int local_device, size_t reserve_bytes)
{
     release();
     device_ = local_device;
     {
          int rc = g_drv.load();
          if (rc != HIQ_OK) return rc;
     }
     reserved_ = reserve_bytes;
     HIQ_CU(g_drv.MemAddressReserve(&base_, reserved_, 0, 0, 0));
     return HIQ_OK;
}

int PeerSlab::map_next_chunk(int fd, size_t bytes)
{
     if (!base_ || mapped_ + bytes > reserved_) {
          ::close(fd);
          return set_error(HIQ_ERR_RUNTIME, "peer slab: chunk does not fit the reserved range");
     }
     CUmemGenericAllocationHandle h;
     CUresult r = g_drv.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                                     CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
     ::close(fd);  // the imported handle keeps the allocation alive
     if (r != CUDA_SUCCESS) return check_cu(r, "cuMemImportFromShareableHandle");
     r = g_drv.MemMap(base_ + mapped_, bytes, 0, h, 0);
     if (r != CUDA_SUCCESS) {
          g_drv.MemRelease(h);
          return check_cu(r, "cuMemMap (peer)");
     }
     CUmemAccessDesc acc = {};
     acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
     acc.location.id = device_;
     acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
     r = g_drv.MemSetAccess(base_ + mapped_, bytes, &acc, 1);
     if (r != CUDA_SUCCESS) {
          g_drv.MemUnmap(base_ + mapped_, bytes);
          g_drv.MemRelease(h);
          return check_cu(r, "cuMemSetAccess (peer)");
     }
     chunks_.emplace_back(h, bytes);
     mapped_ += bytes;
     return HIQ_OK;
}

void PeerSlab::release()
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
