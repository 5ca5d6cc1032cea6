#include "comm.hpp"

#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "hiq_host.hpp"
#include "nccl_api.hpp"

namespace hiq {

static int check_nccl(ncclResult_t r, const char* what)
{
     if (r == ncclSuccess) return HIQ_OK;
     return set_error(HIQ_ERR_CUDA, std::string(what) + ": " + nccl().GetErrorString(r));
}
#define HIQ_NCCL(call)                                 \
     do {                                              \
          int _rc = check_nccl((call), #call);         \
          if (_rc != HIQ_OK) return _rc;               \
     } while (0)

constexpr size_t kStageBytes = 4096;

Comm::~Comm()
{
     if (comm_) nccl().CommDestroy(comm_);
     if (stage_) cudaFree(stage_);
}

int Comm::init(int rank, int world_size, const void* unique_id, int device)
{
     if (world_size < 1 || (world_size & (world_size - 1)) || rank < 0 || rank >= world_size)
          return set_error(HIQ_ERR_ARG, "world size must be a power of two and 0 <= rank < world size");
     rank_ = rank;
     size_ = world_size;
     if (world_size == 1) return HIQ_OK;
     if (!unique_id) return set_error(HIQ_ERR_ARG, "world size > 1 needs the NCCL unique id of rank 0");
     HIQ_CUDA(cudaSetDevice(device));
     ncclUniqueId id;
     static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
     std::memcpy(&id, unique_id, sizeof(id));
     {
          int rc = nccl_load();
          if (rc != HIQ_OK) return rc;
     }
     // bind the descriptor socket BEFORE the NCCL rendezvous: once CommInitRank returns, every rank's socket exists
     if (fds_.open(std::string(static_cast<const char*>(unique_id), 128), rank) != HIQ_OK) {
          // not fatal: the swap falls back to the staged NCCL exchange
     }
     HIQ_NCCL(nccl().CommInitRank(&comm_, world_size, id, rank));
     HIQ_CUDA(cudaMalloc(&stage_, kStageBytes));
     return HIQ_OK;
}

Comm* Comm::shared(int rank, int world_size, const void* unique_id, int device)
{
     static std::mutex mu;
     static auto* cache = new std::map<std::string, Comm*>();  // intentionally leaked: outlives CUDA teardown order issues
     std::lock_guard<std::mutex> lock(mu);
     std::string key = world_size > 1 && unique_id ? std::string(static_cast<const char*>(unique_id), 128) : std::string("single");
     key += ":" + std::to_string(rank) + ":" + std::to_string(world_size) + ":" + std::to_string(device);
     auto it = cache->find(key);
     if (it != cache->end()) return it->second;
     Comm* c = new Comm();
     if (c->init(rank, world_size, unique_id, device) != HIQ_OK) {
          delete c;
          return nullptr;
     }
     (*cache)[key] = c;
     return c;
}

int Comm::allreduce_sum(double* vals, int n, cudaStream_t stream)
{
     if (size_ == 1) return HIQ_OK;
     if (static_cast<size_t>(n) * sizeof(double) > kStageBytes) return set_error(HIQ_ERR_ARG, "allreduce_sum: too many values");
     HIQ_CUDA(cudaMemcpyAsync(stage_, vals, n * sizeof(double), cudaMemcpyHostToDevice, stream));
     HIQ_NCCL(nccl().AllReduce(stage_, stage_, n, ncclDouble, ncclSum, comm_, stream));
     HIQ_CUDA(cudaMemcpyAsync(vals, stage_, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
     HIQ_CUDA(cudaStreamSynchronize(stream));
     return HIQ_OK;
}

int Comm::allreduce_sum_device(double* dev, int n, cudaStream_t stream)
{
     if (size_ == 1) return HIQ_OK;
     HIQ_NCCL(nccl().AllReduce(dev, dev, n, ncclDouble, ncclSum, comm_, stream));
     return HIQ_OK;
}

int Comm::broadcast_bytes(void* host, size_t bytes, int root, cudaStream_t stream)
{
     if (size_ == 1) return HIQ_OK;
     if (bytes > kStageBytes) return set_error(HIQ_ERR_ARG, "broadcast_bytes: message too large");
     if (rank_ == root) HIQ_CUDA(cudaMemcpyAsync(stage_, host, bytes, cudaMemcpyHostToDevice, stream));
     HIQ_NCCL(nccl().Broadcast(stage_, stage_, bytes, ncclChar, root, comm_, stream));
     HIQ_CUDA(cudaMemcpyAsync(host, stage_, bytes, cudaMemcpyDeviceToHost, stream));
     HIQ_CUDA(cudaStreamSynchronize(stream));
     return HIQ_OK;
}

int Comm::allgather(const double* dev_send, double* dev_recv, size_t n, cudaStream_t stream)
{
     if (size_ == 1) {
          if (dev_send != dev_recv)
               HIQ_CUDA(cudaMemcpyAsync(dev_recv, dev_send, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
          return HIQ_OK;
     }
     HIQ_NCCL(nccl().AllGather(dev_send, dev_recv, n, ncclDouble, comm_, stream));
     return HIQ_OK;
}

}  // namespace hiq

extern "C" int hiq_comm_unique_id(void* out128)
{
     if (!out128) return hiq::set_error(HIQ_ERR_ARG, "hiq_comm_unique_id: null output");
     int rc = hiq::nccl_load();
     if (rc != HIQ_OK) return rc;
     ncclUniqueId id;
     ncclResult_t r = hiq::nccl().GetUniqueId(&id);
     if (r != ncclSuccess) return hiq::set_error(HIQ_ERR_CUDA, std::string("ncclGetUniqueId: ") + hiq::nccl().GetErrorString(r));
     std::memcpy(out128, &id, sizeof(id));
     return HIQ_OK;
}
