#include "nccl_api.hpp"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <string>

#include "hiq_host.hpp"

namespace hiq {

static NcclApi g_api;
static bool g_loaded = false;
static std::mutex g_mu;

const NcclApi& nccl() { return g_api; }

template <class Fn>
static bool sym(void* h, const char* name, Fn& fn)
{
     fn = reinterpret_cast<Fn>(dlsym(h, name));
     return fn != nullptr;
}

int nccl_load()
{
     std::lock_guard<std::mutex> lock(g_mu);
     if (g_loaded) return HIQ_OK;
     const char* override_path = std::getenv("HIQ_NCCL_LIB");
     const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
     void* h = nullptr;
     std::string err;
     for (const char* n: names) {
          if (!n || !*n) continue;
          h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
          if (h) break;
          err = dlerror();
     }
     if (!h) return set_error(HIQ_ERR_CUDA, "cannot load NCCL: " + err);
     bool ok = sym(h, "ncclGetUniqueId", g_api.GetUniqueId) && sym(h, "ncclCommInitRank", g_api.CommInitRank) &&
               sym(h, "ncclCommDestroy", g_api.CommDestroy) && sym(h, "ncclGetErrorString", g_api.GetErrorString) &&
               sym(h, "ncclAllReduce", g_api.AllReduce) && sym(h, "ncclBroadcast", g_api.Broadcast) &&
               sym(h, "ncclAllGather", g_api.AllGather) && sym(h, "ncclSend", g_api.Send) && sym(h, "ncclRecv", g_api.Recv) &&
               sym(h, "ncclGroupStart", g_api.GroupStart) && sym(h, "ncclGroupEnd", g_api.GroupEnd) &&
               sym(h, "ncclGetVersion", g_api.GetVersion);
     if (!ok) return set_error(HIQ_ERR_CUDA, "NCCL library lacks a required entry point");
     g_loaded = true;
     return HIQ_OK;
}

}  // namespace hiq
