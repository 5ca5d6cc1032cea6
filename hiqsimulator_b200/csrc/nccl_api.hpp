// NCCL entry points resolved at run time (dlopen) instead of at link time.
//
// Two NCCL builds with the same SONAME live in this image (the system one and the one bundled
// with PyTorch); a link-time dependency would pin whichever the loader meets first and break the
// other user in the same process.  Resolving lazily means: in a process that already loaded an
// NCCL (e.g. through torch.distributed) we share it, otherwise the system library is loaded the
// first time a multi-rank engine is created.  HIQ_NCCL_LIB overrides the library path.
#pragma once
#include <nccl.h>

namespace hiq {

struct NcclApi {
     decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
     decltype(&ncclCommInitRank) CommInitRank = nullptr;
     decltype(&ncclCommDestroy) CommDestroy = nullptr;
     decltype(&ncclGetErrorString) GetErrorString = nullptr;
     decltype(&ncclAllReduce) AllReduce = nullptr;
     decltype(&ncclBroadcast) Broadcast = nullptr;
     decltype(&ncclAllGather) AllGather = nullptr;
     decltype(&ncclSend) Send = nullptr;
     decltype(&ncclRecv) Recv = nullptr;
     decltype(&ncclGroupStart) GroupStart = nullptr;
     decltype(&ncclGroupEnd) GroupEnd = nullptr;
     decltype(&ncclGetVersion) GetVersion = nullptr;
};

// Loads the library on first use; returns HIQ_OK or records an error.
int nccl_load();
const NcclApi& nccl();

}  // namespace hiq
