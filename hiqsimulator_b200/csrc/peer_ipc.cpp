#include "peer_ipc.hpp"

#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#include "hiq_host.hpp"

namespace hiq {

namespace {
struct Wire {
     int32_t src_rank;
     uint32_t index, total;
     uint64_t size, epoch;
};

socklen_t fill_addr(sockaddr_un& a, const std::string& name)
{
     std::memset(&a, 0, sizeof(a));
     a.sun_family = AF_UNIX;
     // abstract namespace: sun_path[0] == 0, no file system entry, vanishes with the process
     const size_t n = std::min(name.size(), sizeof(a.sun_path) - 2);
     std::memcpy(a.sun_path + 1, name.data(), n);
     return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + n);
}
}  // namespace

FdChannel::~FdChannel()
{
     if (sock_ >= 0) ::close(sock_);
}

std::string FdChannel::name_of(int rank) const { return prefix_ + "_" + std::to_string(rank); }

int FdChannel::open(const std::string& world_key, int rank)
{
     if (sock_ >= 0) return HIQ_OK;
     uint64_t h = 1469598103934665603ull;  // FNV-1a of the world key
     for (unsigned char c: world_key) h = (h ^ c) * 1099511628211ull;
     char buf[64];
     std::snprintf(buf, sizeof(buf), "hiq_b200_%016llx", static_cast<unsigned long long>(h));
     prefix_ = buf;
     rank_ = rank;
     sock_ = ::socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
     if (sock_ < 0) return set_error(HIQ_ERR_RUNTIME, std::string("peer ipc: socket(): ") + std::strerror(errno));
     sockaddr_un a;
     const socklen_t len = fill_addr(a, name_of(rank));
     if (::bind(sock_, reinterpret_cast<sockaddr*>(&a), len) != 0) {
          const int e = errno;
          ::close(sock_);
          sock_ = -1;
          return set_error(HIQ_ERR_RUNTIME, std::string("peer ipc: bind(): ") + std::strerror(e));
     }
     return HIQ_OK;
}

int FdChannel::send_fd(int dst_rank, const FdMessage& m, int timeout_ms)
{
     if (sock_ < 0) return set_error(HIQ_ERR_RUNTIME, "peer ipc: channel not open");
     Wire w{rank_, m.index, m.total, m.size, m.epoch};
     iovec iov{&w, sizeof(w)};
     alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
     std::memset(ctrl, 0, sizeof(ctrl));
     sockaddr_un a;
     const socklen_t len = fill_addr(a, name_of(dst_rank));
     msghdr msg{};
     msg.msg_name = &a;
     msg.msg_namelen = len;
     msg.msg_iov = &iov;
     msg.msg_iovlen = 1;
     msg.msg_control = ctrl;
     msg.msg_controllen = sizeof(ctrl);
     cmsghdr* c = CMSG_FIRSTHDR(&msg);
     c->cmsg_level = SOL_SOCKET;
     c->cmsg_type = SCM_RIGHTS;
     c->cmsg_len = CMSG_LEN(sizeof(int));
     std::memcpy(CMSG_DATA(c), &m.fd, sizeof(int));
     const auto t0 = std::chrono::steady_clock::now();
     for (;;) {
          if (::sendmsg(sock_, &msg, MSG_DONTWAIT) >= 0) return HIQ_OK;
          const int e = errno;
          // peer not bound yet / its queue is full: retry for a while
          const bool retry = e == ECONNREFUSED || e == ENOENT || e == EAGAIN || e == ENOBUFS || e == EINTR;
          const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
          if (!retry || ms > timeout_ms)
               return set_error(HIQ_ERR_RUNTIME, "peer ipc: sendmsg to rank " + std::to_string(dst_rank) + ": " + std::strerror(e));
          std::this_thread::sleep_for(std::chrono::milliseconds(2));
     }
}

int FdChannel::recv_fd(FdMessage& m, int timeout_ms)
{
     if (sock_ < 0) return set_error(HIQ_ERR_RUNTIME, "peer ipc: channel not open");
     pollfd p{sock_, POLLIN, 0};
     int pr;
     do {
          pr = ::poll(&p, 1, timeout_ms);
     } while (pr < 0 && errno == EINTR);
     if (pr <= 0) return set_error(HIQ_ERR_RUNTIME, "peer ipc: timed out waiting for a peer's memory handle");
     Wire w{};
     iovec iov{&w, sizeof(w)};
     alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
     msghdr msg{};
     msg.msg_iov = &iov;
     msg.msg_iovlen = 1;
     msg.msg_control = ctrl;
     msg.msg_controllen = sizeof(ctrl);
     const ssize_t n = ::recvmsg(sock_, &msg, MSG_CMSG_CLOEXEC);
     if (n != static_cast<ssize_t>(sizeof(w))) return set_error(HIQ_ERR_RUNTIME, "peer ipc: short message");
     m.fd = -1;
     for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c))
          if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) std::memcpy(&m.fd, CMSG_DATA(c), sizeof(int));
     if (m.fd < 0) return set_error(HIQ_ERR_RUNTIME, "peer ipc: message without a descriptor");
     m.src_rank = w.src_rank;
     m.index = w.index;
     m.total = w.total;
     m.size = w.size;
     m.epoch = w.epoch;
     return HIQ_OK;
}

}  // namespace hiq
