// Two processes exchange a file descriptor over hiq::FdChannel (the transport of the peer-slab handshake).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>

#include "../../hiqsimulator_b200/csrc/peer_ipc.hpp"

int main()
{
     const std::string key = "fd-channel-test-" + std::to_string(getpid());
     pid_t pid = fork();
     if (pid == 0) {  // rank 1: sends 40 descriptors (more than a default datagram queue holds)
          hiq::FdChannel ch;
          if (ch.open(key, 1) != 0) return 2;
          for (int i = 0; i < 40; ++i) {
               int fd = memfd_create("hiq_test", 0);
               char buf[32];
               const int n = std::snprintf(buf, sizeof(buf), "chunk-%d", i);
               if (write(fd, buf, n) != n) return 3;
               hiq::FdMessage m;
               m.index = i;
               m.total = 40;
               m.size = n;
               m.epoch = 7;
               m.fd = fd;
               if (ch.send_fd(0, m, 20000) != 0) return 4;
               close(fd);
          }
          return 0;
     }
     hiq::FdChannel ch;
     usleep(200 * 1000);  // the sender must retry until this socket is bound
     if (ch.open(key, 0) != 0) return 5;
     for (int i = 0; i < 40; ++i) {
          hiq::FdMessage m;
          if (ch.recv_fd(m, 20000) != 0) return 6;
          if (m.src_rank != 1 || m.index != static_cast<uint32_t>(i) || m.total != 40 || m.epoch != 7) return 7;
          char buf[32] = {0};
          lseek(m.fd, 0, SEEK_SET);
          if (read(m.fd, buf, sizeof(buf) - 1) != static_cast<ssize_t>(m.size)) return 8;
          char want[32];
          std::snprintf(want, sizeof(want), "chunk-%d", i);
          if (std::strcmp(buf, want) != 0) return 9;
          close(m.fd);
     }
     int st = 0;
     waitpid(pid, &st, 0);
     if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) return 10;
     std::puts("FD_CHANNEL_OK");
     return 0;
}
