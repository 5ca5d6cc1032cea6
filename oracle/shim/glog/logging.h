// Stand-in for <glog/logging.h> (test infrastructure; glog is not installed).
// Logging is discarded; CHECK() aborts like glog does, which the reference
// schedulers rely on (reference: src/scheduler/swap_scheduler.cpp:119-139).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace hiq_shim {
struct NullStream {
     template <class T>
     NullStream& operator<<(const T&) { return *this; }
     NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
     std::ostringstream s;
     ~FatalStream()
     {
          std::cerr << "CHECK failed: " << s.str() << std::endl;
          std::abort();
     }
     template <class T>
     FatalStream& operator<<(const T& v) { s << v; return *this; }
};
}  // namespace hiq_shim

#define VLOG(n) if (true) {} else hiq_shim::NullStream()
#define DLOG(x) if (true) {} else hiq_shim::NullStream()
#define LOG(x) if (true) {} else hiq_shim::NullStream()
#define CHECK(c) if (c) {} else hiq_shim::FatalStream()

namespace google {
inline void InitGoogleLogging(const char*) {}
inline void ShutdownGoogleLogging() {}
}  // namespace google
