// Stand-in for <boost/format.hpp> (test infrastructure; Boost is not
// installed). Supports the printf-style subset the reference uses to build
// log lines and exception messages (e.g. src/simulator-mpi/SimulatorMPI.cpp:204-208):
// every %-directive is replaced, in order, by the next operand streamed with
// operator<<.
#pragma once
#include <sstream>
#include <string>
#include <vector>

namespace boost {
class format {
public:
     format(const char* f) : fmt_(f) {}
     format(const std::string& f) : fmt_(f) {}
     template <class T>
     format& operator%(const T& v)
     {
          std::ostringstream o;
          o << v;
          args_.push_back(o.str());
          return *this;
     }
     std::string str() const
     {
          std::string out;
          size_t a = 0;
          for (size_t i = 0; i < fmt_.size(); ++i) {
               if (fmt_[i] != '%') { out += fmt_[i]; continue; }
               if (i + 1 < fmt_.size() && fmt_[i + 1] == '%') { out += '%'; ++i; continue; }
               size_t j = i + 1;
               while (j < fmt_.size() && std::string("diouxXeEfgGscpl.0123456789+- #").find(fmt_[j]) != std::string::npos) {
                    char c = fmt_[j];
                    ++j;
                    if (std::string("diouxXeEfgGscp").find(c) != std::string::npos) break;
               }
               out += a < args_.size() ? args_[a] : std::string("?");
               ++a;
               i = j - 1;
          }
          return out;
     }
private:
     std::string fmt_;
     std::vector<std::string> args_;
};
inline std::ostream& operator<<(std::ostream& o, const format& f) { return o << f.str(); }
}  // namespace boost
