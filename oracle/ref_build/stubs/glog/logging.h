// Stand-in for <glog/logging.h>: LOG(...) << ... compiles to nothing and never evaluates its operands (glog itself skips
// them when a severity is disabled), which also keeps the reference's `imuMeasuementBuffer_.end()->timeStamp` log
// arguments (filter.cpp:53,204) from being read.
#ifndef FBUS_REF_STUB_GLOG
#define FBUS_REF_STUB_GLOG
#include <ostream>
namespace fbus_ref_stub {
struct NullStream {
    template <class T> NullStream& operator<<(const T&) { return *this; }
    NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
    NullStream& operator<<(std::ios_base& (*)(std::ios_base&)) { return *this; }
};
struct Voidify {
    void operator&(NullStream&) {}
};
}  // namespace fbus_ref_stub
#define LOG(severity) true ? (void)0 : ::fbus_ref_stub::Voidify() & ::fbus_ref_stub::NullStream()
#define LOG_EVERY_N(severity, n) LOG(severity)
#define LOG_IF(severity, cond) LOG(severity)
#define VLOG(n) LOG(INFO)
#endif
