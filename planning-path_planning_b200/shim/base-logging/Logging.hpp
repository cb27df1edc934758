// Stand-in for base-logging: LOG_*_S stream macros.  Messages are dropped
// unless DYMU_SHIM_LOG_STDERR is defined.  Also pulls in the std headers and
// the `uint` typedef that the DyMu sources assume transitively.
#ifndef DYMU_SHIM_BASE_LOGGING_HPP
#define DYMU_SHIM_BASE_LOGGING_HPP
#include <sys/types.h>
#include <algorithm>
#include <cmath>
#include <iostream>
#include <limits>
#include <string>
namespace base_logging_shim
{
struct NullStream
{
    template <typename T> NullStream& operator<<(const T&) { return *this; }
    NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct LineStream
{
    ~LineStream() { std::cerr << std::endl; }
    template <typename T> LineStream& operator<<(const T& v)
    {
        std::cerr << v;
        return *this;
    }
};
}  // namespace base_logging_shim
#ifdef DYMU_SHIM_LOG_STDERR
#define DYMU_SHIM_LOG base_logging_shim::LineStream()
#else
#define DYMU_SHIM_LOG base_logging_shim::NullStream()
#endif
#define LOG_DEBUG_S DYMU_SHIM_LOG
#define LOG_INFO_S DYMU_SHIM_LOG
#define LOG_WARN_S DYMU_SHIM_LOG
#define LOG_ERROR_S DYMU_SHIM_LOG
#define LOG_FATAL_S DYMU_SHIM_LOG
#endif
