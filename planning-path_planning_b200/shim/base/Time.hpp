// Stand-in for base/Time.hpp: now(), operator-, toSeconds().
#ifndef DYMU_SHIM_BASE_TIME_HPP
#define DYMU_SHIM_BASE_TIME_HPP
#include <chrono>
#include <cstdint>
namespace base
{
struct Time
{
    int64_t microseconds;
    Time() : microseconds(0) {}
    static Time fromMicroseconds(int64_t us)
    {
        Time t;
        t.microseconds = us;
        return t;
    }
    static Time now()
    {
        using namespace std::chrono;
        return fromMicroseconds(
            duration_cast<std::chrono::microseconds>(steady_clock::now().time_since_epoch())
                .count());
    }
    Time operator-(const Time& o) const { return fromMicroseconds(microseconds - o.microseconds); }
    double toSeconds() const { return static_cast<double>(microseconds) * 1e-6; }
};
}  // namespace base
#endif
