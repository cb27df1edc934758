// DyMuTiming.hpp -- DYMU_TIMING=1: wall time per stage of the host layer on stderr (developer aid).
#ifndef DYMU_TIMING_HPP
#define DYMU_TIMING_HPP
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

namespace
{
struct Lap
{
    const char* name;
    double t0;
    static bool on()
    {
        static int v = -1;
        if (v < 0) v = getenv("DYMU_TIMING") ? 1 : 0;
        return v == 1;
    }
    static double now()
    {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit Lap(const char* n) : name(n), t0(on() ? now() : 0.0) {}
    ~Lap()
    {
        if (on()) fprintf(stderr, "[dymu timing] %-28s %8.3f ms\n", name, now() - t0);
    }
};

}  // namespace
#endif
