// windows.h shim for building the UNTOUCHED reference headers with g++ on Linux.
// Camera.txt:17 includes "windows.h"; under MSVC that header (and MSVC's transitive
// includes) provided std::thread / std::chrono / std::atomic, the `max` macro used at
// Camera.txt:249, and sprintf_s used by the hand-patched stb_image_write.h:778.
// Test infrastructure only (oracle/); never part of the product.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <thread>
#include <algorithm>
using std::max;
using std::min;
template <size_t N>
inline int sprintf_s(char (&buf)[N], const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int r = vsnprintf(buf, N, fmt, ap);
    va_end(ap);
    return r;
}
