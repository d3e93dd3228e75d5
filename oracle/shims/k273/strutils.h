/* ORACLE shim for k273/strutils.h */
#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>

namespace K273 {
inline std::string fmtString(const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt);
    std::vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return std::string(buf);
}
}  // namespace K273
