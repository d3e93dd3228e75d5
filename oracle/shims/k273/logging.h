/* ORACLE shim for k273/logging.h: printf-style loggers, silent unless GZREF_VERBOSE is set. */
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace K273 {
inline bool log_enabled() {
    static int on = -1;
    if (on < 0) on = std::getenv("GZREF_VERBOSE") ? 1 : 0;
    return on == 1;
}
#define GZREF_LOGFN(name)                                   \
    inline void name(const char* fmt, ...) {                \
        if (!log_enabled()) return;                         \
        va_list ap; va_start(ap, fmt);                      \
        std::vfprintf(stderr, fmt, ap); va_end(ap);         \
        std::fputc('\n', stderr);                           \
    }                                                       \
    inline void name(const std::string& s) { if (log_enabled()) std::fprintf(stderr, "%s\n", s.c_str()); }
GZREF_LOGFN(l_debug)
GZREF_LOGFN(l_verbose)
GZREF_LOGFN(l_info)
GZREF_LOGFN(l_warning)
GZREF_LOGFN(l_error)
GZREF_LOGFN(l_critical)
}  // namespace K273
