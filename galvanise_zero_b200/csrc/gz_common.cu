#include "gz_common.cuh"

static thread_local char g_err[512] = "";

void gz_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gz_last_error(void) { return g_err; }
