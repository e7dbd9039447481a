// jinc_error.cpp -- thread-local error text behind jinc_last_error().
#include <cstdio>

#include "jinc_internal.h"

namespace {
thread_local char g_err[512] = "";
}

void jinc_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int jinc_fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* jinc_last_error(void) { return g_err; }
extern "C" int jinc_abi_version(void) { return JINC_ABI_VERSION; }
