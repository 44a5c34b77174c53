// libeod_memory.so: error reporting and device queries shared by all entry points.
#include <stdarg.h>
#include <string.h>

#include "eod_common.cuh"

static thread_local char g_err[512] = "";

void eod_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int eod_check_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        eod_set_error("%s: %s", what, cudaGetErrorString(e));
        return EOD_ERR_LAUNCH;
    }
    return EOD_OK;
}

int eod_num_sms()
{
    static int cache[64] = {0};                 // per device ordinal (a process may drive several GPUs)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int &n = cache[dev & 63];
    if (!n) {
        int v = 0;
        n = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : EOD_NUM_SMS_FALLBACK;
    }
    return n;
}

extern "C" int eod_version(void) { return 100; /* 0.1.0 */ }
extern "C" const char *eod_last_error(void) { return g_err; }
