// libf3d: error reporting and version of the C ABI declared in include/f3d.h.
#include <cstdio>
#include <cstring>

#include "../../include/f3d.h"
#include "f3d_host.h"

static thread_local char g_err[512] = "";

int f3d_fail(int code, const char* msg) {
    std::snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int f3d_check_launch(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
        return F3D_ERR_CUDA;
    }
    return F3D_OK;
}

extern "C" const char* f3d_last_error(void) { return g_err; }
extern "C" int f3d_version(void) { return 100; }

