// Host-side helpers shared by the libf3d translation units: thread-local error message and launch checking.
#pragma once
#include <cuda_runtime.h>

int f3d_fail(int code, const char* msg);
int f3d_check_launch(const char* where);
