// Shared by the C-ABI translation units: error reporting and allocation helpers.
#pragma once
#include <cstdio>
#include <string>

#include "sage_internal.cuh"

std::string& sage_err_slot();     // thread-local last-error message (defined in api.cu)

inline int sage_fail(const char* what, const char* file, int line, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    else
        snprintf(buf, sizeof buf, "%s (%s:%d)", what, file, line);
    sage_err_slot() = buf;
    return 1;
}

#define CK(call)                                                           \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return sage_fail(#call, __FILE__, __LINE__, e__); \
    } while (0)
#define REQUIRE(cond, msg)                                                 \
    do {                                                                   \
        if (!(cond)) return sage_fail(msg, __FILE__, __LINE__);            \
    } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t count)
{
    *p = nullptr;
    return cudaMalloc((void**)p, sizeof(T) * (count ? count : 1));
}
