// Exception barrier of the C ABI: every `extern "C" int egx_*` entry point is a function-try-block ending in
// EGX_ABI_CATCH, so that a C++ exception (std::bad_alloc from a host vector, std::length_error from an absurd size)
// never unwinds into a C / Rust / Go / JVM caller: it becomes EGX_CUDA_ERROR ("library failure") with the message in
// egx_last_error().
#pragma once
#include <exception>

#include "../../include/egobox_gpu.h"

void egx_set_error(const char* fmt, ...);

#define EGX_ABI_CATCH                                                  \
    catch (const std::exception& e__) {                                \
        egx_set_error("internal error: %s", e__.what());               \
        return EGX_CUDA_ERROR;                                         \
    }                                                                  \
    catch (...) {                                                      \
        egx_set_error("internal error: unknown C++ exception");        \
        return EGX_CUDA_ERROR;                                         \
    }

// for the two entry points whose int is a COUNT, not a status (egx_device_count, egx_gp_async_slots): 0 on any exception
#define EGX_ABI_CATCH_COUNT \
    catch (...) {           \
        return 0;           \
    }
