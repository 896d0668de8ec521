// The exception barrier of the C ABI (egobox_b200/csrc/abi_guard.h) on a function that throws: the caller sees a status
// and a message, never an exception.  Built and run by tests/test_cabi_exports.py.
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>

#include "abi_guard.h"

extern "C" int thrower(int which) try {
    if (which == 0) throw std::bad_alloc();
    if (which == 1) throw std::length_error("vector::_M_default_append");
    if (which == 2) throw 42;
    return EGX_OK;
}
EGX_ABI_CATCH

int main() {
    int bad = 0;
    bad += !(thrower(0) == EGX_CUDA_ERROR && std::strstr(egx_last_error(), "bad_alloc") != nullptr);
    bad += !(thrower(1) == EGX_CUDA_ERROR && std::strstr(egx_last_error(), "_M_default_append") != nullptr);
    bad += !(thrower(2) == EGX_CUDA_ERROR && std::strstr(egx_last_error(), "unknown C++ exception") != nullptr);
    bad += !(thrower(3) == EGX_OK);
    if (bad == 0) std::printf("abi guard ok\n");
    return bad;
}
