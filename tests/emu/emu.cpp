// Host-side kernel-logic emulator -- TEST INFRASTRUCTURE, never loaded by the package.
//
// The build container has nvcc but no GPU.  This file compiles the SAME kernel
// bodies (dtcwt_b200/csrc/*.cuh, DTCWT_EMU) and the SAME C-ABI wrappers
// (abi_generic.inl, ...) with g++ and runs every "thread" in a loop, so the index
// arithmetic of the CUDA kernels and the whole Python host layer can be exercised
// by `pytest -m "not gpu"`.  "Device" pointers are host pointers here; `stream`
// is ignored.  dtcwt_b200_is_device_build() returns 0 so nothing can mistake this
// for the product library.
#define DTCWT_EMU 1
#include <stdio.h>
#include <string.h>

#include "../../dtcwt_b200/csrc/generic_kernels.cuh"

namespace dtcwt {

template <class Elem>
static int launch_1d(const typename Elem::Args& a, void* /*stream*/) {
    const int64_t total = Elem::total(a);
    for (int64_t gid = 0; gid < total; ++gid) Elem::run(a, gid);
    return DTCWT_B200_OK;
}

}  // namespace dtcwt

#include "../../dtcwt_b200/csrc/abi_generic.inl"

extern "C" {

int dtcwt_b200_is_device_build(void) { return 0; }

const char* dtcwt_b200_error_string(int code) {
    if (code == DTCWT_B200_OK) return "ok";
    if (code == DTCWT_B200_EINVAL) return "dtcwt_b200: invalid argument (shape, tap count or NULL pointer)";
    if (code == DTCWT_B200_EUNSUPPORTED) return "dtcwt_b200: request not supported by this build";
    return "dtcwt_b200[emulator]: unknown error";
}

}  // extern "C"
