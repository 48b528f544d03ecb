// Device build of libdtcwt_b200.so: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdio.h>

#include "generic_kernels.cuh"
#include "fused2d.cuh"
#include "fused3d.cuh"
#include "registration.cuh"
#include "keypoint.cuh"
#include "axis_pass.cuh"

namespace dtcwt {

// One thread per output element; gid spans blockIdx.x (2^31-1 blocks of 256 threads
// covers 5e11 elements).  Used by every generic kernel.
template <class Elem>
__global__ void __launch_bounds__(256) generic_1d_kernel(const typename Elem::Args a, const int64_t total) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < total) Elem::run(a, gid);
}

template <class Elem>
static int launch_1d(const typename Elem::Args& a, void* stream) {
    const int64_t total = Elem::total(a);
    if (total <= 0) return DTCWT_B200_OK;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return DTCWT_B200_EUNSUPPORTED;
    generic_1d_kernel<Elem><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(a, total);
    return (int)cudaGetLastError();
}

// the same with two CTAs per SM asked of the compiler (register-heavy float64 elements: QtildeElem took 171 registers)
template <class Elem>
__global__ void __launch_bounds__(256, 2) generic_1d_kernel_2(const typename Elem::Args a, const int64_t total) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < total) Elem::run(a, gid);
}

template <class Elem>
static int launch_1d_2(const typename Elem::Args& a, void* stream) {
    const int64_t total = Elem::total(a);
    if (total <= 0) return DTCWT_B200_OK;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return DTCWT_B200_EUNSUPPORTED;
    generic_1d_kernel_2<Elem><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, total);
    return (int)cudaGetLastError();
}

// Q~ of the registration (registration.cuh): two lanes per pixel, three sub-bands each; every thread of the grid takes part in
// the shuffles (lanes beyond the last pixel carry zeros)
template <class T>
__global__ void __launch_bounds__(256, 2) qtilde_kernel(const QtildeArgs<T> a) {
    QtildeElem<T>::run_lanes(a, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

template <class T>
static int launch_qtilde(const QtildeArgs<T>& a, void* stream) {
    const int64_t total = QtildeElem<T>::total(a);
    if (total <= 0) return DTCWT_B200_OK;
    const int64_t blocks = (2 * total + 255) / 256;
    if (blocks > 0x7fffffffLL) return DTCWT_B200_EUNSUPPORTED;
    qtilde_kernel<T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}

template <class K>
__global__ void __launch_bounds__(256) axis_kernel(const __grid_constant__ AxisArgs a, const int64_t total) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < total) K::run(a, gid);
}

template <class K>
static int launch_axis(const AxisArgs& a, void* stream) {
    const int64_t total = K::total(a);
    if (total <= 0) return DTCWT_B200_OK;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return DTCWT_B200_EUNSUPPORTED;
    axis_kernel<K><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, total);
    return (int)cudaGetLastError();
}

// depth passes of the fused 3-D levels (fused3d.cuh): one thread per (2 x 2 patch, depth group, image, volume)
template <class K>
__global__ void __launch_bounds__(256) z3_kernel(const __grid_constant__ Z3Args a, const int64_t total) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < total) K::run(a, gid);
}

template <class K>
static int launch_z3(const Z3Args& a, void* stream) {
    const int64_t total = K::total(a);
    if (total <= 0) return DTCWT_B200_OK;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return DTCWT_B200_EUNSUPPORTED;
    z3_kernel<K><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, total);
    return (int)cudaGetLastError();
}

}  // namespace dtcwt

#include "abi_generic.inl"
#include "abi_reg.inl"
#include "fused2d_launch.cuh"
#include "abi_fused2d.inl"
#include "abi_axis.inl"
#include "abi_fused3d.inl"
#include "abi_chain.inl"

#ifdef DTCWT_EMIT_GENERIC
extern "C" {

int dtcwt_b200_is_device_build(void) { return 1; }

const char* dtcwt_b200_error_string(int code) {
    if (code == DTCWT_B200_OK) return "ok";
    if (code == DTCWT_B200_EINVAL) return "dtcwt_b200: invalid argument (shape, tap count or NULL pointer)";
    if (code == DTCWT_B200_EUNSUPPORTED) return "dtcwt_b200: request not supported by this build";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "dtcwt_b200: unknown error";
}

}  // extern "C"
#endif  // DTCWT_EMIT_GENERIC
