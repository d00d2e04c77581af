// Platform layer of libsdtree.
//
// The product is built by nvcc for sm_100a (see build.py).  The SAME sources also
// compile with plain g++ when SDT_HOSTEMU is defined: every "kernel" then runs as a
// serial loop and the CUDA runtime calls below become malloc/memcpy.  That build
// (tests/hostemu, libsdtree_hostemu.so) exists ONLY so that the index logic of the
// kernels can be checked against the oracle in a container without a GPU; it is test
// infrastructure, it is never loaded by the package (practical_path_guiding_lab_b200/
// _lib.py loads libsdtree.so and nothing else) and nothing is measured through it.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef SDT_HOSTEMU
#include <cuda_runtime.h>
#define SDT_HD __host__ __device__ __forceinline__
template <class T>
__host__ __device__ __forceinline__ T sdt_ldg_(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
#define SDT_LDG(p) sdt_ldg_(p)
#else
#include <math.h>
#define SDT_HD inline
#define SDT_LDG(p) (*(p))
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
static inline const char* cudaGetErrorString(cudaError_t) { return "hostemu"; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return 0; }
#endif
