// Shared host-side plumbing for the C ABI: thread-local error string, launch counting, optional per-kernel
// CUDA-event timing (cv_profile_*), launch/check macros.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/cv_b200.h"

int cvb_fail(int code, const char* msg);
int cvb_fail_cuda(cudaError_t e, const char* what);
void cvb_reset_launches();
void cvb_count_launch();
// Per-kernel timing: when enabled, every CVB_LAUNCH is bracketed by two events on its stream.
bool cvb_profile_on();
void cvb_profile_begin(const char* name, cudaStream_t st, double work);
void cvb_profile_end(cudaStream_t st);
// algorithmic work (bytes or flops) of the NEXT launch, consumed by cvb_profile_begin
void cvb_next_work(double w);
double cvb_take_work();
// optional display name of the NEXT launch in the profile table (e.g. a GEMM's shape); consumed by cvb_profile_begin
void cvb_next_name(const char* name);

#define CVB_CHECK(expr)                                          \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return cvb_fail_cuda(_e, #expr);      \
  } while (0)

// kernel launch + launch-error check + launch counter (+ events when profiling)
#define CVB_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                     \
    bool _p = cvb_profile_on();                                            \
    if (_p) cvb_profile_begin(#kernel, stream, cvb_take_work());           \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                    \
    cudaError_t _e = cudaGetLastError();                                   \
    if (_p) cvb_profile_end(stream);                                       \
    if (_e != cudaSuccess) return cvb_fail_cuda(_e, "launch " #kernel);    \
    cvb_count_launch();                                                    \
  } while (0)

// Function attributes (dynamic shared-memory size, carve-out) are per DEVICE: a process that drives several GPUs must set
// them once on each.  `mask` is a function-local static; returns true the first time it is called on the current device.
#include <atomic>
inline bool cvb_once_per_device(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  return (mask.fetch_or(bit) & bit) == 0;
}

