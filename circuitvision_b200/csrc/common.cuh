// Shared host-side plumbing for the C ABI: thread-local error string, launch counting, launch/check macros.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/cv_b200.h"

int cvb_fail(int code, const char* msg);
int cvb_fail_cuda(cudaError_t e, const char* what);
void cvb_reset_launches();
void cvb_count_launch();

#define CVB_CHECK(expr)                                          \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return cvb_fail_cuda(_e, #expr);      \
  } while (0)

// kernel launch + launch-error check + launch counter
#define CVB_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                     \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                    \
    cudaError_t _e = cudaGetLastError();                                   \
    if (_e != cudaSuccess) return cvb_fail_cuda(_e, "launch " #kernel);    \
    cvb_count_launch();                                                    \
  } while (0)
