// Stand-in for <cuda_runtime.h> when the csrc headers are compiled for the host by tests/emu (TEST INFRASTRUCTURE).
#pragma once
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
