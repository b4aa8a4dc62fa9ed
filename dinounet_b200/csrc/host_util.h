// Host-side helpers shared by the C-ABI translation units: error reporting, launch accounting, TMA descriptor encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2u {

int set_error(int code, const char* fmt, ...);
// Index of the current CUDA device (0..63): cudaFuncSetAttribute is PER DEVICE, so "already configured" flags are arrays.
int current_device_index();
// cudaGetLastError() after a launch; bumps the launch counter on success.
int check_launch(const char* what);
// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time libcuda dependency): 16-bit elements,
// 128B swizzle, zero OOB fill.  `strides` are byte strides of dims 1..rank-1.
int encode_tensor_map(CUtensorMap* map, int dtype, int rank, const void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* elem_strides);

}  // namespace b2u
