// Block-level reduction of per-thread 8-channel partial sums for the NHWC reduction kernels. The block is laid out
// as `cg` channel groups x `lanes` pixel lanes (threadIdx.x = lane * cg + g); every thread calls block_channel_sum
// (idle lanes pass zeros). On return, thread t < min(blockDim, cg*8) holds the block total of channel
// t, t + blockDim, ... in out[]: one atomic per channel per block instead of one per thread.
#pragma once
#include <cuda_runtime.h>

namespace pmoe {

constexpr int kRedThreads = 256;

// sm: kRedThreads * 8 floats. Calls __syncthreads twice. For channel index c = threadIdx.x + j * blockDim.x
// (j < kRedMaxIter), the total of channel c is returned in out[j] (only meaningful when c < cg * 8).
constexpr int kRedMaxIter = 8;  // cg * 8 <= 2048 channels
__device__ __forceinline__ void block_channel_sum(const float (&v)[8], float* sm, int cg, int lanes, float (&out)[kRedMaxIter]) {
  float4* s4 = reinterpret_cast<float4*>(sm) + threadIdx.x * 2;
  s4[0] = make_float4(v[0], v[1], v[2], v[3]);
  s4[1] = make_float4(v[4], v[5], v[6], v[7]);
  __syncthreads();
  const int nch = cg * 8;
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    float t = 0.f;
    if (c < nch) {
      for (int l = 0; l < lanes; ++l) t += sm[l * nch + c];
    }
    out[j] = t;
  }
  __syncthreads();
}

}  // namespace pmoe
