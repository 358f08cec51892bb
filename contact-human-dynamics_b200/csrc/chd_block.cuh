// Block-wide reductions (product code, device only).
#pragma once
// ------------------------------------------------------------------ block helpers ----------------
__device__ __forceinline__ double chd_block_sum(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  double r = red[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double chd_block_max(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + off]);
    __syncthreads();
  }
  double r = red[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double chd_block_min(double v, double* red) { return -chd_block_max(-v, red); }

