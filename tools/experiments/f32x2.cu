// Does packed FP32x2 arithmetic (FADD2 / FMUL2 / FFMA2, sm_100) halve the issue slots of an issue-bound FP32 loop?
// Each thread advances two independent (a, b, c) recurrences: once with scalar instructions (6 per iteration), once
// with packed ones (3 per iteration), at 32 and at 16 resident warps per SM.  nvcc -arch=sm_100a -fmad=false.
#include <cstdio>
#include <cuda_runtime.h>

template<bool PACKED, int NINT>
__global__ void k(float * out, int iters)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float2 a = make_float2(1.0f + i * 1e-6f, 1.0f - i * 1e-6f), b = make_float2(0.5f, 0.25f), c = make_float2(1e-3f, 2e-3f);
  const float2 s = make_float2(0.999f, 1.001f);
  unsigned h[8] = {1u + i, 2u + i, 3u, 4u, 5u, 6u, 7u, 8u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < NINT; ++q) {h[q] = (h[q] ^ (h[q] >> 3)) + 0x9e3779b9u;}   // two ALU-pipe instructions each, independent chains
    if (PACKED) {
      a = __ffma2_rn(a, s, c);
      b = __fadd2_rn(b, a);
      c = __fmul2_rn(c, s);
    } else {
      a.x = fmaf(a.x, s.x, c.x); a.y = fmaf(a.y, s.y, c.y);
      b.x = b.x + a.x; b.y = b.y + a.y;
      c.x = c.x * s.x; c.y = c.y * s.y;
    }
  }
  unsigned hs = 0;
  for (int q = 0; q < 8; ++q) {hs += h[q];}
  out[i] = a.x + a.y + b.x + b.y + c.x + c.y + static_cast<float>(hs);
}

int main()
{
  float * out;
  cudaMalloc(&out, sizeof(float) * 148 * 2048);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 16;
  for (int warps = 16; warps <= 32; warps *= 2) {
    for (int packed = 0; packed < 4; ++packed) {
      const dim3 grid(148 * warps / 8), block(256);
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (packed == 0) {k<false, 0><<<grid, block>>>(out, iters);} else if (packed == 1) {k<true, 0><<<grid, block>>>(out, iters);}
        else if (packed == 2) {k<false, 3><<<grid, block>>>(out, iters);} else {k<true, 3><<<grid, block>>>(out, iters);}
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      const double flop_pairs = 3.0 * iters * 148.0 * warps * 32;   // (a, b, c) updates of two recurrences per thread
      printf("%2d warps/SM %s%s: %.3f ms, %.2f T pair-updates/s\n", warps, (packed & 1) ? "packed" : "scalar", packed >= 2 ? " + 6 integer instructions per iteration" : "", best, flop_pairs / best * 1e-9);
    }
  }
  return 0;
}
