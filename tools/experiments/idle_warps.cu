// Ad-hoc: how 15 idle warps of a 512-thread block should wait while one warp walks a shared-memory chain.
// Measured on B200: waiting at __syncthreads() slows the working warp 2.5x.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void * p) {return static_cast<uint32_t>(__cvta_generic_to_shared(p));}

template<int MODE>  // 0 __syncthreads, 1 idle warps exit, 2 mbarrier try_wait, 3 nanosleep polling of a flag, 4 named barrier (bar.sync 1)
__global__ void k_prefix(float * out, long long * cyc, int n)
{
  extern __shared__ float sm[];
  __shared__ uint64_t bar;
  __shared__ volatile int flag;
  for (int i = threadIdx.x; i < n * 128; i += blockDim.x) {sm[i] = 1e-3f * i;}
  if (threadIdx.x == 0) {
    flag = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float * p = sm + threadIdx.x;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i + 8 <= n; i += 8, p += 8 * 128) {
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {d[j] = p[j * 128];}
#pragma unroll
      for (int j = 0; j < 8; ++j) {acc += d[j]; d[j] = acc;}
#pragma unroll
      for (int j = 0; j < 8; ++j) {p[j * 128] = d[j];}
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
      cyc[0] = t1 - t0;
      if (MODE == 2) {asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");}
      if (MODE == 3) {__threadfence_block(); flag = 1;}
    }
    out[threadIdx.x] = sm[threadIdx.x + (n - 1) * 128];
    if (MODE == 0) {__syncthreads();}
    if (MODE == 4) {asm volatile("bar.sync 1, 512;");}
  } else {
    if (MODE == 0) {__syncthreads();}
    if (MODE == 1) {return;}
    if (MODE == 2) {
      asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n"
        ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    if (MODE == 3) {while (flag == 0) {__nanosleep(200);}}
    if (MODE == 4) {asm volatile("bar.sync 1, 512;");}
    out[threadIdx.x] = 1.f;
  }
}

int main()
{
  float * out; long long * cyc; long long h;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
  const char * names[5] = {"idle warps at __syncthreads()", "idle warps exit", "idle warps in mbarrier.try_wait", "idle warps nanosleep-poll a flag", "idle warps at bar.sync 1"};
  for (int rep = 0; rep < 2; ++rep) {
    k_prefix<0><<<1, 512, 56 * 128 * 4>>>(out, cyc, 56); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-36s %5lld cycles for 56 steps\n", names[0], h);
    k_prefix<1><<<1, 512, 56 * 128 * 4>>>(out, cyc, 56); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-36s %5lld\n", names[1], h);
    k_prefix<2><<<1, 512, 56 * 128 * 4>>>(out, cyc, 56); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-36s %5lld\n", names[2], h);
    k_prefix<3><<<1, 512, 56 * 128 * 4>>>(out, cyc, 56); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-36s %5lld\n", names[3], h);
    k_prefix<4><<<1, 512, 56 * 128 * 4>>>(out, cyc, 56); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-36s %5lld\n", names[4], h);
    k_prefix<0><<<1, 32, 56 * 128 * 4>>>(out, cyc, 56); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-36s %5lld\n", "(32-thread block)", h);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
