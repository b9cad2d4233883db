// Ad-hoc: event-to-event time of (almost) empty launches in the shapes the fused kernel uses.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>
namespace cg = cooperative_groups;

__global__ void k_null(int * out) {if (out && threadIdx.x == 9999) {*out = 1;}}
__global__ void k_cluster(int * out)
{
  extern __shared__ unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  if (threadIdx.x == 0) {smem[0] = 1;}
  cluster.sync();
  if (out && threadIdx.x == 9999) {*out = smem[0];}
}
__global__ void k_spin(long long cycles)
{
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}

template<typename F>
static float p50(F launch, cudaStream_t s, int n = 200)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  std::vector<float> v;
  for (int i = 0; i < n; ++i) {
    cudaEventRecord(a, s);
    launch();
    cudaEventRecord(b, s);
    cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, a, b);
    v.push_back(ms * 1000.f);
  }
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}

int main()
{
  cudaStream_t s; cudaStreamCreate(&s);
  int * d; cudaMalloc(&d, 4);
  printf("null <<<1,32>>>                      : %.2f us\n", p50([&] {k_null<<<1, 32, 0, s>>>(d);}, s));
  printf("null <<<16,512>>>                    : %.2f us\n", p50([&] {k_null<<<16, 512, 0, s>>>(d);}, s));
  printf("spin 20us <<<16,512>>>               : %.2f us\n", p50([&] {k_spin<<<16, 512, 0, s>>>(40000);}, s));
  for (int smem : {1024, 100 * 1024, 170 * 1024, 220 * 1024}) {
    for (int cs : {1, 8, 16}) {
      cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(k_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(16, 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      float t = p50([&] {cudaLaunchKernelEx(&cfg, k_cluster, d);}, s);
      printf("cluster %2d x 512 thr, %3d KB smem     : %.2f us  (%s)\n", cs, smem / 1024, t, cudaGetErrorString(cudaGetLastError()));
    }
  }
  // graph: memcpy node + kernel vs kernel only
  {
    int * h; cudaMallocHost(&h, 4096); int * dd; cudaMalloc(&dd, 4096);
    cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(16, 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = 170 * 1024; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int with_copy = 0; with_copy < 2; ++with_copy) {
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
      if (with_copy) {cudaMemcpyAsync(dd, h, 4096, cudaMemcpyHostToDevice, s);}
      cudaLaunchKernelEx(&cfg, k_cluster, d);
      cudaStreamEndCapture(s, &g);
      cudaGraphInstantiate(&ge, g, 0);
      printf("graph [%s cluster16 170KB]       : %.2f us\n", with_copy ? "H2D 4KB +" : "         ", p50([&] {cudaGraphLaunch(ge, s);}, s));
    }
    printf("stream H2D 4KB + cluster16 170KB      : %.2f us\n", p50([&] {cudaMemcpyAsync(dd, h, 4096, cudaMemcpyHostToDevice, s); cudaLaunchKernelEx(&cfg, k_cluster, d);}, s));
    printf("stream H2D 4KB only                   : %.2f us\n", p50([&] {cudaMemcpyAsync(dd, h, 4096, cudaMemcpyHostToDevice, s);}, s));
  }
  // spin kernel of known duration inside events: overhead = measured - 20 us
  printf("spin 40000 cyc <<<16,512>>> again     : %.2f us\n", p50([&] {k_spin<<<16, 512, 0, s>>>(40000);}, s));
  return 0;
}
