// Ad-hoc: cost of shipping ~16 KB per launch as kernel parameters vs a stream-ordered H2D copy.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <chrono>
#include <vector>
namespace cg = cooperative_groups;

template<int N> struct Packet {alignas(16) unsigned char bytes[N];};

template<int N>
__global__ void k_param(const __grid_constant__ Packet<N> pkt, int n16, unsigned * out)
{
  extern __shared__ uint4 sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const uint4 * src = reinterpret_cast<const uint4 *>(pkt.bytes);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) {sm[i] = src[i];}
  __syncthreads();
  cluster.sync();
  if (threadIdx.x == 0 && out) {out[blockIdx.x] = sm[n16 - 1].x + sm[0].y;}
}
__global__ void k_glob(const uint4 * src, int n16, unsigned * out)
{
  extern __shared__ uint4 sm[];
  cg::cluster_group cluster = cg::this_cluster();
  for (int i = threadIdx.x; i < n16; i += blockDim.x) {sm[i] = src[i];}
  __syncthreads();
  cluster.sync();
  if (threadIdx.x == 0 && out) {out[blockIdx.x] = sm[n16 - 1].x + sm[0].y;}
}

template<typename F>
static void report(const char * name, F launch, cudaStream_t s, int n = 300)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  std::vector<float> dev, host, total;
  for (int i = 0; i < n; ++i) {
    const auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(a, s);
    launch();
    cudaEventRecord(b, s);
    const auto t1 = std::chrono::steady_clock::now();
    cudaStreamSynchronize(s);
    const auto t2 = std::chrono::steady_clock::now();
    float ms; cudaEventElapsedTime(&ms, a, b);
    dev.push_back(ms * 1000.f);
    host.push_back(std::chrono::duration<float, std::micro>(t1 - t0).count());
    total.push_back(std::chrono::duration<float, std::micro>(t2 - t0).count());
  }
  auto med = [](std::vector<float> v) {std::sort(v.begin(), v.end()); return v[v.size() / 2];};
  printf("%-44s device %.2f us  host enqueue %.2f us  host total %.2f us  (%s)\n", name, med(dev), med(host), med(total),
    cudaGetErrorString(cudaGetLastError()));
}

template<int N>
static void run_param(cudaStream_t s, unsigned * d_out, const char * name)
{
  static Packet<N> pkt;
  for (int i = 0; i < N; ++i) {pkt.bytes[i] = static_cast<unsigned char>(i);}
  cudaFuncSetAttribute(k_param<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  cudaFuncSetAttribute(k_param<N>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(16, 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = 170 * 1024; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  report(name, [&] {pkt.bytes[5]++; cudaLaunchKernelEx(&cfg, k_param<N>, pkt, N / 16, d_out);}, s);
}

int main()
{
  cudaStream_t s; cudaStreamCreate(&s);
  unsigned * d_out; cudaMalloc(&d_out, 64 * 4);
  run_param<4000>(s, d_out, "params 4000 B, cluster16 170KB smem");
  run_param<12 * 1024>(s, d_out, "params 12 KB");
  run_param<16 * 1024>(s, d_out, "params 16 KB");
  run_param<31 * 1024>(s, d_out, "params 31 KB");
  // the same bytes by a stream-ordered copy + kernel (graph, kernel only)
  {
    const int N = 12 * 1024;
    unsigned char * h; cudaMallocHost(&h, N); unsigned char * d; cudaMalloc(&d, N);
    cudaFuncSetAttribute(k_glob, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
    cudaFuncSetAttribute(k_glob, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(16, 1, 1); cfg.blockDim = dim3(512, 1, 1); cfg.dynamicSmemBytes = 170 * 1024; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    cudaLaunchKernelEx(&cfg, k_glob, reinterpret_cast<const uint4 *>(d), N / 16, d_out);
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    report("H2D 12 KB + graph[kernel]", [&] {h[5]++; cudaMemcpyAsync(d, h, N, cudaMemcpyHostToDevice, s); cudaGraphLaunch(ge, s);}, s);
    report("H2D 12 KB + direct kernel", [&] {h[5]++; cudaMemcpyAsync(d, h, N, cudaMemcpyHostToDevice, s); cudaLaunchKernelEx(&cfg, k_glob, reinterpret_cast<const uint4 *>(d), N / 16, d_out);}, s);
    report("graph[kernel] only", [&] {cudaGraphLaunch(ge, s);}, s);
    report("direct kernel only", [&] {cudaLaunchKernelEx(&cfg, k_glob, reinterpret_cast<const uint4 *>(d), N / 16, d_out);}, s);
  }
  return 0;
}
