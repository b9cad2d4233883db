// What does each way of synchronising a thread-block cluster compile to on sm_100a?  Not run: compile and read the SASS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o cb.cubin cluster_barrier_sass.cu
//   cuobjdump -sass cb.cubin | grep -E 'Function|MEMBAR|UCGABAR|CCTL|ERRBAR|SYNCS|STAS'
// v_sync / v_fence_relaxed / v_mbar: MEMBAR.ALL.GPU on every warp; v_relaxed: none (and no ordering); v_async (st.async +
// complete_tx on the receiver's mbarrier, what k_fused_cluster uses): STAS, no fence.
#include <cooperative_groups.h>
#include <cstdint>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t smem_u32(const void * p){return static_cast<uint32_t>(__cvta_generic_to_shared(p));}

extern "C" __global__ void __cluster_dims__(4,1,1) v_sync(float * out){
  __shared__ float x; cg::cluster_group c = cg::this_cluster();
  if(threadIdx.x==0) x = out[blockIdx.x];
  c.sync();
  out[blockIdx.x+100] = *c.map_shared_rank(&x,(c.block_rank()+1)%4);
}
extern "C" __global__ void __cluster_dims__(4,1,1) v_relaxed(float * out){
  __shared__ float x; cg::cluster_group c = cg::this_cluster();
  if(threadIdx.x==0) x = out[blockIdx.x];
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  out[blockIdx.x+100] = *c.map_shared_rank(&x,(c.block_rank()+1)%4);
}
extern "C" __global__ void __cluster_dims__(4,1,1) v_fence_relaxed(float * out){
  __shared__ float x; cg::cluster_group c = cg::this_cluster();
  if(threadIdx.x==0) x = out[blockIdx.x];
  asm volatile("fence.acq_rel.cluster;\n\tbarrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  out[blockIdx.x+100] = *c.map_shared_rank(&x,(c.block_rank()+1)%4);
}
// mbarrier all-to-all: every CTA's thread q arrives on rank q's barrier
extern "C" __global__ void __cluster_dims__(4,1,1) v_mbar(float * out){
  __shared__ float x; __shared__ __align__(8) uint64_t bar; cg::cluster_group c = cg::this_cluster();
  if(threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(4)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  c.sync();
  if(threadIdx.x==0) x = out[blockIdx.x];
  __syncthreads();
  if(threadIdx.x<4){
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&bar)), "r"(threadIdx.x));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
  }
  {
    uint32_t done=0;
    while(!done){
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  out[blockIdx.x+100] = *c.map_shared_rank(&x,(c.block_rank()+1)%4);
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank){uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;}
__device__ __forceinline__ void st_async_b32(uint32_t dst, uint32_t v, uint32_t bar){asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" :: "r"(dst), "r"(v), "r"(bar) : "memory");}
__device__ __forceinline__ void st_async_v4(uint32_t dst, int4 v, uint32_t bar){asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" :: "r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t * bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAITC_LOOP:\n"
    "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
    "@p bra WAITC_DONE;\n"
    "bra WAITC_LOOP;\n"
    "WAITC_DONE:\n"
    "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
extern "C" __global__ void __cluster_dims__(4,1,1) v_async(float * out){
  __shared__ __align__(16) int4 tab[16]; __shared__ float tab2[16]; __shared__ __align__(8) uint64_t bar[2]; cg::cluster_group c = cg::this_cluster();
  const int rank = c.block_rank(), n = c.num_blocks();
  if(threadIdx.x==0){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[1])), "r"(1));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(16*n) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(4*n) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  float v = out[blockIdx.x*blockDim.x+threadIdx.x];
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  if(threadIdx.x < n){
    st_async_v4(mapa_u32(smem_u32(&tab[rank]), threadIdx.x), make_int4(rank, 1, 2, __float_as_int(v)), mapa_u32(smem_u32(&bar[0]), threadIdx.x));
    st_async_b32(mapa_u32(smem_u32(&tab2[rank]), threadIdx.x), __float_as_uint(v), mapa_u32(smem_u32(&bar[1]), threadIdx.x));
  }
  mbar_wait_cluster(&bar[0], 0);
  mbar_wait_cluster(&bar[1], 0);
  out[blockIdx.x+100] = tab[threadIdx.x & 3].x + tab2[threadIdx.x&3];
}
