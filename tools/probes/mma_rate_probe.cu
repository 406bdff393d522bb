// Micro-benchmark: clocks per tcgen05.mma (kind::f16, M = 128 per CTA, K = 16, operands in shared memory) as a function
// of N, of the A-operand layout (dense 16 KB tile vs the halo-tile addressing of the conv kernels: pixel rows 128 B apart,
// 8-row groups one halo pitch apart) and of cta_group (1: one SM; 2: CTA pair, M = 256, each CTA holds half of B).
// It answers the question DESIGN.md section 11 item 1 asks: are the N = 128 SS-MMAs of the halo kernels bound by
// shared-memory operand bandwidth (then cta_group::2 lifts them), or by something else.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mma_rate_probe.bin tools/probes/mma_rate_probe.cu
// All CTAs of the grid issue concurrently; operand values are zeros (only the issue rate is measured).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct ProbeArgs {
  int N;            // MMA N (whole pair for cta_group::2)
  int iters;        // x 12 MMAs
  uint32_t a_sbo;   // 8-row-group stride of the A operand (1024 dense; 2560 = pitch-10 halo, two image rows apart)
  uint32_t a_tap;   // byte step between the three "taps" (16384 dense tiles; 128 = next halo pixel)
  unsigned long long* out;
};

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(const ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;                 // 64 KB
  uint8_t* smem_b = smem + 65536;         // 3 x 32 KB
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

  for (int i = threadIdx.x; i < (65536 + 3 * 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_smem;

  if (warp == 1 && lane == 0) {
    if (cta_rank == 0) {
      const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
      const uint32_t idesc = idesc_bf16(CG == 2 ? 256 : 128, p.N);
      const uint32_t b_tile = (uint32_t)(p.N / CG) * 128u;   // bytes of one K = 64 tile of this CTA's share of B
      const long long t0 = clock64();
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = desc_sw128(a_base + t * p.a_tap + k * 32, p.a_sbo);
            const uint64_t bd = desc_sw128(b_base + t * b_tile + k * 32, 1024);
            if (CG == 1)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                           "l"(ad), "l"(bd), "r"(idesc), "r"(1)
                           : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                           "l"(ad), "l"(bd), "r"(idesc), "r"(1)
                           : "memory");
          }
        }
      }
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                     "h"((uint16_t)3)
                     : "memory");
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      p.out[blockIdx.x / CG] = (unsigned long long)(t1 - t0);
    } else {
      mbar_wait(&bar, 0);   // the peer's shared memory and TMEM are in use until the leader's MMAs have retired
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
  }
}

template <int CG>
static void run(const char* label, int N, uint32_t a_sbo, uint32_t a_tap, int iters, unsigned long long* d_out, int n_sm) {
  const int smem_bytes = 65536 + 3 * 32768 + 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  ProbeArgs a{N, iters, a_sbo, a_tap, d_out};
  const int grid = (n_sm / CG) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {   // second launch is the measurement
    CK(cudaMemset(d_out, 0, sizeof(unsigned long long) * n_sm));
    CK(cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG>, a));
    CK(cudaDeviceSynchronize());
  }
  unsigned long long h[256];
  CK(cudaMemcpy(h, d_out, sizeof(unsigned long long) * (grid / CG), cudaMemcpyDeviceToHost));
  double sum = 0, mx = 0, mn = 1e30;
  for (int i = 0; i < grid / CG; ++i) { sum += (double)h[i]; if (h[i] > mx) mx = (double)h[i]; if (h[i] < mn) mn = (double)h[i]; }
  const double n_mma = 12.0 * iters;
  const double per = sum / (grid / CG) / n_mma;
  // MACs per SM per clock: M(128 per CTA) x N x 16 per instruction per CTA
  printf("%-34s cta_group::%d N=%3d  clocks/MMA avg %.1f (min %.1f max %.1f)  -> %.0f MAC/clk/SM\n", label, CG, N, per, mn / n_mma, mx / n_mma,
         128.0 * N * 16.0 / per);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 1500;
  const bool with_cg2 = argc > 2 ? atoi(argv[2]) != 0 : true;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int n_sm = prop.multiProcessorCount;
  printf("%s, %d SMs, %d x 12 MMAs per CTA\n", prop.name, n_sm, iters);
  unsigned long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(unsigned long long) * 256));
  for (int N : {64, 128, 256}) run<1>("dense A tiles (SBO 1024)", N, 1024, 16384, iters, d_out, n_sm);
  for (int N : {64, 128, 256}) run<1>("halo A (pixel taps, SBO 2560)", N, 2560, 128, iters, d_out, n_sm);
  if (with_cg2) {
    for (int N : {128, 256}) run<2>("dense A tiles (SBO 1024)", N, 1024, 16384, iters, d_out, n_sm);
    for (int N : {128, 256}) run<2>("halo A (pixel taps, SBO 2560)", N, 2560, 128, iters, d_out, n_sm);
  }
  return 0;
}
