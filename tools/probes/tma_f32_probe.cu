// Probe: fp32 TMA box loads (no swizzle). Usage: probe <bw> <dtype 0=f32 1=u32> <x> <y> <rank 2|3|4> <promo 0..3>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int n, int bytes, int x, int y) {
  __shared__ __align__(1024) float buf[4096];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(smem_u32(buf)), "l"(&tm), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(0), "r"(0) : "memory");
    else if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(buf)), "l"(&tm), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(0) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(buf)), "l"(&tm), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
  const int bw = atoi(argv[1]), dt = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]), rank = atoi(argv[5]), promo = atoi(argv[6]);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)sym;
  const int W = 512, H = 512, C = 3, B = 2;
  std::vector<float> h((size_t)W * H * C * B);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 4096 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t gd[4] = {W, H, C, B}, gs[3] = {W * 4ull, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  const int bc = rank >= 3 ? 3 : 1;
  cuuint32_t bx[4] = {(cuuint32_t)bw, 18, (cuuint32_t)bc, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = fn(&tm, dt ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int n = bw * 18 * bc;
  if (rank == 4) probe<4><<<1, 128>>>(tm, o, n, n * 4, x, y);
  else if (rank == 3) probe<3><<<1, 128>>>(tm, o, n, n * 4, x, y);
  else probe<2><<<1, 128>>>(tm, o, n, n * 4, x, y);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> g(n);
  cudaMemcpy(g.data(), o, n * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < bc; ++c) for (int yy = 0; yy < 18; ++yy) for (int xx = 0; xx < bw; ++xx) {
    const int Y = yy + y, X = xx + x;
    const float exp = (Y < 0 || X < 0) ? 0.f : h[((size_t)c * H + Y) * W + X];
    if (g[(c * 18 + yy) * bw + xx] != exp) ++bad;
  }
  printf("bw %d dt %d x %d y %d rank %d promo %d: encode %d sync %s bad %d\n", bw, dt, x, y, rank, promo, (int)r, cudaGetErrorString(e), bad);
  return 0;
}
