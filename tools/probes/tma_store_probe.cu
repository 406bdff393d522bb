// Probe: strided (elementStrides 2,2) TMA tensor STORE of a SWIZZLE_64B staged bf16 tile. Usage: probe <px> <py> <swz 0|1>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int px, int py, int swz) {
  __shared__ __align__(1024) __nv_bfloat16 buf[4 * 8 * 32];
  const int lane = threadIdx.x;  // 32 threads: lane = row*8 + col
  for (int k = 0; k < 4; ++k) {  // 16-byte piece k of this lane's 64-byte pixel
    const int slot = swz ? (k ^ ((lane >> 1) & 3)) : k;
    __nv_bfloat16* dst = buf + lane * 32 + slot * 8;
    for (int j = 0; j < 8; ++j) dst[j] = __float2bfloat16((float)(lane * 32 + k * 8 + j));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(&tm), "r"(smem_u32(buf)), "r"(0), "r"(px), "r"(py), "r"(0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main(int argc, char** argv) {
  const int px = atoi(argv[1]), py = atoi(argv[2]), swz = atoi(argv[3]);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)sym;
  const int W2 = 32, H2 = 16, C = 32;
  __nv_bfloat16* d; cudaMalloc(&d, (size_t)W2 * H2 * C * 2); cudaMemset(d, 0, (size_t)W2 * H2 * C * 2);
  CUtensorMap tm;
  cuuint64_t gd[4] = {C, W2, H2, 1}, gs[3] = {C * 2ull, (cuuint64_t)W2 * C * 2, (cuuint64_t)W2 * H2 * C * 2};
  cuuint32_t bx[4] = {32, 16, 8, 1}, es[4] = {1, 2, 2, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  probe<<<1, 32>>>(tm, px, py, swz);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<__nv_bfloat16> h((size_t)W2 * H2 * C);
  cudaMemcpy(h.data(), d, h.size() * 2, cudaMemcpyDeviceToHost);
  int bad = 0, nonzero = 0;
  for (int y = 0; y < H2; ++y) for (int x = 0; x < W2; ++x) for (int c = 0; c < C; ++c) {
    float exp = 0.f;
    const int ry = y - py, rx = x - px;
    if (ry >= 0 && rx >= 0 && ry % 2 == 0 && rx % 2 == 0 && ry / 2 < 4 && rx / 2 < 8) exp = (float)(((ry / 2) * 8 + rx / 2) * 32 + c);
    const float got = __bfloat162float(h[((size_t)y * W2 + x) * C + c]);
    if (got != 0.f) ++nonzero;
    if (got != __bfloat162float(__float2bfloat16(exp))) ++bad;
  }
  printf("px %d py %d swz %d: encode %d sync %s nonzero %d bad %d\n", px, py, swz, (int)r, cudaGetErrorString(e), nonzero, bad);
  return 0;
}
