// Experiment (not product code): MN-major, NON-swizzled UMMA A-descriptor over the same strip:
//   out[m][n] = sum_p x8view[p][m] * G[p][n],  x8view[p][m] = strip[16*p + 2*m]  (m < 64)
// core matrix (8 positions x 8 m): rows 16 B apart, contiguous 128 B at 16*(p0 + mb); next along
// MN +16 B, next along K (8 positions) +128 B.  Which of LBO / SBO is which is what is tested.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../bpreveal_b200/csrc/common.cuh"
#include "../../bpreveal_b200/csrc/ptx.cuh"

using namespace bp;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__global__ void __launch_bounds__(128, 1)
exp_kernel(const __nv_bfloat16* x8, const __grid_constant__ CUtensorMap tmG, float* out, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* strip = smem;            // 160 positions * 16 B
  uint8_t* gs = smem + 4096;        // [128 positions x 64 n] swizzled, 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4096 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<64>(slot);
  for (int i = threadIdx.x; i < 160; i += blockDim.x)
    reinterpret_cast<uint4*>(strip)[i] = reinterpret_cast<const uint4*>(x8)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 16384);
    tma_load_2d(&tmG, gs, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  if (threadIdx.x == 0) {
    tc_fence_after();
    for (int k = 0; k < 8; ++k) {  // 16 positions per step
      const uint32_t a_addr = smem_u32(strip) + k * 256;
      const uint64_t ad = variant == 0 ? desc_noswz(a_addr, 16, 128) : desc_noswz(a_addr, 128, 16);
      umma_bf16(tmem, ad, umma_smem_desc(smem_u32(gs) + k * 2048, 16384, 1024),
                umma_idesc_bf16(64, 64, 1, 1), k != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  uint32_t r[32];
  // M=64: accumulator rows sit in lanes 0-15 of each TMEM quadrant: row = 16*warp + lane
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(tmem + c + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    if (lane < 16)
      for (int i = 0; i < 32; ++i) out[(16 * warp + lane) * 64 + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem);
}

int main() {
  const int P = 160, N = 64;
  std::vector<__nv_bfloat16> hx(P * 8), hg(128 * N);
  std::vector<float> fx(P * 8), fg(128 * N);
  srand(3);
  for (int i = 0; i < P * 8; ++i) { float v = (rand() % 9 - 4) / 4.f; hx[i] = __float2bfloat16(v); fx[i] = v; }
  for (int i = 0; i < 128 * N; ++i) { float v = (rand() % 17 - 8) / 8.f; hg[i] = __float2bfloat16(v); fg[i] = v; }
  __nv_bfloat16 *dx, *dg; float* dout;
  cudaMalloc(&dx, P * 8 * 2); cudaMalloc(&dg, 128 * N * 2); cudaMalloc(&dout, 64 * 64 * 4);
  cudaMemcpy(dx, hx.data(), P * 8 * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dg, hg.data(), 128 * N * 2, cudaMemcpyHostToDevice);
  CUtensorMap tg;
  if (!make_tmap_2d(&tg, dg, N, 128, 64, 128)) { printf("tmap failed: %s\n", bpb_last_error()); return 1; }
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dout, 0, 64 * 64 * 4);
    exp_kernel<<<1, 128, 32 * 1024>>>(dx, tg, dout, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d (%s): kernel: %s\n", variant, variant == 0 ? "LBO=16 SBO=128" : "LBO=128 SBO=16", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> ho(64 * 64);
    cudaMemcpy(ho.data(), dout, 64 * 64 * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int p = 0; p < 128; ++p) ref += fx[p * 8 + m] * fg[p * N + n];
        maxerr = fmax(maxerr, fabs(ref - ho[m * 64 + n]));
      }
    printf("  maxerr=%g %s\n", maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
  }
  return 0;
}
