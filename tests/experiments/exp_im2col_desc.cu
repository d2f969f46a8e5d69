// Experiment (not product code): can a K-major, NON-swizzled UMMA A-descriptor read the implicit
// im2col matrix of a width-W window straight from the channels-last strip it is a view of?
//   strip: x8[p][8] bf16, 16 bytes per position.  im2col row t = x8[t .. t+W-1][0..7] (8W values).
//   no-swizzle K-major canonical layout: 8x8 core matrices whose rows are 16 B apart;
//   element (row r, k) of the view lives at byte 16*(r + k/8) + 2*(k%8): a core matrix
//   (rows r0..r0+7, k-block kb) is the contiguous 128 B at 16*(r0 + kb); the next core matrix along K
//   is +16 B (LBO), along M +128 B (SBO).  Core matrices overlap in memory, which reads do not mind.
// One 128 x 64 x K MMA chain (K = 16 per instruction, start address += 32 B per step) against a
// swizzled K-major B tile loaded by TMA; compared with a host reference.
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
  return d;  // layout type 0: no swizzle
}

constexpr int KSTEPS = 13;  // K = 208 >= 8 * 25

__global__ void __launch_bounds__(128, 1)
exp_kernel(const __nv_bfloat16* x8, const __grid_constant__ CUtensorMap tmW, float* out, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* strip = smem;                 // 160 positions * 16 B = 2560 B
  uint8_t* ws = smem + 4096;             // 4 tiles [64 rows x 64 k] swizzled = 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4096 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(slot);
  for (int i = threadIdx.x; i < 160 * 16 / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(strip)[i] = reinterpret_cast<const uint4*>(x8)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 32768);
    for (int c = 0; c < 4; ++c) tma_load_2d(&tmW, ws + c * 8192, &bars[0], c * 64, 0);
  }
  mbar_wait(&bars[0], 0);
  if (threadIdx.x == 0) {
    tc_fence_after();
    for (int k = 0; k < KSTEPS; ++k) {
      const uint32_t a_addr = smem_u32(strip) + k * 32;
      const uint32_t b_addr = smem_u32(ws) + (k / 4) * 8192 + (k % 4) * 32;
      const uint64_t ad = variant == 0 ? desc_noswz(a_addr, 16, 128) : desc_noswz(a_addr, 128, 16);
      umma_bf16(tmem, ad, umma_smem_desc(b_addr, 0, 1024), umma_idesc_bf16(128, 64, 0, 0), k != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  uint32_t r[32];
  float* o = out + (warp * 32 + lane) * 64;
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(tmem + c + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) o[c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem);
}

int main() {
  const int P = 160, W = 25, N = 64, K = 256;
  std::vector<__nv_bfloat16> hx(P * 8), hw(N * K);
  std::vector<float> fx(P * 8), fw(N * K, 0.f);
  srand(3);
  for (int p = 0; p < P; ++p) {
    const int base = rand() % 4;
    for (int c = 0; c < 8; ++c) { float v = (c == base) ? 1.f : 0.f; hx[p * 8 + c] = __float2bfloat16(v); fx[p * 8 + c] = v; }
  }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float v = (k < 8 * W && (k % 8) < 4) ? (rand() % 17 - 8) / 8.f : 0.f;
      hw[n * K + k] = __float2bfloat16(v); fw[n * K + k] = v;
    }
  __nv_bfloat16 *dx, *dw; float* dout;
  cudaMalloc(&dx, P * 8 * 2); cudaMalloc(&dw, N * K * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dx, hx.data(), P * 8 * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), N * K * 2, cudaMemcpyHostToDevice);
  CUtensorMap tw;
  if (!make_tmap_2d(&tw, dw, K, N, 64, 64)) { printf("tmap failed: %s\n", bpb_last_error()); return 1; }
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dout, 0, 128 * 64 * 4);
    exp_kernel<<<1, 128, 48 * 1024>>>(dx, tw, dout, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d (%s): kernel: %s\n", variant, variant == 0 ? "LBO=16 SBO=128" : "LBO=128 SBO=16", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> ho(128 * 64);
    cudaMemcpy(ho.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int t = 0; t < 128; ++t)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < 16 * KSTEPS; ++k) ref += fx[t * 8 + k] * fw[n * K + k];
        maxerr = fmax(maxerr, fabs(ref - ho[t * 64 + n]));
      }
    printf("  maxerr=%g %s\n", maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
  }
  return 0;
}
