// Experiment (not product code): can a K-major SWIZZLE_128B UMMA operand start at a row that is
// not a multiple of 8 (start address not 1024-byte aligned), and what must the descriptor's
// base_offset field be?  Loads 256 rows x 64 bf16 with TMA, then for each row shift s runs one
// 128x64x64 MMA whose A descriptor starts at row s, with base_offset = 0 and = (addr >> 7) & 7.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../bpreveal_b200/csrc/common.cuh"
#include "../bpreveal_b200/csrc/ptx.cuh"

using namespace bp;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = umma_smem_desc(saddr, 0, sbo);
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1)
exp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
           float* out, int n_shift, const int* shifts) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* xs = smem;               // 256 rows * 128 B
  uint8_t* ws = smem + 32768;       // 64 rows * 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 32768 + 8192);
    tma_load_3d(&tmX, xs, &bars[0], 0, 0, 0);
    tma_load_3d(&tmX, xs + 16384, &bars[0], 0, 128, 0);
    tma_load_2d(&tmW, ws, &bars[0], 0, 0);
  }
  mbar_wait(&bars[0], 0);
  uint32_t phase = 0;
  for (int si = 0; si < n_shift; ++si) {
    for (int variant = 0; variant < 2; ++variant) {
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t a_addr = smem_u32(xs) + shifts[si] * 128;
        const uint32_t bo = variant ? ((a_addr >> 7) & 7) : 0;
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, desc_bo(a_addr + k * 32, 1024, bo), umma_smem_desc(smem_u32(ws) + k * 32, 0, 1024),
                    umma_idesc_bf16(128, 64, 0, 0), k != 0);
        umma_commit(&bars[1]);
      }
      mbar_wait(&bars[1], phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t r[32];
      float* o = out + ((size_t)(si * 2 + variant) * 128 + warp * 32 + lane) * 64;
      for (int c = 0; c < 64; c += 32) {
        tmem_ld32(tmem + c + ((uint32_t)(warp * 32) << 16), r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) o[c + i] = __uint_as_float(r[i]);
      }
      tc_fence_before();
      __syncthreads();
    }
  }
  if (warp == 1) tmem_dealloc<64>(tmem);
}

int main() {
  const int R = 256, F = 64;
  std::vector<__nv_bfloat16> hx(R * F), hw(F * F);
  std::vector<float> fx(R * F), fw(F * F);
  srand(1);
  for (int i = 0; i < R * F; ++i) { float v = (rand() % 17 - 8) / 8.f; hx[i] = __float2bfloat16(v); fx[i] = v; }
  for (int i = 0; i < F * F; ++i) { float v = (rand() % 9 - 4) / 4.f; hw[i] = __float2bfloat16(v); fw[i] = v; }
  __nv_bfloat16 *dx, *dw;
  cudaMalloc(&dx, R * F * 2);
  cudaMalloc(&dw, F * F * 2);
  cudaMemcpy(dx, hx.data(), R * F * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), F * F * 2, cudaMemcpyHostToDevice);
  std::vector<int> shifts = {0, 8, 1, 2, 3, 4, 5, 7, 9, 23, 64, 100};
  int* dsh;
  cudaMalloc(&dsh, shifts.size() * 4);
  cudaMemcpy(dsh, shifts.data(), shifts.size() * 4, cudaMemcpyHostToDevice);
  float* dout;
  const size_t n_out = shifts.size() * 2 * 128 * 64;
  cudaMalloc(&dout, n_out * 4);
  CUtensorMap tx, tw;
  if (!make_tmap_3d(&tx, dx, F, R, 1, 64, 128) || !make_tmap_2d(&tw, dw, F, F, 64, 64)) {
    printf("tmap failed: %s\n", bpb_last_error());
    return 1;
  }
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  exp_kernel<<<1, 128, 48 * 1024, 0>>>(tx, tw, dout, (int)shifts.size(), dsh);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> ho(n_out);
  cudaMemcpy(ho.data(), dout, n_out * 4, cudaMemcpyDeviceToHost);
  for (size_t si = 0; si < shifts.size(); ++si) {
    for (int variant = 0; variant < 2; ++variant) {
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += fx[(shifts[si] + m) * F + k] * fw[n * F + k];
          double got = ho[((si * 2 + variant) * 128 + m) * 64 + n];
          maxerr = fmax(maxerr, fabs(ref - got));
        }
      printf("shift %3d base_offset=%s maxerr=%g %s\n", shifts[si], variant ? "(addr>>7)&7" : "0",
             maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  }
  return 0;
}
