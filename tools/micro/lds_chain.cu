// Micro-benchmark: dependent table walk in shared memory, the inner loop of scan_rev_fast.
//   lds_chain <mode> : per (warps per SM) prints cycles per dependent step.
// mode 0: LDS.U8  + IMAD  (one-byte entries, 288-byte rows)
// mode 1: LDS.32  + LOP3  (four-byte entries holding row addresses)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
extern __shared__ __align__(1024) unsigned char smem[];
template <int MODE>
__global__ void chain(const uint32_t* text, uint32_t* out, long long* cycles, int steps16, int spread) {
  const uint32_t tb = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  // table: 8 states; next = (state + byte) & 7 pattern, arbitrary but data dependent
  if (MODE == 0) {
    for (uint32_t i = threadIdx.x; i < 8 * 256; i += blockDim.x) {
      uint32_t r = i >> 8, b = i & 255;
      asm volatile("st.shared.u8 [%0], %1;" ::"r"(tb + r * 288 + b), "r"((r + b) & 7));
    }
  } else {
    for (uint32_t i = threadIdx.x; i < 8 * 256; i += blockDim.x) {
      uint32_t r = i >> 8, b = i & 255, nx = (r + b) & 7;
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(tb + (r << 10) + ((b ^ (r & 31)) << 2)), "r"(tb + (nx << 10) + ((nx & 31) << 2)));
    }
  }
  __syncthreads();
  const uint32_t s0 = spread ? (threadIdx.x & 7) : 3;
  uint32_t e = MODE == 0 ? s0 : tb + (s0 << 10) + (s0 << 2);
  uint32_t w[4];
  for (int i = 0; i < 4; i++) w[i] = text[(threadIdx.x * spread + i) & 1023];
  uint32_t bits = 0;
  const uint32_t thr = MODE == 0 ? 6 : tb + (6 << 10);
  long long t0 = clock64();
  for (int it = 0; it < steps16; it++) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (MODE == 0) {
          uint32_t x = __byte_perm(w[g], tb, 0x7650 + k), addr;
          asm("mad.lo.u32 %0, %1, 288, %2;" : "=r"(addr) : "r"(e), "r"(x));
          asm volatile("ld.shared.u8 %0, [%1];" : "=r"(e) : "r"(addr));
        } else {
          uint32_t idx = k == 0 ? (w[g] << 2) : (w[g] >> (8 * k - 2));
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"((idx & 0x3FCu) ^ e));
        }
        if (e >= thr) bits |= 1u << (4 * g + k);
      }
      w[g] = w[g] * 1664525u + 1013904223u + bits;  // keep bytes changing
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = e + bits;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main(int argc, char** argv) {
  uint32_t* text; uint32_t* out; long long* cyc;
  cudaMalloc(&text, 4096); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  uint32_t h[1024];
  for (int i = 0; i < 1024; i++) h[i] = 0x61626364u + i * 0x01030507u;  // letters-ish
  for (int i = 0; i < 1024; i++) { uint32_t v = 0; for (int b = 0; b < 4; b++) v |= (uint32_t)(97 + ((i * 7 + b * 13 + (i >> 3)) % 26)) << (8 * b); h[i] = v; }
  cudaMemcpy(text, h, 4096, cudaMemcpyHostToDevice);
  const int steps16 = 2000;
  cudaFuncSetAttribute(chain<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(chain<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int spread = 0; spread < 2; spread++)
  for (int mode = 0; mode < 2; mode++)
    for (int threads : {32, 256, 512, 1024}) {
      long long c = 0;
      for (int rep = 0; rep < 2; rep++) {
        if (mode == 0) chain<0><<<148, threads, 16 * 1024>>>(text, out, cyc, steps16, spread);
        else chain<1><<<148, threads, 16 * 1024>>>(text, out, cyc, steps16, spread);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("spread %d mode %d (%s) warps/SM %2d: %.1f cycles per dependent step, %.2f bytes/cycle/SM\n", spread, mode, mode ? "LDS.32+LOP3" : "LDS.U8+IMAD",
             threads / 32, (double)c / (steps16 * 16), threads * 16.0 * steps16 / c);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
