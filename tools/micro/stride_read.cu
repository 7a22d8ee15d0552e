// Micro-benchmark: every thread streams its own SEG-byte segment (thread t owns [t*SEG, (t+1)*SEG))
// in pieces of P bytes, the access pattern of a lane-per-segment DFA scan.  How does the achieved
// HBM read bandwidth depend on the piece size and on the segment size?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int P>
__global__ void __launch_bounds__(1024) reader(const uint4* text, uint64_t n_seg, uint32_t seg, uint32_t* out) {
  uint32_t acc = 0;
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n_seg; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint4* p = text + t * (seg / 16);
    for (uint32_t o = 0; o < seg / 16; o += P / 16) {
      uint4 v[P / 16];
#pragma unroll
      for (int j = 0; j < P / 16; j++) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(p + o + j));
#pragma unroll
      for (int j = 0; j < P / 16; j++) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
  }
  if (acc == 0x12345678) out[0] = acc;
}
template <int P>
void run(const uint4* text, uint64_t n, uint32_t seg, uint32_t* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {512, 1024}) {
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      reader<P><<<148 * (2048 / threads), threads>>>(text, n / seg, seg, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("seg %5u piece %4d B threads/block %4d (2048/SM): %.0f GB/s\n", seg, P, threads, n / best / 1e6);
  }
}
int main() {
  const uint64_t n = 8ull << 30;
  uint4* text; uint32_t* out;
  cudaMalloc(&text, n); cudaMalloc(&out, 4); cudaMemset(text, 1, n);
  for (uint32_t seg : {4096u, 1024u, 512u}) {
    run<16>(text, n, seg, out); run<64>(text, n, seg, out); run<128>(text, n, seg, out); run<256>(text, n, seg, out);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
