// Microbenchmark: peak rate of legacy mma.sync.m16n8k16 bf16 on sm_100a (decides whether the attention
// kernels can stay on mma.sync or must move to tcgen05).  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b0 = 5, b1 = 6;
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 1 << 24);
  for (int threads : {128, 256, 512, 1024}) {
    int iters = 4096, grid = 148 * 2;
    k<<<grid, threads>>>(out, 16); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<grid, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flop = (double)grid * (threads / 32) * iters * 8 * 4096.0;
    printf("threads %4d grid %d: %.3f ms  %.1f TFLOP/s\n", threads, grid, ms, flop / ms / 1e9);
  }
  return 0;
}
