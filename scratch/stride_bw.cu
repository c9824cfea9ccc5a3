// micro-benchmark: achievable HBM bandwidth for the column-pass access pattern
// tile = N points x SEG bytes, point stride = STRIDE bytes; each CTA copies its tile in place (read + write)
#include <cstdio>
#include <cuda_runtime.h>
template <int SEG16>  // 16-byte elements per segment (8 = 128 B)
__global__ void k_copy(float4* data, size_t stride16, int npts, int tiles_per_row, size_t row16) {
  int t = blockIdx.x;
  int row = t / tiles_per_row, kt = t % tiles_per_row;
  float4* g = data + (size_t)row * row16 + (size_t)kt * SEG16;
  const int per = npts * SEG16;
  float4 buf[16];
  int cnt = 0;
  for (int w = threadIdx.x; w < per; w += blockDim.x) { int p = w / SEG16, c = w % SEG16; buf[cnt++] = g[p * stride16 + c]; if (cnt == 16) break; }
  cnt = 0;
  for (int w = threadIdx.x; w < per; w += blockDim.x) { int p = w / SEG16, c = w % SEG16; float4 v = buf[cnt++]; v.x += 1.f; g[p * stride16 + c] = v; if (cnt == 16) break; }
}
int main() {
  const int N = 512; const size_t H = 256; // float2 per row
  size_t total16 = (size_t)N * N * H / 2;  // float4 count (512^3/2 float2 = 64M float2 = 32M float4 = 537MB)
  float4* d; cudaMalloc(&d, total16 * 16); cudaMemset(d, 0, total16 * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto kern, int seg16, bool zpass) {
    // y-pass: row = z plane: row16 = N*H/2, stride16 = H/2 ; z-pass: row = y: row16 = H/2, stride16 = N*H/2
    size_t stride16 = zpass ? (size_t)N * H / 2 : H / 2, row16 = zpass ? H / 2 : (size_t)N * H / 2;
    int tiles_per_row = (H / 2) / seg16; int grid = N * tiles_per_row; int threads = N * seg16 / 16;
    for (int i = 0; i < 3; i++) kern<<<grid, threads>>>(d, stride16, N, tiles_per_row, row16);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) kern<<<grid, threads>>>(d, stride16, N, tiles_per_row, row16);
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("%-28s %s  %.3f ms  %.0f GB/s\n", name, zpass ? "z-pass(stride 1MB)" : "y-pass(stride 2KB)", ms, 2.0 * total16 * 16 / ms / 1e6);
  };
  run("seg 128B (16 cols)", k_copy<8>, 8, false);  run("seg 128B (16 cols)", k_copy<8>, 8, true);
  run("seg 256B (32 cols)", k_copy<16>, 16, false); run("seg 256B (32 cols)", k_copy<16>, 16, true);
  run("seg 512B (64 cols)", k_copy<32>, 32, false); run("seg 512B (64 cols)", k_copy<32>, 32, true);
  run("seg 64B (8 cols)", k_copy<4>, 4, false); run("seg 64B (8 cols)", k_copy<4>, 4, true);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
