// tools/pcie_bench.cu -- pinned H2D / D2H bandwidth alone and concurrently (measuring stick for e2e).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_=(x); if(e_!=cudaSuccess){fprintf(stderr,"%s at %d\n",cudaGetErrorString(e_),__LINE__);exit(1);} } while(0)
int main() {
  size_t n = (size_t)256 << 20;
  void *h1, *h2, *d1, *d2;
  CK(cudaMallocHost(&h1, n)); CK(cudaMallocHost(&h2, n)); CK(cudaMalloc(&d1, n)); CK(cudaMalloc(&d2, n));
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float ms;
  for (int chunkMB : {256, 32, 8}) {
    size_t c = (size_t)chunkMB << 20;
    // H2D alone
    CK(cudaEventRecord(a, s1));
    for (int r = 0; r < 4; ++r) for (size_t o = 0; o < n; o += c) CK(cudaMemcpyAsync((char*)d1 + o, (char*)h1 + o, c, cudaMemcpyHostToDevice, s1));
    CK(cudaEventRecord(b, s1)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
    printf("chunk %3d MB  H2D alone %.1f GB/s", chunkMB, 4.0 * n / ms * 1e-6);
    CK(cudaEventRecord(a, s2));
    for (int r = 0; r < 4; ++r) for (size_t o = 0; o < n; o += c) CK(cudaMemcpyAsync((char*)h2 + o, (char*)d2 + o, c, cudaMemcpyDeviceToHost, s2));
    CK(cudaEventRecord(b, s2)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
    printf("  D2H alone %.1f GB/s", 4.0 * n / ms * 1e-6);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a, 0));
    for (int r = 0; r < 4; ++r) for (size_t o = 0; o < n; o += c) {
      CK(cudaMemcpyAsync((char*)d1 + o, (char*)h1 + o, c, cudaMemcpyHostToDevice, s1));
      CK(cudaMemcpyAsync((char*)h2 + o, (char*)d2 + o, c, cudaMemcpyDeviceToHost, s2));
    }
    CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2));
    CK(cudaEventRecord(b, 0)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
    printf("  both: %.1f GB/s each direction\n", 4.0 * n / ms * 1e-6);
  }
  return 0;
}
