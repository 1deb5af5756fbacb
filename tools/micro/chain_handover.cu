// Period of a chain of DEPENDENT kernels (link i+1 consumes what link i wrote), launched back to back from a CUDA
// graph with programmatic dependent launch, two ways of ordering the data (development tool):
//   mode 0  griddepcontrol.wait          — the dependent's wait returns after the previous grid has completed and
//                                          flushed (what the ordered GEMM chain paid in round 1, ≈1.1 µs)
//   mode 1  device-side completion count — every CTA of link i ends with  bar.sync ; red.release.gpu(counter, 1) ;
//                                          link i+1 (already resident under PDL) polls ld.acquire.gpu(counter) until
//                                          all CTAs of links <= i have arrived, then reads; no griddepcontrol.wait
// Each link reads `bytes_in` per CTA of the previous link's output (L2 hits), adds 1 and writes its own slab, so the
// final value proves that every hand-over was observed (value == number of links).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_handover chain_handover.cu && ./chain_handover
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// x, y: [n_vec] uint4, every CTA reads the whole of x (n_vec * 16 bytes, the "activation row") and writes its own
// slice of y (n_vec / gridDim.x vectors).
template <int MODE>
__global__ void link(unsigned* counter, unsigned expect, const uint4* x, uint4* y, int n_vec, int spin) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (spin > 0) {   // emulated pre-wait prologue (weight prefetch etc.)
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
  }
  if (MODE == 0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
  } else {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      while (static_cast<int>(ld_acquire(counter) - expect) < 0) {
        if (clock64() - t0 > 2000000000ll) __trap();
      }
    }
    __syncthreads();
  }
  // consume: everybody reads the row (L2), reduce a checksum that must equal the link index everywhere
  unsigned acc = 0;
  for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
    const uint4 v = ld_cg(x + i);
    acc = max(acc, max(max(v.x, v.y), max(v.z, v.w)));
    acc = max(acc, 0u) & 0x7fffffffu;
    if (min(min(v.x, v.y), min(v.z, v.w)) != acc) acc |= 0x80000000u;   // a stale element poisons the result
  }
  // produce: my slice of y = x + 1
  __shared__ unsigned val_sh;
  const int poisoned = __syncthreads_or(static_cast<int>(acc >> 31));
  if (threadIdx.x == 0) val_sh = poisoned ? 0xdeadbeefu : (acc & 0x7fffffffu) + 1u;
  __syncthreads();
  const unsigned v = val_sh;
  const int lo = static_cast<int>(static_cast<long long>(blockIdx.x) * n_vec / gridDim.x);
  const int hi = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * n_vec / gridDim.x);
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) y[i] = make_uint4(v, v, v, v);
  if (MODE == 1) {
    __syncthreads();
    if (threadIdx.x == 0) red_release(counter, 1u);
  }
}

template <int MODE>
float run_chain(int grid, int threads, int n_vec, int links, int spin, int reps, unsigned* result) {
  unsigned* counter;
  uint4* buf[2];
  cudaMalloc(&counter, 128);
  cudaMalloc(&buf[0], n_vec * sizeof(uint4));
  cudaMalloc(&buf[1], n_vec * sizeof(uint4));
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  cudaMemsetAsync(counter, 0, 128, st);
  cudaMemsetAsync(buf[0], 0, n_vec * sizeof(uint4), st);
  for (int i = 0; i < links; ++i) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.stream = st; cfg.attrs = attr;
    cfg.numAttrs = i == 0 ? 0 : 1;   // the first link follows the memsets: fully serialised
    cudaLaunchKernelEx(&cfg, link<MODE>, counter, static_cast<unsigned>(i * grid), (const uint4*)buf[i & 1], buf[(i + 1) & 1], n_vec, spin);
  }
  cudaStreamEndCapture(st, &graph);
  cudaGraphInstantiate(&exec, graph, 0);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(exec, st);
  cudaStreamSynchronize(st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(exec, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  uint4 h;
  cudaMemcpy(&h, buf[links & 1], sizeof(h), cudaMemcpyDeviceToHost);
  *result = h.x;
  cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
  cudaFree(counter); cudaFree(buf[0]); cudaFree(buf[1]); cudaStreamDestroy(st);
  return ms * 1e3f / (reps * links);   // µs per link
}

int main() {
  const int links = 100, reps = 20;
  printf("chain of %d dependent links per graph, %d replays; us per link (final value must equal %d)\n", links, reps, links);
  for (int spin : {0, 2000}) {
    for (int grid : {32, 128, 148, 296}) {
      for (int threads : {128, 448}) {
        const int n_vec = 512;   // 8 KB row (one fp16 activation row of K = 4096)
        unsigned r0 = 0, r1 = 0;
        const float t0 = run_chain<0>(grid, threads, n_vec, links, spin, reps, &r0);
        const float t1 = run_chain<1>(grid, threads, n_vec, links, spin, reps, &r1);
        printf("grid %3d x %3d thr, row %5d B, pre-wait spin %4d cyc: griddepcontrol.wait %.3f us (final %u) | counter %.3f us (final %u)\n",
               grid, threads, n_vec * 16, spin, t0, r0, t1, r1);
      }
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
