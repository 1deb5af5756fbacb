// Hand-over latency between dependent pieces of work, three ways (development tool):
//   (a) programmatic dependent launch: kernel i+1's griddepcontrol.wait returning after kernel i's last CTA exits
//       (what the ordered GEMM chain pays today, ≈1.1 µs in the round-1 timelines),
//   (b) a device-side flag between two resident CTAs on different SMs: st.release.gpu by the producer, ld.acquire.gpu
//       polling by the consumer — what a persistent chain kernel would pay per dependency,
//   (c) the same flag with the producer's payload (4 KB) written first and read by the consumer after the acquire.
// All times from %globaltimer (ns), median of many repetitions.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o flag_handover flag_handover.cu && ./flag_handover
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// (a) chain of tiny kernels under PDL: each stamps "my wait returned" and "I am about to exit"
__global__ void pdl_link(unsigned long long* stamps, int i) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) stamps[2 * i] = gtimer();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) stamps[2 * i + 1] = gtimer();
}

// (b)/(c) ping-pong between CTA 0 and CTA 1 (different SMs: one CTA per SM at this block size / smem)
__global__ void __launch_bounds__(128) flag_pingpong(unsigned* flags, uint4* payload, unsigned long long* out, int reps, int with_payload) {
  extern __shared__ uint8_t pad[];   // large dynamic smem: forces the two CTAs onto different SMs
  (void)pad;
  const int me = blockIdx.x, other = 1 - me;
  unsigned* my_flag = flags + me * 32;          // separate 128-byte lines
  unsigned* their_flag = flags + other * 32;
  uint4* my_buf = payload + me * 256;
  const uint4* their_buf = payload + other * 256;
  unsigned long long t0 = 0;
  uint4 sink = make_uint4(0, 0, 0, 0);
  for (int r = 1; r <= reps; ++r) {
    if ((r & 1) == me) {                        // my turn to produce
      if (with_payload) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) my_buf[i] = make_uint4(r, r, r, r);
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        t0 = gtimer();
        st_release(my_flag, static_cast<unsigned>(r));
      }
    } else {                                    // wait for the other CTA's epoch r
      if (threadIdx.x == 0) {
        while (ld_acquire(their_flag) < static_cast<unsigned>(r)) {}
      }
      __syncthreads();
      if (with_payload) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) { const uint4 v = their_buf[i]; sink.x ^= v.x; }
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) { out[me] = gtimer(); out[2 + me] = sink.x + t0; }
}

int main() {
  const int links = 200;
  unsigned long long* stamps;
  cudaMallocManaged(&stamps, sizeof(unsigned long long) * 2 * links);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  for (int grid : {1, 128}) {
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = 0; i < links; ++i) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, pdl_link, stamps, i);
      }
      cudaStreamSynchronize(st);
    }
    std::vector<double> gaps;
    for (int i = 1; i < links; ++i) gaps.push_back(double(stamps[2 * i]) - double(stamps[2 * (i - 1) + 1]));
    std::sort(gaps.begin(), gaps.end());
    printf("(a) PDL hand-over, grid %3d: last-CTA-exit -> next wait returned: median %.0f ns (p10 %.0f, p90 %.0f)\n", grid,
           gaps[gaps.size() / 2], gaps[gaps.size() / 10], gaps[gaps.size() * 9 / 10]);
  }

  unsigned* flags; uint4* payload; unsigned long long* out;
  cudaMalloc(&flags, 64 * sizeof(unsigned)); cudaMalloc(&payload, 512 * sizeof(uint4)); cudaMallocManaged(&out, 4 * sizeof(unsigned long long));
  const int smem = 120 * 1024;
  cudaFuncSetAttribute(flag_pingpong, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int with_payload = 0; with_payload < 2; ++with_payload) {
    const int reps = 20000;
    for (int pass = 0; pass < 2; ++pass) {
      cudaMemset(flags, 0, 64 * sizeof(unsigned));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0, st);
      flag_pingpong<<<2, 128, smem, st>>>(flags, payload, out, reps, with_payload);
      cudaEventRecord(e1, st);
      cudaStreamSynchronize(st);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      if (pass == 1)
        printf("(%c) flag hand-over between two CTAs%s: %.0f ns per one-way hand-over (%d hand-overs in %.3f ms)\n",
               with_payload ? 'c' : 'b', with_payload ? " + 4 KB payload write/read" : "", ms * 1e6 / reps, reps, ms);
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
