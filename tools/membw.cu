// Memory-path microbenchmark for B200 (sm_100a): what the RoIAlign kernels can expect from HBM, L2 and the bulk-copy
// engine.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membw tools/membw.cu ; run: tools/membw
//   read_ldg   : LDG.128 streaming reads, persistent grid
//   read_bulk  : cp.async.bulk (UBLKCP) global -> shared ring, one issuing thread per CTA
//   write_stg  : STG.128 streaming writes
// Each is run over a DRAM-sized buffer (4 GiB, one pass) and an L2-resident buffer (32 MiB, many passes).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e = (x);                                                                        \
    if (e != cudaSuccess) {                                                                     \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));                                     \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

__global__ void __launch_bounds__(256) read_ldg(const uint4* __restrict__ p, size_t n16, int passes, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int ps = 0; ps < passes; ps++) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
      uint4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
      acc.x ^= a.x ^ b.x ^ c.x ^ d.x;
      acc.y ^= a.y ^ b.y ^ c.y ^ d.y;
      acc.z ^= a.z ^ b.z ^ c.z ^ d.z;
      acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
    }
    for (; i < n16; i += stride) {
      uint4 a = __ldg(p + i);
      acc.x ^= a.x; acc.y ^= a.y; acc.z ^= a.z; acc.w ^= a.w;
    }
  }
  if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) *sink = acc;
}

__global__ void __launch_bounds__(256) write_stg(uint4* __restrict__ p, size_t n16, int passes) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int ps = 0; ps < passes; ps++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
      p[i] = make_uint4(ps, 1, 2, 3);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(128) read_bulk(const char* __restrict__ p, size_t bytes, int passes, uint32_t* sink) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t full[STAGES];
  const size_t nchunks = bytes / CHUNK;
  const size_t per = (nchunks + gridDim.x - 1) / gridDim.x;
  const size_t c0 = blockIdx.x * per, c1 = c0 + per < nchunks ? c0 + per : nchunks;
  const long n = c1 > c0 ? (long)(c1 - c0) * passes : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long i) {
    const int st = (int)(i % STAGES);
    const size_t c = c0 + (size_t)(i % (long)(c1 - c0));
    mbar_expect_tx(&full[st], CHUNK);
    bulk_g2s(smem + (size_t)st * CHUNK, p + c * CHUNK, CHUNK, &full[st]);
  };
  if (threadIdx.x == 0)
    for (long i = 0; i < STAGES - 1 && i < n; i++) issue(i);
  uint32_t acc = 0;
  for (long i = 0; i < n; i++) {
    const int st = (int)(i % STAGES);
    if (threadIdx.x == 0 && i + STAGES - 1 < n) issue(i + STAGES - 1);
    mbar_wait(&full[st], (uint32_t)((i / STAGES) & 1));
    acc ^= *reinterpret_cast<const uint32_t*>(smem + (size_t)st * CHUNK + 4 * threadIdx.x);
    __syncthreads();
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  const int sms = pr.multiProcessorCount;
  printf("device %s, %d SMs, L2 %.0f MB\n", pr.name, sms, pr.l2CacheSize / 1048576.0);
  const size_t big = (size_t)4 << 30;
  char* buf;
  CK(cudaMalloc(&buf, big));
  CK(cudaMemset(buf, 1, big));
  uint4* sink;
  CK(cudaMalloc(&sink, 64));
  struct Case { const char* name; size_t bytes; int passes; } cases[] = {
      {"HBM 4 GiB x1", big, 1}, {"L2 16 MiB x128", (size_t)16 << 20, 128}, {"L2 32 MiB x64", (size_t)32 << 20, 64},
      {"L2 64 MiB x32", (size_t)64 << 20, 32}, {"L2? 96 MiB x32", (size_t)96 << 20, 32}, {"192 MiB x16", (size_t)192 << 20, 16}};
  for (auto& c : cases) {
    for (int cps = 2; cps <= 8; cps *= 2) {
      float ms = time_ms([&] { read_ldg<<<sms * cps, 256>>>((const uint4*)buf, c.bytes / 16, c.passes, sink); }, 3);
      printf("read_ldg   %-16s %d CTA/SM x256thr : %8.1f GB/s\n", c.name, cps, c.bytes * (double)c.passes / ms / 1e6);
    }
    {
      constexpr int CH = 16384, ST = 4;
      auto k = read_bulk<CH, ST>;
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST));
      for (int cps = 1; cps <= 3; cps++) {
        float ms = time_ms([&] { k<<<sms * cps, 128, CH * ST>>>(buf, c.bytes, c.passes, (uint32_t*)sink); }, 3);
        printf("read_bulk  %-16s %d CTA/SM 16KBx4     : %8.1f GB/s\n", c.name, cps, c.bytes * (double)c.passes / ms / 1e6);
      }
    }
    {
      constexpr int CH = 4096, ST = 8;
      auto k = read_bulk<CH, ST>;
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST));
      for (int cps = 2; cps <= 4; cps += 2) {
        float ms = time_ms([&] { k<<<sms * cps, 128, CH * ST>>>(buf, c.bytes, c.passes, (uint32_t*)sink); }, 3);
        printf("read_bulk  %-16s %d CTA/SM 4KBx8      : %8.1f GB/s\n", c.name, cps, c.bytes * (double)c.passes / ms / 1e6);
      }
    }
    {
      constexpr int CH = 1024, ST = 16;
      auto k = read_bulk<CH, ST>;
      for (int cps = 4; cps <= 8; cps += 4) {
        float ms = time_ms([&] { k<<<sms * cps, 128, CH * ST>>>(buf, c.bytes, c.passes, (uint32_t*)sink); }, 3);
        printf("read_bulk  %-16s %d CTA/SM 1KBx16     : %8.1f GB/s\n", c.name, cps, c.bytes * (double)c.passes / ms / 1e6);
      }
    }
    float ms = time_ms([&] { write_stg<<<sms * 8, 256>>>((uint4*)buf, c.bytes / 16, c.passes); }, 3);
    printf("write_stg  %-16s 8 CTA/SM x256thr : %8.1f GB/s\n", c.name, c.bytes * (double)c.passes / ms / 1e6);
  }
  printf("ok\n");
  return 0;
}
