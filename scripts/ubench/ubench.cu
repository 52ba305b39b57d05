// ubench.cu -- micro-benchmarks behind the round-2 design decisions: what bounds a random
// table gather on B200 -- L2 requests, L2 sectors, shared-memory bank conflicts or DSMEM -- measured with the
// access shapes the colorlut evaluators could use.  One "item" = one pixel (Q1) or one pixel-channel (Q2/Q3);
// every run covers 8 294 400 pixels (one 4K frame) so the numbers compare directly with the kernel timings.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/ubench/ubench scripts/ubench/ubench.cu
//   run  : scripts/ubench/ubench > gpurun_out/ubench.jsonl
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // cheap integer hash (lowbias32)
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
struct V256 { unsigned long long a, b, c, d; };
__device__ __forceinline__ V256 ldg256(const void *p) {
  V256 v;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t fold(const V256 &v) { return (uint32_t)(v.a ^ v.b ^ v.c ^ v.d) ^ (uint32_t)((v.a ^ v.b ^ v.c ^ v.d) >> 32); }

// ---- Q1: gathers from an L2-resident table of `lines` 128-byte lines ------------------------------------------
// MODE 0: 1 x 4 B   1: 1 x 32 B   2: 3 x 32 B same line   3: 4 x 32 B same line
//      4: 4 x 32 B in 4 different lines (the x-pair layout: entries i, i+33, i+1089, i+1122 of 32 B each)
//      5: 2 lines x 2 sectors (64-byte xy-quad entries i, i+1089)
// COH: coherent indices (64 consecutive items share one cell) instead of random ones
template <int MODE, bool COH>
__global__ void __launch_bounds__(256) q1_kernel(const uint8_t *__restrict__ tab, uint32_t lines, uint32_t items, uint32_t *out) {
  uint32_t acc = 0;
  constexpr int U = 4;
  for (uint32_t i0 = (blockIdx.x * 256 + threadIdx.x); i0 < items; i0 += gridDim.x * 256 * U) {
    uint32_t idx[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t it = i0 + u * gridDim.x * 256;
      idx[u] = (COH ? mix32(it >> 6) : mix32(it)) % lines;
    }
    if (MODE == 0) {
      uint32_t v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = __ldg(reinterpret_cast<const uint32_t *>(tab + (size_t)idx[u] * 128));
#pragma unroll
      for (int u = 0; u < U; u++) acc ^= v[u];
    } else if (MODE <= 3) {
      constexpr int NS = MODE == 1 ? 1 : (MODE == 2 ? 3 : 4);
      V256 v[U][NS];
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int s = 0; s < NS; s++) v[u][s] = ldg256(tab + (size_t)idx[u] * 128 + 32 * s);
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int s = 0; s < NS; s++) acc ^= fold(v[u][s]);
    } else if (MODE == 4) {
      V256 v[U][4];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint8_t *b = tab + (size_t)idx[u] * 32;   // 32-byte entries: the table is used as lines*4 entries
        v[u][0] = ldg256(b); v[u][1] = ldg256(b + 33 * 32); v[u][2] = ldg256(b + 1089 * 32); v[u][3] = ldg256(b + 1122 * 32);
      }
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int s = 0; s < 4; s++) acc ^= fold(v[u][s]);
    } else {
      V256 v[U][4];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint8_t *b = tab + (size_t)idx[u] * 64;   // 64-byte entries
        v[u][0] = ldg256(b); v[u][1] = ldg256(b + 32); v[u][2] = ldg256(b + 1089 * 64); v[u][3] = ldg256(b + 1089 * 64 + 32);
      }
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int s = 0; s < 4; s++) acc ^= fold(v[u][s]);
    }
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}

// ---- Q2: one channel of a 34^3 f32 LUT (157 216 B) resident in shared memory; item = pixel-channel ------------
// MODE 0: 8 x LDS.32 at base + {0,1,34,35,1156,1157,1190,1191}   1: 4 x LDS.64 (x-pairs, even base)   2: 1 x LDS.32
template <int MODE, bool COH>
__global__ void __launch_bounds__(1024) q2_kernel(uint32_t items, uint32_t span, uint32_t *out) {
  extern __shared__ __align__(16) float lut[];
  constexpr int NW = 34 * 34 * 34;
  for (int i = threadIdx.x; i < NW; i += blockDim.x) lut[i] = (float)i;
  __syncthreads();
  float acc = 0.f;
  const uint32_t range = span ? span : (uint32_t)(NW - 1192);
  constexpr int U = 4;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < items; i0 += stride * U) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t it = i0 + u * stride;
      uint32_t b = (COH ? mix32(it >> 6) : mix32(it)) % range;
      if (MODE == 0) {
        acc += lut[b] + lut[b + 1] + lut[b + 34] + lut[b + 35] + lut[b + 1156] + lut[b + 1157] + lut[b + 1190] + lut[b + 1191];
      } else if (MODE == 1) {
        b &= ~1u;
        const float2 p0 = *reinterpret_cast<const float2 *>(lut + b), p1 = *reinterpret_cast<const float2 *>(lut + b + 34);
        const float2 p2 = *reinterpret_cast<const float2 *>(lut + b + 1156), p3 = *reinterpret_cast<const float2 *>(lut + b + 1190);
        acc += p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y;
      } else {
        acc += lut[b];
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(acc);
}

// ---- Q3: the same 8 corner reads with the LUT split over a 2-CTA cluster (z-halves): DSMEM ---------------------
template <bool COH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024) q3_kernel(uint32_t items, uint32_t *out) {
  extern __shared__ __align__(16) float lut[];
  constexpr int NW = 34 * 34 * 17 + 1192;
  cg::cluster_group cl = cg::this_cluster();
  for (int i = threadIdx.x; i < NW; i += blockDim.x) lut[i] = (float)i;
  cl.sync();
  const float *mine = lut, *other = cl.map_shared_rank(lut, cl.block_rank() ^ 1);
  float acc = 0.f;
  constexpr int U = 4;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < items; i0 += stride * U) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t it = i0 + u * stride;
      const uint32_t h = COH ? mix32(it >> 6) : mix32(it);
      const uint32_t b = h % (uint32_t)(34 * 34 * 17);
      const float *t = (h >> 31) ? other : mine;   // half of the items live in the other CTA's half
      acc += t[b] + t[b + 1] + t[b + 34] + t[b + 35] + t[b + 1156] + t[b + 1157] + t[b + 1190] + t[b + 1191];
    }
  }
  cl.sync();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(acc);
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a));
    launch();
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
  const uint32_t PIX = 3840u * 2160u;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint32_t *out;
  CK(cudaMalloc(&out, 4u << 20));
  // Q1
  for (int big = 0; big < 2; big++) {
    const uint32_t lines = big ? 275000u : 36000u;   // 33^3 / 65^3 cells of 128 B
    uint8_t *tab;
    const size_t bytes = (size_t)lines * 128 + 1200 * 64 * 2;
    CK(cudaMalloc(&tab, bytes));
    CK(cudaMemset(tab, 1, bytes));
    const int grid = sms * 8;
#define Q1(M, C) printf("{\"q\":\"q1\",\"table_MB\":%.1f,\"mode\":%d,\"coherent\":%d,\"us_per_4k_frame\":%.2f}\n", bytes / 1e6, M, C, \
                        1e3f * time_ms([&] { q1_kernel<M, C><<<grid, 256>>>(tab, lines, PIX, out); }))
    Q1(0, false); Q1(1, false); Q1(2, false); Q1(3, false); Q1(4, false); Q1(5, false);
    Q1(0, true); Q1(2, true); Q1(3, true); Q1(4, true);
#undef Q1
    CK(cudaFree(tab));
  }
  // Q2: items = 3 channels x PIX
  {
    const int smem = 34 * 34 * 34 * 4;
#define Q2(M, C, SPAN) do { CK(cudaFuncSetAttribute(q2_kernel<M, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    printf("{\"q\":\"q2\",\"mode\":%d,\"coherent\":%d,\"span\":%d,\"us_per_4k_frame_3ch\":%.2f}\n", M, C, SPAN, \
           1e3f * time_ms([&] { q2_kernel<M, C><<<sms, 1024, smem>>>(3 * PIX, SPAN, out); })); } while (0)
    Q2(0, false, 0); Q2(0, true, 0); Q2(1, false, 0); Q2(1, true, 0);
    Q2(2, false, 0); Q2(2, false, 2048); Q2(2, false, 8192); Q2(2, true, 0);
#undef Q2
  }
  // Q3
  {
    const int smem = (34 * 34 * 17 + 1192) * 4;
    CK(cudaFuncSetAttribute(q3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(q3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = sms - sms % 2;
    printf("{\"q\":\"q3_dsmem\",\"coherent\":0,\"us_per_4k_frame_3ch\":%.2f}\n", 1e3f * time_ms([&] { q3_kernel<false><<<grid, 1024, smem>>>(3 * PIX, out); }));
    printf("{\"q\":\"q3_dsmem\",\"coherent\":1,\"us_per_4k_frame_3ch\":%.2f}\n", 1e3f * time_ms([&] { q3_kernel<true><<<grid, 1024, smem>>>(3 * PIX, out); }));
  }
  return 0;
}
