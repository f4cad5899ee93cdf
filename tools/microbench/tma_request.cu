// Micro-benchmark: what does ONE SM's TMA unit sustain as a function of the box shape?
// One CTA per SM; an elected thread keeps DEPTH tiled loads in flight from an L2-resident bf16 matrix [R, 64 * C]
// viewed as a 3-D tensor (64 elements = 128 B innermost, R rows, C column chunks); box = [64, rows, chunks].
// Prints cycles per request and bytes per clock for several (rows, chunks).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_request tools/microbench/tma_request.cu -lcuda && /tmp/tma_request
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(128, 1)
tma_kernel(const __grid_constant__ CUtensorMap tm, int box_bytes, int depth, int iters, int rows_total, int box_rows,
           int dims, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // rounds of `depth` back-to-back requests, then one wait per request: time per round = depth * service + latency
    const int row_tiles = rows_total / box_rows;
    long long t0 = 0;
    for (int round = 0; round < iters + 1; ++round) {
      if (round == 1) t0 = clock64();
      for (int slot = 0; slot < depth; ++slot) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[slot])), "r"(box_bytes) : "memory");
        const int r0 = ((blockIdx.x * 7 + round * depth + slot) % row_tiles) * box_rows;
        if (dims == 3)
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(smem_u32(smem + static_cast<size_t>(slot) * box_bytes)), "l"(reinterpret_cast<uint64_t>(&tm)),
                         "r"(smem_u32(&bar[slot])), "r"(0), "r"(r0), "r"(0) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(smem + static_cast<size_t>(slot) * box_bytes)), "l"(reinterpret_cast<uint64_t>(&tm)),
                         "r"(smem_u32(&bar[slot])), "r"(0), "r"(r0) : "memory");
      }
      for (int slot = 0; slot < depth; ++slot) {
        uint32_t done = 0;
        while (!done)
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                       : "=r"(done) : "r"(smem_u32(&bar[slot])), "r"(round & 1) : "memory");
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int R = 4096, C = 16;                       // 4096 rows x 1024 bf16 = 8 MB: L2 resident
  void* buf;
  CK(cudaMalloc(&buf, static_cast<size_t>(R) * C * 128));
  CK(cudaMemset(buf, 0, static_cast<size_t>(R) * C * 128));
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn encode = reinterpret_cast<EncodeFn>(fnp);
  CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  printf("%6s %6s %9s %5s %5s\n", "rows", "chunks", "box_bytes", "dims", "CTAs");
  const int shapes[][2] = {{8, 1}, {16, 1}, {64, 1}, {128, 1}, {16, 2}, {16, 4}, {64, 2}, {64, 4}, {16, 8}, {64, 8}, {128, 2}, {128, 4}, {32, 4}};
  for (auto& sh : shapes) {
    const int rows = sh[0], chunks = sh[1];
    const int box_bytes = rows * chunks * 128;
    for (int dims : {2, 3}) {
      if (dims == 2 && chunks != 1) continue;
      CUtensorMap tm;
      cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(R), static_cast<cuuint64_t>(C)};
      cuuint64_t gstr[2] = {static_cast<cuuint64_t>(C) * 128, 128};   // row pitch, chunk pitch (bytes)
      cuuint32_t box[3] = {64, static_cast<cuuint32_t>(rows), static_cast<cuuint32_t>(chunks)};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d for %d x %d\n", (int)r, rows, chunks); continue; }
      for (int grid : {16, 148}) {
        double per_round[2] = {0, 0};
        int depths[2] = {2, 6};
        for (int di = 0; di < 2; ++di) {
          const int depth = depths[di];
          if (static_cast<long>(depth) * box_bytes > 200 * 1024) { per_round[di] = -1; continue; }
          const int iters = 200;
          for (int rep = 0; rep < 2; ++rep) {
            tma_kernel<<<grid, 128, 208 * 1024>>>(tm, box_bytes, depth, iters, R, rows, dims, cyc);
            CK(cudaDeviceSynchronize());
          }
          long long h[148];
          CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
          double mean = 0;
          for (int i = 0; i < grid; ++i) mean += h[i];
          per_round[di] = mean / grid / iters;
        }
        if (per_round[1] < 0) { printf("%6d %6d %9d %4dD %5d | round(2) %8.1f\n", rows, chunks, box_bytes, dims, grid, per_round[0]); continue; }
        const double service = (per_round[1] - per_round[0]) / 4.0;   // extra cycles per extra request in flight
        printf("%6d %6d %9d %4dD %5d | round(2) %8.1f round(6) %8.1f -> %7.1f cyc/request, %6.1f B/clk/SM, %6.1f GB/s/SM\n", rows,
               chunks, box_bytes, dims, grid, per_round[0], per_round[1], service, box_bytes / service, box_bytes / service * 1.965);
      }
    }
  }
  return 0;
}
