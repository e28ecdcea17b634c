// Step A of the tcgen05 path: a standalone int8 GEMM on tcgen05.mma.kind::i8 with TMA-staged
// SWIZZLE_128B K-major operands and int32 accumulators in TMEM.
//   C[M][N] (int32, row-major) = sum_k A[M][K] * B[N][K]      (both operands K-major int8)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/i8gemm_test tools/i8gemm_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int BM = 128, BN = 128, BKB = 128;   // BKB = K bytes (= int8 elements) per stage row = one 128B swizzle row
constexpr int STAGES = 4;
constexpr int UMMA_K = 32;                     // int8 elements per MMA
constexpr int TILE_A_BYTES = BM * BKB, TILE_B_BYTES = BN * BKB;
constexpr int STAGE_BYTES = TILE_A_BYTES + TILE_B_BYTES;
constexpr int NTHREADS = 192;                  // warp0 TMA, warp1 MMA, warps 2..5 epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B canonical layout: 8-row groups of 128-byte rows (1024 B apart)
__device__ __forceinline__ uint64_t make_desc(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);   // start address
  d |= (uint64_t)1 << 16;                            // LBO (ignored for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                  // SBO = 1024 B
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
i8gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
              int32_t* __restrict__ C, int M, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.y, tile_n = blockIdx.x;
  const int num_kb = K / BKB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one warp allocates TMEM (128 columns of int32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0 && lane == 0) {
    // ---- TMA producer ----
    for (int kb = 0; kb < num_kb; kb++) {
      const int s = kb % STAGES;
      if (kb >= STAGES) mbar_wait(&empty_bar[s], ((kb / STAGES) - 1) & 1);
      uint8_t* sa = smem + s * STAGE_BYTES;
      uint8_t* sb = sa + TILE_A_BYTES;
      mbar_expect_tx(&full_bar[s], STAGE_BYTES);
      tma_load_2d(&mapA, &full_bar[s], sa, kb * BKB, tile_m * BM);
      tma_load_2d(&mapB, &full_bar[s], sb, kb * BKB, tile_n * BN);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer (single thread) ----
    // idesc: c_format S32 (2) @4, a/b signed int8 (1) @7/@10, K-major both, N>>3 @17, M>>4 @24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int kb = 0; kb < num_kb; kb++) {
      const int s = kb % STAGES;
      mbar_wait(&full_bar[s], (kb / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint8_t* sa = smem + s * STAGE_BYTES;
      const uint8_t* sb = sa + TILE_A_BYTES;
      const uint64_t da = make_desc(sa), db = make_desc(sb);
#pragma unroll
      for (int k = 0; k < BKB / UMMA_K; k++) {
        // advance along K inside the 128-byte swizzle row: +32 bytes = +2 in 16-byte units
        umma_i8(tmem_base, da + (uint64_t)(k * UMMA_K >> 4), db + (uint64_t)(k * UMMA_K >> 4), idesc, (kb | k) != 0);
      }
      umma_commit(&empty_bar[s]);   // frees the smem stage when these MMAs have read it
    }
    umma_commit(&acc_bar);          // accumulator complete
  } else if (warp >= 2) {
    // ---- epilogue: TMEM -> registers -> global ----
    mbar_wait(&acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = tile_m * BM + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      int32_t* dst = C + (size_t)row * N + tile_n * BN + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<int4*>(dst + j) = make_int4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* ptr, uint64_t rows, uint64_t kbytes, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {kbytes, rows};
  cuuint64_t strides[1] = {kbytes};
  cuuint32_t box[2] = {(cuuint32_t)BKB, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

int run(int M, int N, int K, bool verify) {
  EncodeTiledFn enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
  uint32_t seed = 12345;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (int8_t)((int)(seed >> 24) - 128 == -128 ? -127 : (int)(seed >> 24) - 128); };
  for (auto& x : hA) x = rnd();
  for (auto& x : hB) x = rnd();
  int8_t *dA, *dB; int32_t* dC;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
  CUtensorMap mA = make_map(enc, dA, M, K, BM), mB = make_map(enc, dB, N, K, BN);
  const int smem_bytes = STAGES * STAGE_BYTES + 1024;
  CK(cudaFuncSetAttribute(i8gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  dim3 grid(N / BN, M / BM);
  i8gemm_kernel<<<grid, NTHREADS, smem_bytes>>>(mA, mB, dC, M, N, K);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (verify) {
    std::vector<int32_t> hC((size_t)M * N);
    CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int i = 0; i < M; i++)
      for (int j = 0; j < N; j++) {
        long long acc = 0;
        for (int k = 0; k < K; k++) acc += (int)hA[(size_t)i * K + k] * (int)hB[(size_t)j * K + k];
        if ((int32_t)acc != hC[(size_t)i * N + j]) { if (bad < 5) printf("mismatch (%d,%d): got %d want %lld\n", i, j, hC[(size_t)i * N + j], acc); bad++; }
      }
    printf("verify M=%d N=%d K=%d: %lld mismatches of %lld\n", M, N, K, bad, (long long)M * N);
    if (bad) return 1;
  } else {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) i8gemm_kernel<<<grid, NTHREADS, smem_bytes>>>(mA, mB, dC, M, N, K);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; i++) i8gemm_kernel<<<grid, NTHREADS, smem_bytes>>>(mA, mB, dC, M, N, K);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("time M=%d N=%d K=%d: %.3f ms  %.1f TOP/s (2MNK)\n", M, N, K, ms, 2.0 * M * N * K / ms * 1e-9);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return 0;
}

int main() {
  if (run(128, 128, 128, true)) return 1;
  if (run(256, 384, 1024, true)) return 1;
  if (run(4096, 4096, 4096, false)) return 1;
  if (run(8192, 8192, 8192, false)) return 1;
  printf("I8GEMM_OK\n");
  return 0;
}
