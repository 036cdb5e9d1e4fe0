// TMA (cp.async.bulk.tensor) + mbarrier helpers for the FP64 tile pipelines, and the host-side tensor-map encoder.
//
// Operand tiles are fetched as 2-D boxes whose inner extent is 16 doubles = 128 bytes with the hardware's
// 128-byte swizzle: the 16-byte chunk c of box row r lands at chunk position c ^ (r & 7) (stage buffers are
// 1024-byte aligned).  Shared memory is dense (no padding); the m8n8k4 fragment loads stay bank-conflict free
// because the rows a half-warp touches are chosen two apart (gemm_core.cuh, frag_off), which makes the XOR
// term run over all eight chunk positions.
//
// The driver entry point cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint, so the library
// keeps linking against cudart only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace db {
namespace tma {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// 2-D tiled load: box at (c0 = inner coordinate, c1 = row coordinate) -> shared, completion on `bar`.
__device__ __forceinline__ void load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// 3-D tiled load (c0 innermost) -> shared, completion on `bar`.
__device__ __forceinline__ void load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], "
      "[%5];\n" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// Bounded wait: a pipeline bug or a lost transaction must end in a launch failure, never in a hung GPU.
// Returns after the phase completed; traps after ~4 s.
__device__ __forceinline__ void mbar_wait_bounded(unsigned bar, unsigned parity) {
  unsigned ok;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  for (;;) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0xfffu) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) asm volatile("trap;");
    }
  }
}
// One lane of the (converged) warp: the form the compiler recognises as "exactly one thread", so that the
// operands of the TMA / tcgen05 instructions issued under it stay in uniform registers.
__device__ __forceinline__ bool elect_one() {
  unsigned pred = 0;
  asm volatile(
      "{\n .reg .pred px;\n elect.sync _|px, 0xffffffff;\n selp.u32 %0, 1, 0, px;\n}\n"
      : "=r"(pred)
      :
      : "memory");
  return pred != 0;
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Row-major fp64 matrix [rows][cols] with leading dimension ld (doubles); box = 16 doubles x box_rows rows,
// 128-byte swizzle.  Returns false when the driver refuses (the callers turn that into DISTBO_ECUDA).
inline bool encode_f64_rows(CUtensorMap* map, const double* base, uint64_t rows, uint64_t cols, uint64_t ld,
                            uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(double)};
  cuuint32_t box[2] = {16, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// int8 planes [planes][rows][cols] (cols contiguous); box = 64 bytes x box_rows x box_planes, 64-byte swizzle: the
// layout tcgen05.mma reads as K-major SWIZZLE_64B operands, one plane after the other.
inline bool encode_u8_planes(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t planes,
                             uint32_t box_rows, uint32_t box_planes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {cols, rows, planes};
  cuuint64_t strides[2] = {cols, cols * rows};
  cuuint32_t box[3] = {64, box_rows, box_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tma

// tcgen05 (5th-generation tensor cores, accumulators in tensor memory) - the few forms this library issues.
namespace tc5 {

__device__ __forceinline__ void alloc(unsigned smem_slot, unsigned cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_slot), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void dealloc(unsigned taddr, unsigned cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// arrive on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, signed 8-bit operands, 32-bit integer accumulators; K = 32 per instruction.
// COLLECT: 0 = plain; 1 = keep the A tile in the tensor core's collector buffer (collector::a::fill);
//          2 = take A from the collector, last use (collector::a::lastuse) - the A tile is not re-read from
//          shared memory (SASS: UTCIMMA gdesc.A_KEEP / .A_REUSE).
template <int COLLECT>
__device__ __forceinline__ void mma_i8(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                       unsigned idesc, unsigned accumulate) {
  if (COLLECT == 1)
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(
            d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
  else if (COLLECT == 2)
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(
            d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
  else
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns: thread t of the warp gets lane (taddr.lane + t), columns taddr.col .. +7
__device__ __forceinline__ void ld_32x32b_x8(unsigned taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major operand stored as rows of 64 bytes with the 64-byte swizzle
// (what tma::encode_u8_planes delivers): start address, 8-row group stride 512 bytes, descriptor version 1.
__device__ __forceinline__ unsigned long long desc_k_sw64(unsigned smem_addr) {
  return (unsigned long long)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((unsigned long long)(512 >> 4) << 32) |
         (1ull << 46) | (4ull << 61);
}
// instruction descriptor: s8 x s8 -> s32, both operands K-major, M x N tile
__host__ __device__ constexpr unsigned idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

}  // namespace tc5
}  // namespace db
