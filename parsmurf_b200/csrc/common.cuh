// parsmurf_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace psm {

// ---------------------------------------------------------------------------------------------
// Sorted-list entry (one in-bag training row of one tree, in one feature's ascending order):
//   [63] label (1 = negative class, ranger classID 1)   [62:48] bootstrap multiplicity (1..32767)
//   [47:24] dense rank of the row's value in this feature   [23:0] row id inside the partition matrix
// 8 bytes, so split search streams one coalesced u64 per (row, candidate feature) and never gathers.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long entry_t;
constexpr uint32_t kMaxRows = 1u << 24;
constexpr uint32_t kMaxCount = (1u << 15) - 1;

__host__ __device__ __forceinline__ uint32_t e_row(entry_t e) { return (uint32_t)(e & 0xFFFFFFu); }
__host__ __device__ __forceinline__ uint32_t e_rank(entry_t e) { return (uint32_t)((e >> 24) & 0xFFFFFFu); }
__host__ __device__ __forceinline__ uint32_t e_count(entry_t e) { return (uint32_t)((e >> 48) & 0x7FFFu); }
__host__ __device__ __forceinline__ uint32_t e_neg(entry_t e) { return (uint32_t)(e >> 63); }
__host__ __device__ __forceinline__ entry_t e_make(uint32_t row, uint32_t rank, uint32_t count, uint32_t neg) {
	return (entry_t)row | ((entry_t)rank << 24) | ((entry_t)count << 48) | ((entry_t)neg << 63);
}

// Entry codecs.  E = unsigned long long is the general layout above.  E = uint32_t is the compact layout used when a
// partition's training matrix has <= 16384 rows and no bootstrap multiplicity exceeds 15 (all BASELINE configs with
// m = 26): [31:28] multiplicity | [27:14] dense rank | [13:0] row; the label is implied by the row (rows below nPosOut
// are positives).  Halving the entry halves the HBM traffic of split search and of the per-level list partition.
template <typename E> struct Ent;
template <> struct Ent<unsigned long long> {
	static constexpr uint32_t kMaxRowsE = kMaxRows, kMaxCountE = kMaxCount;
	__host__ __device__ static __forceinline__ uint32_t row(unsigned long long e) { return e_row(e); }
	__host__ __device__ static __forceinline__ uint32_t rank(unsigned long long e) { return e_rank(e); }
	__host__ __device__ static __forceinline__ uint32_t count(unsigned long long e) { return e_count(e); }
	__host__ __device__ static __forceinline__ uint32_t neg(unsigned long long e, uint32_t) { return e_neg(e); }
	__host__ __device__ static __forceinline__ unsigned long long make(uint32_t r, uint32_t k, uint32_t c, uint32_t n) { return e_make(r, k, c, n); }
};
template <> struct Ent<uint32_t> {
	static constexpr uint32_t kMaxRowsE = 1u << 14, kMaxCountE = 15;
	__host__ __device__ static __forceinline__ uint32_t row(uint32_t e) { return e & 0x3FFFu; }
	__host__ __device__ static __forceinline__ uint32_t rank(uint32_t e) { return (e >> 14) & 0x3FFFu; }
	__host__ __device__ static __forceinline__ uint32_t count(uint32_t e) { return e >> 28; }
	__host__ __device__ static __forceinline__ uint32_t neg(uint32_t e, uint32_t nPosOut) { return (e & 0x3FFFu) >= nPosOut ? 1u : 0u; }
	__host__ __device__ static __forceinline__ uint32_t make(uint32_t r, uint32_t k, uint32_t c, uint32_t) { return r | (k << 14) | (c << 28); }
};

// Packed 16-byte forest node, the only thing prediction reads:
//   internal: a = split threshold (f64), b = feature id (low 32) | left child id (high 32, >= 1); right = left + 1
//   leaf:     a = fraction of class 0 (positive), b = bits(fraction of class 1) with the sign bit set
struct __align__(16) Node16 {
	double a;
	unsigned long long b;
};
__host__ __device__ __forceinline__ bool node_is_leaf(const Node16& nd) { return (long long)nd.b < 0; }

// per-partition descriptor of one batch
struct PartDesc {
	uint32_t negOffset;   // first index into trngNegIdx
	uint32_t nPosOut;     // (fp+1)*P
	uint32_t nNegOut;
	uint32_t Ntr;         // nPosOut + nNegOut
	uint64_t dOffset;     // element offset of this partition's column-major matrix D [m+1][Ntr]
	uint64_t drawOffset;  // element offset into the draws array
	uint64_t sOffset;     // element offset of the sorted (rank,row) lists S [m][Ntr]
};

// level work item: one open node of one tree
struct Item {
	uint32_t tree;
	uint32_t node;
	uint32_t start;   // segment start inside the tree's per-feature lists
	uint32_t cnt;     // entries in the segment
	uint32_t npos;    // samples (with multiplicity) of class 0
	uint32_t nneg;    // samples of class 1
	uint32_t nPosOut; // rows below this id are positives (label of compact list entries)
	uint32_t pl;      // partition index inside the batch (tree / T), carried from the root item
};

struct SearchRes {
	double score;
	uint32_t pos;     // best split after list position `pos` (relative to segment start)
	uint32_t lpos;    // left class-0 samples
	uint32_t lneg;    // left class-1 samples
	uint32_t pad;
};

struct SplitDec {
	uint32_t split;       // 1 = node splits
	uint32_t nleft;       // entries going left
	uint32_t lpos, lneg;  // left samples per class
};

// what the list partition needs to know about one split node, in one 16-byte load
struct __align__(16) PartTask {
	uint32_t tree;
	uint32_t start;
	uint32_t cnt;     // 0 = nothing to move (no split, or both children terminal); top bits: PT_DEAD_LEFT / PT_DEAD_RIGHT
	uint32_t nleft;
};
constexpr uint32_t PT_DEAD_LEFT = 1u << 31, PT_DEAD_RIGHT = 1u << 30, PT_CNT_MASK = (1u << 24) - 1u;   // rows per matrix < 2^24

// per-tree growth state
struct TreeState {
	uint32_t nNodes;      // nodes created so far
	uint32_t itemBase;    // first item of this tree in the current level's item array
	uint32_t itemCount;   // items of this tree in the current level
	uint32_t mtPos;       // next unread index in the mt19937_64 state block (0..312)
	uint32_t nEntries;    // in-bag distinct rows
	uint32_t overflow;
	uint32_t pad[2];
};

// Opt a kernel in to the full 227 KB of dynamic shared memory.  The attribute belongs to the function, not to the calling
// thread or stream: several host threads (fold lanes) launch the same kernel with different sizes, so every caller sets the
// SAME ceiling instead of its own size -- setting the exact size would let one thread lower it under another's launch.
constexpr int kMaxSmemOptin = 227 * 1024;   // per block, static + dynamic
template <typename F> inline cudaError_t allow_max_dyn_smem(F* fn) {
	cudaFuncAttributes a;
	cudaError_t e = cudaFuncGetAttributes(&a, (const void*)fn);
	if (e != cudaSuccess) return e;
	return cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin - (int)a.sharedSizeBytes);
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
	unsigned m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}
__device__ __forceinline__ unsigned long long ld_nc_u64(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}

// order-preserving u64 key of an f64 (-0.0 canonicalised to +0.0 so == on doubles is == on keys)
__host__ __device__ __forceinline__ unsigned long long f64_key(double v) {
#ifdef __CUDA_ARCH__
	unsigned long long b = (unsigned long long)__double_as_longlong(v);
#else
	unsigned long long b;
	memcpy(&b, &v, 8);
#endif
	if ((b << 1) == 0) b = 0;
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

}  // namespace psm
