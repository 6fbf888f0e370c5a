// K3 feature_rank -- per (partition, feature) ascending order of the training rows + dense ranks.
//
// Stands in for Data::sort (reference src/ranger/Utility/Data.cpp:187-215, reached from
// ForestProbability::initInternal, src/ranger/Forest/ForestProbability.cpp:97-99): ranger stores, per
// column, the sorted unique values and each row's index into them.  Here each column becomes a list of
// (dense rank, row) in ascending value order -- the "sorted-index" form the level-wise split search streams.
// Equal doubles get equal ranks (-0.0 == +0.0 as in std::unique's operator==).
//
// One CTA per (partition, feature) segment: LSD radix sort, 8 passes x 8 bits over the order-preserving
// u64 image of the f64 values, payload = row id, ping-pong buffers in global memory (L2 resident: a segment
// is Ntr*12 B), stable in-tile ranking with __match_any_sync.  Then a block scan turns "value changed" flags
// into dense ranks.  Byte work only -- no tensor cores.
#include <cstdlib>
#include "kernels.cuh"

namespace psm {

constexpr int ST = 256;   // threads per CTA = elements per tile

__global__ void __launch_bounds__(ST) feature_sort_kernel(const double* __restrict__ Dall, const PartDesc* __restrict__ parts,
                                                           uint32_t m, unsigned long long* __restrict__ keysA,
                                                           unsigned long long* __restrict__ keysB, uint32_t* __restrict__ valsA,
                                                           uint32_t* __restrict__ valsB, entry_t* __restrict__ Sall) {
	__shared__ uint32_t base[257];
	__shared__ uint32_t wc[8][257];
	__shared__ uint32_t wsum[8];
	__shared__ uint32_t carry_s;
	const uint32_t seg = blockIdx.x;
	const uint32_t pl = seg / m, f = seg % m;
	const PartDesc pd = parts[pl];
	const uint32_t N = pd.Ntr;
	const size_t so = pd.sOffset + (size_t)f * N;
	const double* src = Dall + pd.dOffset + (size_t)f * N;
	unsigned long long* kin = keysA + so;
	unsigned long long* kout = keysB + so;
	uint32_t* vin = valsA + so;
	uint32_t* vout = valsB + so;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const unsigned lt = lanemask_lt();

	for (uint32_t i = tid; i < N; i += ST) { kin[i] = f64_key(src[i]); vin[i] = i; }
	for (int i = tid; i < 8 * 257; i += ST) (&wc[0][0])[i] = 0;
	__syncthreads();

	const uint32_t ntiles = (N + ST - 1) / ST;
	for (int pass = 0; pass < 8; pass++) {
		const int shift = pass * 8;
		for (int i = tid; i < 257; i += ST) base[i] = 0;
		__syncthreads();
		for (uint32_t i = tid; i < N; i += ST) atomicAdd(&base[(uint32_t)(kin[i] >> shift) & 255u], 1u);
		__syncthreads();
		// exclusive scan of the 256 digit counts (warp 0..7 each scan 32 bins, then add warp offsets)
		{
			uint32_t v = base[tid];
			uint32_t inc = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
			if (lane == 31) wsum[w] = inc;
			__syncthreads();
			uint32_t off = 0;
			for (int ww = 0; ww < w; ww++) off += wsum[ww];
			base[tid] = off + inc - v;
			__syncthreads();
		}
		for (uint32_t t = 0; t < ntiles; t++) {
			const uint32_t i = t * ST + tid;
			const bool valid = i < N;
			unsigned long long key = valid ? kin[i] : 0ull;
			uint32_t val = valid ? vin[i] : 0u;
			const uint32_t d = valid ? ((uint32_t)(key >> shift) & 255u) : 256u;
			const unsigned peers = __match_any_sync(0xffffffffu, d);
			const uint32_t rk = __popc(peers & lt);
			const uint32_t cw = __popc(peers);
			const uint32_t stamp = (uint32_t)pass * ntiles + t + 1;   // unique per (pass, tile); fits: 8*ntiles < 2^20 for N < 2^25
			if (rk == 0) wc[w][d] = (stamp << 6) | cw;                 // cw <= 32
			__syncthreads();
			if (valid) {
				uint32_t off = base[d];
				for (int ww = 0; ww < w; ww++) { uint32_t s = wc[ww][d]; if ((s >> 6) == stamp) off += s & 63u; }
				kout[off + rk] = key;
				vout[off + rk] = val;
			}
			__syncthreads();
			if (rk == 0 && valid) atomicAdd(&base[d], cw);
		}
		__syncthreads();
		unsigned long long* tk = kin; kin = kout; kout = tk;
		uint32_t* tv = vin; vin = vout; vout = tv;
	}
	// 8 passes: result is back in keysA/valsA (== kin/vin).  Dense ranks by block scan of "value changed".
	entry_t* S = Sall + so;
	if (tid == 0) carry_s = 0;
	__syncthreads();
	for (uint32_t t = 0; t < ntiles; t++) {
		const uint32_t i = t * ST + tid;
		uint32_t flag = (i > 0 && i < N) ? (kin[i] != kin[i - 1]) : 0u;
		uint32_t inc = flag;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
		if (lane == 31) wsum[w] = inc;
		__syncthreads();
		uint32_t off = carry_s;
		for (int ww = 0; ww < w; ww++) off += wsum[ww];
		const uint32_t rank = off + inc;
		if (i < N) S[i] = e_make(vin[i], rank, 0, 0);
		__syncthreads();
		if (tid == ST - 1) carry_s = rank;
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------
// Shared-memory path (Ntr <= 16384, i.e. every BASELINE config with m = 26): the whole segment lives in shared memory
// (u64 keys + u16 row ids, up to 160 KB) and is sorted by a bitonic network, so HBM sees the column once (read) and
// the list once (write) instead of 8 radix passes.  Stability is irrelevant: equal values get equal dense ranks and
// split candidates only exist between different ranks.  Padding keys are all-ones.
// ------------------------------------------------------------------------------------------------
constexpr int BT = 1024;

// NPAD = padded segment length (power of two <= 16384); every thread performs NPAD/2/BT compare-exchanges per stage,
// fully unrolled so their shared-memory loads are issued back to back.
template <uint32_t NPAD>
__global__ void __launch_bounds__(BT, 1) feature_sort_smem_kernel(const double* __restrict__ Dall, const PartDesc* __restrict__ parts,
                                                                   uint32_t m, entry_t* __restrict__ Sall, unsigned short* __restrict__ rankTab) {
	extern __shared__ __align__(16) unsigned char smraw[];
	unsigned long long* keys = reinterpret_cast<unsigned long long*>(smraw);            // [NPAD]
	unsigned short* ids = reinterpret_cast<unsigned short*>(smraw + (size_t)NPAD * 8);   // [NPAD]
	__shared__ uint32_t wsum[BT / 32];
	__shared__ uint32_t carry_s;
	const uint32_t seg = blockIdx.x;
	const uint32_t pl = seg / m, f = seg % m;
	const PartDesc pd = parts[pl];
	const uint32_t N = pd.Ntr;
	const double* src = Dall + pd.dOffset + (size_t)f * N;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	for (uint32_t i = tid; i < NPAD; i += BT) {
		keys[i] = (i < N) ? f64_key(src[i]) : 0xFFFFFFFFFFFFFFFFull;
		ids[i] = (unsigned short)i;
	}
	__syncthreads();
	constexpr uint32_t CE = (NPAD / 2 + BT - 1) / BT;   // compare-exchanges per thread per stage
	for (uint32_t k = 2; k <= NPAD; k <<= 1) {
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			unsigned long long a[CE], b[CE];
			uint32_t lo[CE];
#pragma unroll
			for (uint32_t q = 0; q < CE; q++) {
				const uint32_t t = tid + q * BT;
				lo[q] = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
				a[q] = keys[lo[q] & (NPAD - 1)];
				b[q] = keys[(lo[q] | j) & (NPAD - 1)];
			}
#pragma unroll
			for (uint32_t q = 0; q < CE; q++) {
				const uint32_t t = tid + q * BT;
				const bool up = (lo[q] & k) == 0;
				if (t < NPAD / 2 && (a[q] > b[q]) == up && a[q] != b[q]) {
					const uint32_t hi = lo[q] | j;
					keys[lo[q]] = b[q]; keys[hi] = a[q];
					const unsigned short ia = ids[lo[q]]; ids[lo[q]] = ids[hi]; ids[hi] = ia;
				}
			}
			__syncthreads();
		}
	}
	entry_t* S = Sall + pd.sOffset + (size_t)f * N;
	if (tid == 0) carry_s = 0;
	__syncthreads();
	for (uint32_t t0 = 0; t0 < N; t0 += BT) {
		const uint32_t i = t0 + tid;
		uint32_t flag = (i > 0 && i < N) ? (keys[i] != keys[i - 1]) : 0u;
		uint32_t inc = flag;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
		if (lane == 31) wsum[w] = inc;
		__syncthreads();
		uint32_t off = carry_s;
		for (int ww = 0; ww < w; ww++) off += wsum[ww];
		const uint32_t rank = off + inc;
		if (i < N) {
			S[i] = e_make(ids[i], rank, 0, 0);
			if (rankTab) rankTab[pd.sOffset + (size_t)f * N + ids[i]] = (unsigned short)rank;
		}
		__syncthreads();
		if (tid == BT - 1) carry_s = rank;
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------
// Split variant: a bitonic network needs a power-of-two length, and padding Ntr = 9600 (BASELINE configs[1]) to 16384 wastes
// 41% of every stage.  Here the segment is cut into run A = the largest power of two <= N and run B = the rest (padded to
// its own power of two Bp <= A).  Both runs go through the same network -- stages with k <= Bp cover A and B, later ones
// only A -- and a merge-path pass (A first on ties) fuses the two sorted runs, the dense ranks and the entry write-out.
// For N = 9600: 10240 slots instead of 16384, 91 stages of 4-5 compare-exchanges per thread instead of 105 of 8.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pow2_floor(uint32_t v) { return v ? 1u << (31 - __clz(v)) : 0u; }
__device__ __host__ inline uint32_t split_slots(uint32_t N) {
	if (N == 0) return 0;
	uint32_t A = 1; while (A * 2 <= N) A *= 2;
	uint32_t nb = N - A, Bp = 0;
	if (nb) { Bp = 1; while (Bp < nb) Bp *= 2; }
	return A + Bp;
}

template <int CEMAX, int MINB>
__global__ void __launch_bounds__(BT, MINB) feature_sort_split_kernel(const double* __restrict__ Dall, const PartDesc* __restrict__ parts,
                                                                       uint32_t m, uint32_t slotsMax, entry_t* __restrict__ Sall,
                                                                       const uint32_t* __restrict__ onlyFlagged, unsigned short* __restrict__ rankTab) {
	extern __shared__ __align__(16) unsigned char smraw[];
	unsigned long long* keys = reinterpret_cast<unsigned long long*>(smraw);                 // [slotsMax]
	unsigned short* ids = reinterpret_cast<unsigned short*>(smraw + (size_t)slotsMax * 8);    // [slotsMax]
	__shared__ uint32_t wsum[BT / 32];
	const uint32_t seg = blockIdx.x;
	if (onlyFlagged && onlyFlagged[seg] == 0u) return;     // second pass after the radix kernel: only the segments it gave up on
	const uint32_t pl = seg / m, f = seg % m;
	const PartDesc pd = parts[pl];
	const uint32_t N = pd.Ntr;
	if (N == 0) return;
	const double* src = Dall + pd.dOffset + (size_t)f * N;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const uint32_t A = pow2_floor(N), nb = N - A;
	const uint32_t Bp = nb ? (nb > 1 ? 1u << (32 - __clz(nb - 1)) : 1u) : 0u;
	const uint32_t slots = A + Bp;                            // run B starts at slot A == source index A
	for (uint32_t i = tid; i < slots; i += BT) {
		keys[i] = (i < N) ? f64_key(src[i]) : 0xFFFFFFFFFFFFFFFFull;
		ids[i] = (unsigned short)i;
	}
	__syncthreads();
	for (uint32_t k = 2; k <= A; k <<= 1) {
		const uint32_t count = (k <= Bp) ? slots / 2 : A / 2;   // compare-exchanges of this stage
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			unsigned long long a[CEMAX], b[CEMAX];
			uint32_t lo[CEMAX];
#pragma unroll
			for (int q = 0; q < CEMAX; q++) {
				const uint32_t t = tid + q * BT;
				lo[q] = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
				if (t < count) { a[q] = keys[lo[q]]; b[q] = keys[lo[q] | j]; }
			}
#pragma unroll
			for (int q = 0; q < CEMAX; q++) {
				const uint32_t t = tid + q * BT;
				const bool up = ((lo[q] & k) == 0) | (k == A);   // k == A is the last stage of both runs (Bp == A): both end ascending
				if (t < count && (a[q] > b[q]) == up && a[q] != b[q]) {
					const uint32_t hi = lo[q] | j;
					keys[lo[q]] = b[q]; keys[hi] = a[q];
					const unsigned short ia = ids[lo[q]]; ids[lo[q]] = ids[hi]; ids[hi] = ia;
				}
			}
			__syncthreads();
		}
	}
	// merge-path: thread t produces outputs [o0, o1) of the merged order; i = elements taken from run A, o - i from run B
	const unsigned long long* kA = keys;
	const unsigned long long* kB = keys + A;
	const uint32_t ipt = (N + BT - 1) / BT;
	const uint32_t o0 = min(N, (uint32_t)tid * ipt), o1 = min(N, o0 + ipt);
	uint32_t i0;
	{
		uint32_t lo2 = o0 > nb ? o0 - nb : 0u, hi2 = min(o0, A);
		while (lo2 < hi2) {
			const uint32_t mid = (lo2 + hi2) >> 1;
			if (kA[mid] <= kB[o0 - mid - 1]) lo2 = mid + 1; else hi2 = mid;
		}
		i0 = lo2;
	}
	const uint32_t j0 = o0 - i0;
	unsigned long long prev0 = 0;                                // key of output o0 - 1 (the larger of the two last-taken keys)
	if (o0 > 0) {
		const unsigned long long pa = i0 ? kA[i0 - 1] : 0ull, pb = j0 ? kB[j0 - 1] : 0ull;
		prev0 = pa > pb ? pa : pb;
	}
	// pass 1: number of value changes inside this thread's range
	uint32_t changes = 0;
	{
		uint32_t i = i0, j = j0;
		unsigned long long prev = prev0;
		for (uint32_t o = o0; o < o1; o++) {
			const unsigned long long ka = i < A ? kA[i] : 0xFFFFFFFFFFFFFFFFull, kb = j < nb ? kB[j] : 0xFFFFFFFFFFFFFFFFull;
			const bool takeA = i < A && (j >= nb || ka <= kb);
			const unsigned long long key = takeA ? ka : kb;
			i += takeA ? 1u : 0u; j += takeA ? 0u : 1u;
			changes += (o > 0 && key != prev) ? 1u : 0u;
			prev = key;
		}
	}
	uint32_t inc = changes;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	uint32_t rank = inc - changes;
	for (int ww = 0; ww < w; ww++) rank += wsum[ww];
	// pass 2: the same walk, now writing (row, dense rank)
	entry_t* S = Sall + pd.sOffset + (size_t)f * N;
	{
		uint32_t i = i0, j = j0;
		unsigned long long prev = prev0;
		for (uint32_t o = o0; o < o1; o++) {
			const unsigned long long ka = i < A ? kA[i] : 0xFFFFFFFFFFFFFFFFull, kb = j < nb ? kB[j] : 0xFFFFFFFFFFFFFFFFull;
			const bool takeA = i < A && (j >= nb || ka <= kb);
			const unsigned long long key = takeA ? ka : kb;
			const uint32_t id = takeA ? ids[i] : ids[A + j];
			i += takeA ? 1u : 0u; j += takeA ? 0u : 1u;
			rank += (o > 0 && key != prev) ? 1u : 0u;
			prev = key;
			S[o] = e_make(id, rank, 0, 0);
			if (rankTab) rankTab[pd.sOffset + (size_t)f * N + id] = (unsigned short)rank;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Radix variant (r02; the default for Ntr <= 16384).  The bitonic network above costs Ntr*log^2(Ntr) compare-exchanges:
// 2.0e9 warp instructions per BASELINE configs[1] CV, 76% issue-active (profiles/r01_summary.md).  Here the segment is sorted by
// the HIGH word of the order-preserving u64 image of the values (sign, exponent, 20 mantissa bits) with a stable LSD radix sort
// in shared memory -- 4 passes of 8 bits, every warp owning a contiguous slice of the segment and private digit counters, ranks
// inside a warp from __match_any_sync -- and the few neighbours that agree in the high word are then put in order by the LOW
// word with an odd-even transposition that only ever exchanges elements of one such run (two quiet phases end it).  Runs that
// are still moving after kFixIters phases (thousands of values that differ only below 2^-20 relative: adversarial) flag the
// segment, and the bitonic kernel redoes exactly the flagged segments.
// Shared memory: 12 bytes per element (u32 key + u16 row id, ping-pong) + 32 KB of counters.
// ------------------------------------------------------------------------------------------------
constexpr int RXT = 1024, RXW = RXT / 32;
constexpr int kFixIters = 48;

__global__ void __launch_bounds__(RXT, 1) feature_sort_radix_kernel(const double* __restrict__ Dall, const PartDesc* __restrict__ parts, uint32_t m,
                                                                     uint32_t maxNp, entry_t* __restrict__ Sall, uint32_t* __restrict__ slowFlags,
                                                                     unsigned short* __restrict__ rankTab) {
	extern __shared__ __align__(16) unsigned char smraw[];
	uint32_t* kA = reinterpret_cast<uint32_t*>(smraw);                 // [maxNp]
	uint32_t* kB = kA + maxNp;                                         // [maxNp]
	unsigned short* idA = reinterpret_cast<unsigned short*>(kB + maxNp);   // [maxNp]
	unsigned short* idB = idA + maxNp;                                 // [maxNp]
	uint32_t* wc = reinterpret_cast<uint32_t*>(idB + maxNp);           // [RXW][256] per-warp digit counters -> offsets inside the digit
	__shared__ uint32_t base[256];
	__shared__ uint32_t wsum[RXW];
	__shared__ uint32_t carry_s;
	const uint32_t seg = blockIdx.x;
	const uint32_t pl = seg / m, f = seg % m;
	const PartDesc pd = parts[pl];
	const uint32_t N = pd.Ntr;
	if (N == 0) return;
	const double* src = Dall + pd.dOffset + (size_t)f * N;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const unsigned lt = lanemask_lt();
	for (uint32_t i = tid; i < N; i += RXT) { kA[i] = (uint32_t)(f64_key(src[i]) >> 32); idA[i] = (unsigned short)i; }
	// every warp owns `chunk` consecutive elements (a multiple of 32, so that all its rounds but the last are full)
	const uint32_t chunk = (((N + RXW - 1) / RXW) + 31u) & ~31u;
	const uint32_t c0 = min(N, (uint32_t)w * chunk), c1 = min(N, c0 + chunk);
	uint32_t* myc = wc + w * 256;
	for (int pass = 0; pass < 4; pass++) {
		const int shift = pass * 8;
		for (int i = tid; i < RXW * 256; i += RXT) wc[i] = 0;
		__syncthreads();
		for (uint32_t i = c0 + lane; i < c1; i += 32) atomicAdd(&myc[(kA[i] >> shift) & 255u], 1u);
		__syncthreads();
		if (tid < 256) {
			// digit tid: counts of the 32 warps -> exclusive offsets inside the digit, then the digit's base
			uint32_t run = 0;
#pragma unroll 8
			for (int ww = 0; ww < RXW; ww++) { const uint32_t c = wc[ww * 256 + tid]; wc[ww * 256 + tid] = run; run += c; }
			uint32_t inc = run;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
			if (lane == 31) wsum[w] = inc;
			base[tid] = inc - run;                       // exclusive inside the warp of 32 digits
		}
		__syncthreads();
		if (tid < 256) { uint32_t off = 0; for (int ww = 0; ww < w; ww++) off += wsum[ww]; base[tid] += off; }
		__syncthreads();
		for (uint32_t i0 = c0; i0 < c1; i0 += 32) {
			const uint32_t i = i0 + lane;
			const bool valid = i < c1;
			const uint32_t key = valid ? kA[i] : 0u;
			const unsigned short id = valid ? idA[i] : (unsigned short)0;
			const uint32_t d = valid ? ((key >> shift) & 255u) : (256u + (uint32_t)lane);   // invalid lanes match nobody
			const unsigned peers = __match_any_sync(0xffffffffu, d);
			const uint32_t rk = __popc(peers & lt);
			uint32_t pos = 0;
			if (valid) pos = base[d] + myc[d] + rk;
			__syncwarp();
			if (valid && rk == 0) myc[d] += __popc(peers);
			__syncwarp();
			if (valid) { kB[pos] = key; idB[pos] = id; }
		}
		__syncthreads();
		{ uint32_t* tk = kA; kA = kB; kB = tk; unsigned short* ti = idA; idA = idB; idB = ti; }
	}
	// sorted by the high word (back in the first pair of arrays); low words of the sorted order into the free key array
	for (uint32_t i = tid; i < N; i += RXT) kB[i] = (uint32_t)f64_key(src[idA[i]]);
	__syncthreads();
	// odd-even transposition restricted to runs of equal high words
	bool settled = false;
	{
		int quiet = 0;
		for (int it = 0; it < kFixIters; it++) {
			const uint32_t phase = (uint32_t)it & 1u;
			int moved = 0;
			for (uint32_t i = 2u * tid + phase; i + 1 < N; i += 2u * RXT) {
				if (kA[i] == kA[i + 1] && kB[i] > kB[i + 1]) {
					const uint32_t t = kB[i]; kB[i] = kB[i + 1]; kB[i + 1] = t;
					const unsigned short u = idA[i]; idA[i] = idA[i + 1]; idA[i + 1] = u;
					moved = 1;
				}
			}
			if (__syncthreads_or(moved)) quiet = 0;
			else if (++quiet == 2) { settled = true; break; }
		}
	}
	if (!settled) {                                   // uniform: every thread saw the same votes
		if (tid == 0) slowFlags[seg] = 1u;
		return;
	}
	// dense ranks (equal doubles share a rank) and the (rank, row) entries
	entry_t* S = Sall + pd.sOffset + (size_t)f * N;
	if (tid == 0) carry_s = 0;
	__syncthreads();
	for (uint32_t t0 = 0; t0 < N; t0 += RXT) {
		const uint32_t i = t0 + tid;
		const uint32_t flag = (i > 0 && i < N) ? ((kA[i] != kA[i - 1] || kB[i] != kB[i - 1]) ? 1u : 0u) : 0u;
		uint32_t inc = flag;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
		if (lane == 31) wsum[w] = inc;
		__syncthreads();
		uint32_t off = carry_s;
		for (int ww = 0; ww < w; ww++) off += wsum[ww];
		const uint32_t rank = off + inc;
		if (i < N) {
			S[i] = e_make(idA[i], rank, 0, 0);
			if (rankTab) rankTab[pd.sOffset + (size_t)f * N + idA[i]] = (unsigned short)rank;
		}
		__syncthreads();
		if (tid == RXT - 1) carry_s = rank;
		__syncthreads();
	}
}

template <int CEMAX, int MINB>
static int launch_sort_split(const double* Dall, const PartDesc* parts, uint32_t nPartsBatch, uint32_t m, uint32_t slotsMax, entry_t* Sall,
                             const uint32_t* onlyFlagged, unsigned short* rankTab, cudaStream_t s) {
	const size_t smem = (size_t)slotsMax * 10;
	if (allow_max_dyn_smem(feature_sort_split_kernel<CEMAX, MINB>) != cudaSuccess) return -1;
	feature_sort_split_kernel<CEMAX, MINB><<<nPartsBatch * m, BT, smem, s>>>(Dall, parts, m, slotsMax, Sall, onlyFlagged, rankTab);
	return 0;
}

template <uint32_t NPAD>
static int launch_sort_smem_N(const double* Dall, const PartDesc* parts, uint32_t nPartsBatch, uint32_t m, entry_t* Sall, unsigned short* rankTab,
                             cudaStream_t s) {
	const size_t smem = (size_t)NPAD * 10;
	if (allow_max_dyn_smem(feature_sort_smem_kernel<NPAD>) != cudaSuccess) return -1;
	feature_sort_smem_kernel<NPAD><<<nPartsBatch * m, BT, smem, s>>>(Dall, parts, m, Sall, rankTab);
	return 0;
}

int launch_feature_sort_smem(const double* Dall, const PartDesc* parts, uint32_t nPartsBatch, uint32_t m, uint32_t maxNtr, entry_t* Sall,
                             uint32_t* slowFlags, unsigned short* rankTab, cudaStream_t s) {
	if (maxNtr > 16384) return 1;   // caller falls back to the global-memory radix sort
	const uint32_t* onlyFlagged = nullptr;
	if (slowFlags && !getenv("PSM_SORT_BITONIC")) {
		const uint32_t maxNp = (maxNtr + 3u) & ~3u;
		const size_t smemR = (size_t)maxNp * 12 + (size_t)RXW * 256 * 4;
		if (allow_max_dyn_smem(feature_sort_radix_kernel) != cudaSuccess) return -1;
		cudaMemsetAsync(slowFlags, 0, (size_t)nPartsBatch * m * 4, s);
		feature_sort_radix_kernel<<<nPartsBatch * m, RXT, smemR, s>>>(Dall, parts, m, maxNp, Sall, slowFlags, rankTab);
		onlyFlagged = slowFlags;       // the bitonic kernel below then only redoes what the radix kernel flagged
	}
	if (!getenv("PSM_SORT_POW2") || onlyFlagged) {
		const uint32_t slotsMax = split_slots(maxNtr);        // monotone in N, so every segment of the batch fits
		const int ce = (int)((slotsMax / 2 + BT - 1) / BT);
		const bool two = (size_t)slotsMax * 10 <= 110 * 1024 && !getenv("PSM_SORT_ONE_CTA");   // two CTAs per SM when shared memory allows
#define PSM_SORT_CASE(C)                                                                                        \
	case C: return two ? launch_sort_split<C, 2>(Dall, parts, nPartsBatch, m, slotsMax, Sall, onlyFlagged, rankTab, s)   \
	                   : launch_sort_split<C, 1>(Dall, parts, nPartsBatch, m, slotsMax, Sall, onlyFlagged, rankTab, s);
		switch (ce < 1 ? 1 : ce) {
			PSM_SORT_CASE(1) PSM_SORT_CASE(2) PSM_SORT_CASE(3) PSM_SORT_CASE(4) PSM_SORT_CASE(5) PSM_SORT_CASE(6) PSM_SORT_CASE(7) PSM_SORT_CASE(8)
		}
#undef PSM_SORT_CASE
	}
	if (maxNtr <= 2048) return launch_sort_smem_N<2048>(Dall, parts, nPartsBatch, m, Sall, rankTab, s);
	if (maxNtr <= 4096) return launch_sort_smem_N<4096>(Dall, parts, nPartsBatch, m, Sall, rankTab, s);
	if (maxNtr <= 8192) return launch_sort_smem_N<8192>(Dall, parts, nPartsBatch, m, Sall, rankTab, s);
	return launch_sort_smem_N<16384>(Dall, parts, nPartsBatch, m, Sall, rankTab, s);
}

void launch_feature_sort(const double* Dall, const PartDesc* parts, uint32_t nPartsBatch, uint32_t m, unsigned long long* keysA,
                         unsigned long long* keysB, uint32_t* valsA, uint32_t* valsB, entry_t* Sall, cudaStream_t s) {
	feature_sort_kernel<<<nPartsBatch * m, ST, 0, s>>>(Dall, parts, m, keysA, keysB, valsA, valsB, Sall);
}

}  // namespace psm
