// K7 curves -- AUROC / AUPRC of (labels, scores) as the reference's Curves class computes them
// (reference src/curves.cpp:7-36 ctor: descending sort; :103-122 evalAUROC_ok; :77-100 evalAUPRC;
//  src/curves.h:73-128 cumulSum / traps_integrate).
//
// Faithfully kept quirks: no tie grouping; FP[last] is overwritten with totN (curves.cpp:87,112); the PR curve
// gets a prepended (recall 0, precision 1) point (:92-93); ROC terms are formed in *float* and summed in double
// (curves.h:119-128 with T=float), PR terms in double; NaN precisions skip their trapezoids.
//
// Pipeline: [device radix sort by descending score | host permutation] -> gather labels -> device-wide inclusive
// scan (TP) -> per-element trapezoid terms -> deterministic two-level fp64 reduction.  HBM-bound byte work.
#include "kernels.cuh"

#include <algorithm>

namespace psm {

constexpr int CB = 256;          // threads
constexpr int CI = 16;           // items per thread
constexpr int CTILE = CB * CI;   // elements per block

__global__ void gather_labels_kernel(const uint32_t* __restrict__ labels, const uint32_t* __restrict__ perm, uint64_t N,
                                     unsigned char* __restrict__ out) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) out[i] = (unsigned char)labels[perm[i]];
}
void launch_gather_labels(const uint32_t* labels, const uint32_t* perm, uint64_t N, unsigned char* out, cudaStream_t s) {
	if (!N) return;
	gather_labels_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(labels, perm, N, out);
}

// fold-local labels: out[i] = labels[idx[i]]
__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t N, uint32_t* __restrict__ out) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) out[i] = src[idx[i]];
}
void launch_gather_u32(const uint32_t* src, const uint32_t* idx, uint64_t N, uint32_t* out, cudaStream_t s) {
	if (!N) return;
	gather_u32_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(src, idx, N, out);
}
// example ids handed in through the C ABI: *bad becomes non-zero if any id is >= n (checked on the device: the arrays are
// already there, and a host loop over the 14M ids of BASELINE configs[2] would cost more than the fold set-up itself)
__global__ void check_ids_kernel(const uint32_t* __restrict__ ids, uint64_t N, uint32_t n, uint32_t* __restrict__ bad) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N && ids[i] >= n) atomicOr(bad, 1u);
}
void launch_check_ids(const uint32_t* ids, uint64_t N, uint32_t n, uint32_t* bad, cudaStream_t s) {
	if (!N) return;
	check_ids_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(ids, N, n, bad);
}
__global__ void gather_f64_kernel(const double* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t N, double* __restrict__ out) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) out[i] = src[idx[i]];
}
void launch_gather_f64(const double* src, const uint32_t* idx, uint64_t N, double* out, cudaStream_t s) {
	if (!N) return;
	gather_f64_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(src, idx, N, out);
}
// classXProb[testIdx[i]] += fold scores (sampler.cpp:121-132 scatter by original example id)
__global__ void scatter_scores_kernel(const double* __restrict__ class12, const uint32_t* __restrict__ testIdx, uint32_t Nte, uint32_t n,
                                      double* __restrict__ glob) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Nte) return;
	const uint32_t t = testIdx[i];
	glob[t] = __dadd_rn(glob[t], class12[i]);
	glob[(size_t)n + t] = __dadd_rn(glob[(size_t)n + t], class12[(size_t)Nte + i]);
}
void launch_scatter_scores(const double* class12, const uint32_t* testIdx, uint32_t Nte, uint32_t n, double* glob, cudaStream_t s) {
	if (!Nte) return;
	scatter_scores_kernel<<<(Nte + 255) / 256, 256, 0, s>>>(class12, testIdx, Nte, n, glob);
}

__global__ void add_f64_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t count) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) dst[i] = __dadd_rn(dst[i], src[i]);
}
void launch_add_f64(double* dst, const double* src, size_t count, cudaStream_t s) {
	if (!count) return;
	add_f64_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(dst, src, count);
}

__device__ __forceinline__ uint32_t block_reduce_u32(uint32_t v, uint32_t* sh) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
	__syncthreads();
	uint32_t t = 0;
	for (int w = 0; w < CB / 32; w++) t += sh[w];
	__syncthreads();
	return t;
}

__global__ void __launch_bounds__(CB) tp_blocksum_kernel(const unsigned char* __restrict__ l, uint64_t N, uint32_t* __restrict__ blockSums) {
	__shared__ uint32_t sh[CB / 32];
	uint64_t b0 = (uint64_t)blockIdx.x * CTILE;
	uint32_t v = 0;
	for (int k = 0; k < CI; k++) { uint64_t i = b0 + (uint64_t)k * CB + threadIdx.x; if (i < N) v += l[i]; }
	uint32_t t = block_reduce_u32(v, sh);
	if (threadIdx.x == 0) blockSums[blockIdx.x] = t;
}

// single block: exclusive scan of `a` in place.  Every thread owns a contiguous run of counters (serial sum, one block scan of the
// 1024 run totals, serial write-back): three barriers in all.  (The first version walked the array 1024 counters at a time with
// two barriers per step -- 33 us for the 63k digit counters of a 1M-key radix pass, 78 such launches per BASELINE configs[1] CV.)
__global__ void __launch_bounds__(1024) scan_blocksums_kernel(uint32_t* __restrict__ a, uint32_t n) {
	__shared__ uint32_t wsum[32];
	const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const uint32_t per = (n + 1023) / 1024;
	const uint32_t b = min(tid * per, n), e = min(b + per, n);
	uint32_t tot = 0;
	for (uint32_t i = b; i < e; i++) tot += a[i];
	uint32_t inc = tot;
	for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	if (w == 0) {
		const uint32_t v = wsum[lane];
		uint32_t s = v;
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= (uint32_t)o) s += t; }
		wsum[lane] = s - v;                                // exclusive prefix of the warp totals
	}
	__syncthreads();
	uint32_t off = wsum[w] + inc - tot;
	for (uint32_t i = b; i < e; i++) { const uint32_t v = a[i]; a[i] = off; off += v; }
}

// TP[i] (inclusive) for a block's tile, element order i = b0 + t*CI + k (thread-contiguous so the scan is a
// per-thread serial prefix + one block scan of thread totals)
__global__ void __launch_bounds__(CB) curve_terms_kernel(const unsigned char* __restrict__ l, uint64_t N, uint32_t totP, uint32_t totN,
                                                          const uint32_t* __restrict__ blockOffsets, double* __restrict__ partials) {
	__shared__ uint32_t wsum[CB / 32];
	__shared__ double dsum[2][CB / 32];
	const uint64_t b0 = (uint64_t)blockIdx.x * CTILE;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	// this kernel counted block sums with a strided mapping; here use the contiguous mapping for the scan
	uint32_t loc[CI + 1];
	uint32_t run = 0;
	const uint64_t t0 = b0 + (uint64_t)tid * CI;
#pragma unroll
	for (int k = 0; k < CI + 1; k++) {       // one element of look-ahead for the i+1 terms
		uint64_t i = t0 + k;
		uint32_t v = (i < N) ? l[i] : 0;
		run += v;
		loc[k] = run;
	}
	uint32_t mine = loc[CI - 1], inc = mine;
	for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	uint32_t off = blockOffsets[blockIdx.x] + inc - mine;
	for (int ww = 0; ww < w; ww++) off += wsum[ww];
	const float fP = __uint2float_rn(totP), fN = __uint2float_rn(totN);
	const double dP = (double)totP;
	double roc = 0.0, pr = 0.0;
#pragma unroll
	for (int k = 0; k < CI; k++) {
		const uint64_t i = t0 + k;
		if (i >= N) break;
		const uint32_t tp0 = off + loc[k];
		uint32_t fp0 = (uint32_t)(i + 1) - tp0;
		if (i == N - 1) fp0 = totN;
		// PR segment ending at point i+1 (points: 0 = (0,1); j+1 = (rec_j, prec_j))
		double precPrev, recPrev;
		if (i == 0) { precPrev = 1.0; recPrev = 0.0; }
		else {
			const uint32_t tpm = tp0 - l[i];
			const uint32_t fpm = (uint32_t)i - tpm;
			precPrev = __ddiv_rn((double)tpm, (double)(tpm + fpm));
			recPrev = __ddiv_rn((double)tpm, dP);
		}
		const double prec = __ddiv_rn((double)tp0, (double)(tp0 + fp0));
		const double rec = __ddiv_rn((double)tp0, dP);
		if (!(isnan(precPrev) || isnan(prec)))
			pr += __dmul_rn(__dmul_rn(__dadd_rn(precPrev, prec), __dsub_rn(rec, recPrev)), 0.5);
		// ROC segment from i to i+1
		if (i + 1 < N) {
			const uint32_t tp1 = off + loc[k + 1];
			uint32_t fp1 = (uint32_t)(i + 2) - tp1;
			if (i + 1 == N - 1) fp1 = totN;
			const float x0 = __fdiv_rn(__uint2float_rn(fp0), fN), x1 = __fdiv_rn(__uint2float_rn(fp1), fN);
			const float y0 = __fdiv_rn(__uint2float_rn(tp0), fP), y1 = __fdiv_rn(__uint2float_rn(tp1), fP);
			if (!(isnan(y0) || isnan(y1)))
				roc += __dmul_rn((double)__fmul_rn(__fadd_rn(y0, y1), __fsub_rn(x1, x0)), 0.5);
		}
	}
	for (int o = 16; o > 0; o >>= 1) { roc += __shfl_xor_sync(0xffffffffu, roc, o); pr += __shfl_xor_sync(0xffffffffu, pr, o); }
	if (lane == 0) { dsum[0][w] = roc; dsum[1][w] = pr; }
	__syncthreads();
	if (tid == 0) {
		double a = 0, b = 0;
		for (int ww = 0; ww < CB / 32; ww++) { a += dsum[0][ww]; b += dsum[1][ww]; }
		partials[2 * (size_t)blockIdx.x] = a;
		partials[2 * (size_t)blockIdx.x + 1] = b;
	}
}

__global__ void __launch_bounds__(256) curve_final_kernel(const double* __restrict__ partials, uint32_t nBlocks, double* __restrict__ out2) {
	__shared__ double sh[2][256];
	double a = 0, b = 0;
	// fixed assignment: thread t sums blocks t, t+256, ... in order -> deterministic
	for (uint32_t i = threadIdx.x; i < nBlocks; i += 256) { a += partials[2 * (size_t)i]; b += partials[2 * (size_t)i + 1]; }
	sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s) { sh[0][threadIdx.x] += sh[0][threadIdx.x + s]; sh[1][threadIdx.x] += sh[1][threadIdx.x + s]; }
		__syncthreads();
	}
	if (threadIdx.x == 0) { out2[0] = sh[0][0]; out2[1] = sh[1][0]; }
}

// tie_mode 2 ("auto"): after the stable device sort, count the adjacent pairs with EQUAL score and DIFFERENT label.  A run of
// equal scores holds both labels iff it contains such a pair; if there is none, every run of tied scores is single-label, the
// order inside it cannot change any TP/FP prefix, and the curves are provably those of the reference's std::sort whatever its
// tie order (src/curves.cpp:31-33).  Otherwise the caller redoes the permutation with the identical host std::sort call.
__global__ void __launch_bounds__(256) tie_mixed_count_kernel(const unsigned long long* __restrict__ keysSorted, const unsigned char* __restrict__ l,
                                                               uint64_t N, unsigned long long* __restrict__ counter) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool mixed = (i + 1 < N) && keysSorted[i] == keysSorted[i + 1] && l[i] != l[i + 1];
	const unsigned b = __ballot_sync(0xffffffffu, mixed);
	if (b && (threadIdx.x & 31) == 0) atomicAdd(counter, (unsigned long long)__popc(b));
}
// tie_mode 2, lower end of the bracket: the labels in "ties ordered negatives first" order, derived from the positives-first
// order without a second sort.  Inside a run of equal scores the positives-first order is [positives][negatives]; negatives
// first is the run reversed, so  out[i] = l[s + e - 1 - i]  with [s, e) the run of element i (two binary searches in the
// sorted keys, only for elements that have an equal neighbour).  Replaces 30 launches of a second radix sort per curves call.
__global__ void __launch_bounds__(256) tie_reverse_labels_kernel(const unsigned long long* __restrict__ keysSorted, const unsigned char* __restrict__ l,
                                                                  uint64_t N, unsigned char* __restrict__ out) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= N) return;
	const unsigned long long k = keysSorted[i];
	const bool tiedL = i > 0 && keysSorted[i - 1] == k, tiedR = i + 1 < N && keysSorted[i + 1] == k;
	if (!tiedL && !tiedR) { out[i] = l[i]; return; }
	uint64_t lo = 0, hi = i;                               // s = first index whose key equals k (the keys ascend)
	while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (keysSorted[mid] < k) lo = mid + 1; else hi = mid; }
	const uint64_t s = lo;
	lo = i + 1; hi = N;                                    // e = first index past i whose key exceeds k
	while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if (keysSorted[mid] <= k) lo = mid + 1; else hi = mid; }
	out[i] = l[s + lo - 1 - i];
}
void launch_tie_reverse_labels(const unsigned long long* keysSorted, const unsigned char* labelsSorted, uint64_t N, unsigned char* out, cudaStream_t s) {
	if (!N) return;
	tie_reverse_labels_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(keysSorted, labelsSorted, N, out);
}
void launch_tie_mixed_count(const unsigned long long* keysSorted, const unsigned char* labelsSorted, uint64_t N, unsigned long long* counter, cudaStream_t s) {
	cudaMemsetAsync(counter, 0, 8, s);
	if (N < 2) return;
	tie_mixed_count_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(keysSorted, labelsSorted, N, counter);
}

// block sums must use the same element->block mapping as curve_terms_kernel (contiguous CTILE per block): they do.
int launch_curves(const unsigned char* labelsSorted, uint64_t N, uint32_t totP, uint32_t totN, uint32_t* tpScratch, double* partials,
                  size_t partialsCap, double* out2, cudaStream_t s) {
	if (N == 0) return -1;
	const uint32_t nBlocks = (uint32_t)((N + CTILE - 1) / CTILE);
	if ((size_t)nBlocks * 2 > partialsCap) return -2;
	tp_blocksum_kernel<<<nBlocks, CB, 0, s>>>(labelsSorted, N, tpScratch);
	scan_blocksums_kernel<<<1, 1024, 0, s>>>(tpScratch, nBlocks);
	curve_terms_kernel<<<nBlocks, CB, 0, s>>>(labelsSorted, N, totP, totN, tpScratch, partials);
	curve_final_kernel<<<1, 256, 0, s>>>(partials, nBlocks, out2);
	return 0;
}

// ------------------------------------------------------------------------------------------------
// Device-wide stable LSD radix sort of scores in DESCENDING order (tie_mode 1): key = ~f64_key(score),
// payload = original index; 8 passes x 8 bits; per pass: per-block digit histogram -> scan -> stable scatter.
// ------------------------------------------------------------------------------------------------
constexpr int RB = 256;
constexpr int RTILES = 16;                 // tiles of RB elements per block
constexpr int RCHUNK = RB * RTILES;

// Per-pass digit scan: hist is [256 digits][nBlocks] block counts.  One warp per digit turns its row into exclusive prefixes
// (coalesced 32-counter steps with a running carry) and leaves the digit's total in hist[256 * nBlocks + digit]; the scatter
// kernel adds the exclusive scan of those 256 totals itself.  (A single 1024-thread block scanning all 256 * nBlocks counters
// took 33 us per pass at 1M keys -- 84 such launches per BASELINE configs[1] CV, profiles/r01_summary.md.)
__global__ void __launch_bounds__(256) radix_scan_digits_kernel(uint32_t* __restrict__ hist, uint32_t nBlocks) {
	const uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (d >= 256) return;
	const uint32_t lane = threadIdx.x & 31;
	uint32_t* row = hist + (size_t)d * nBlocks;
	uint32_t carry = 0;
	for (uint32_t b = 0; b < nBlocks; b += 32) {
		const uint32_t i = b + lane;
		const uint32_t v = i < nBlocks ? row[i] : 0u;
		uint32_t inc = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
		if (i < nBlocks) row[i] = carry + inc - v;
		carry += __shfl_sync(0xffffffffu, inc, 31);
	}
	if (lane == 0) hist[(size_t)256 * nBlocks + d] = carry;
}

__global__ void sort_init_kernel(const double* __restrict__ scores, uint64_t N, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) { keys[i] = ~f64_key(scores[i]); vals[i] = (uint32_t)i; }
}

__global__ void __launch_bounds__(RB) radix_hist_kernel(const unsigned long long* __restrict__ keys, uint64_t N, int shift, uint32_t nBlocks,
                                                         uint32_t* __restrict__ hist /*[256][nBlocks]*/) {
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t b0 = (uint64_t)blockIdx.x * RCHUNK;
	for (int t = 0; t < RTILES; t++) {
		uint64_t i = b0 + (uint64_t)t * RB + threadIdx.x;
		if (i < N) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
	}
	__syncthreads();
	hist[(size_t)threadIdx.x * nBlocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RB) radix_scatter_kernel(const unsigned long long* __restrict__ kin, const uint32_t* __restrict__ vin, uint64_t N,
                                                            int shift, uint32_t nBlocks, const uint32_t* __restrict__ hist,
                                                            unsigned long long* __restrict__ kout, uint32_t* __restrict__ vout) {
	__shared__ uint32_t base[257];
	__shared__ uint32_t wc[8][257];
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const unsigned lt = lanemask_lt();
	{
		// exclusive scan of the 256 digit totals (radix_scan_digits_kernel) + this block's offset inside its digit
		const uint32_t tot = hist[(size_t)256 * nBlocks + tid];
		uint32_t inc = tot;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		if (lane == 31) wc[0][w] = inc;
		__syncthreads();
		uint32_t off = inc - tot;
		for (int ww = 0; ww < w; ww++) off += wc[0][ww];
		__syncthreads();
		base[tid] = off + hist[(size_t)tid * nBlocks + blockIdx.x];
	}
	if (tid == 0) base[256] = 0;
	for (int i = tid; i < 8 * 257; i += RB) (&wc[0][0])[i] = 0;
	__syncthreads();
	const uint64_t b0 = (uint64_t)blockIdx.x * RCHUNK;
	for (int t = 0; t < RTILES; t++) {
		const uint64_t i = b0 + (uint64_t)t * RB + tid;
		const bool valid = i < N;
		unsigned long long key = valid ? kin[i] : 0ull;
		uint32_t val = valid ? vin[i] : 0u;
		const uint32_t d = valid ? ((uint32_t)(key >> shift) & 255u) : 256u;
		const unsigned peers = __match_any_sync(0xffffffffu, d);
		const uint32_t rk = __popc(peers & lt), cw = __popc(peers);
		const uint32_t stamp = (uint32_t)t + 1;
		if (rk == 0) wc[w][d] = (stamp << 6) | cw;
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
}

// returns 0 with the sorted permutation in valsA (8 passes -> back in A)
int launch_sort_scores_desc(const double* scores, uint64_t N, unsigned long long* keysA, unsigned long long* keysB, uint32_t* valsA,
                            uint32_t* valsB, uint32_t* hist, cudaStream_t s) {
	if (N == 0 || N >= (1ull << 32)) return -1;
	const uint32_t nBlocks = (uint32_t)((N + RCHUNK - 1) / RCHUNK);
	sort_init_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(scores, N, keysA, valsA);
	unsigned long long *kin = keysA, *kout = keysB;
	uint32_t *vin = valsA, *vout = valsB;
	for (int pass = 0; pass < 8; pass++) {
		radix_hist_kernel<<<nBlocks, RB, 0, s>>>(kin, N, pass * 8, nBlocks, hist);
		radix_scan_digits_kernel<<<32, 256, 0, s>>>(hist, nBlocks);
		radix_scatter_kernel<<<nBlocks, RB, 0, s>>>(kin, vin, N, pass * 8, nBlocks, hist, kout, vout);
		unsigned long long* tk = kin; kin = kout; kout = tk;
		uint32_t* tv = vin; vin = vout; vout = tv;
	}
	return 0;
}

// tie_mode 2 bracket: the same descending sort, but with the ties of equal score ordered by LABEL -- positives first (the order
// that maximises both curves) or negatives first (the one that minimises them).  LSD radix sorting is stable, so it is enough to
// start from an order that has the labels partitioned: one extra radix pass on a 1-bit label key, then the keys are rebuilt in
// that order and the eight score passes follow.
__global__ void sort_init_label_kernel(const uint32_t* __restrict__ labels, uint64_t N, int posFirst, unsigned long long* __restrict__ keys,
                                       uint32_t* __restrict__ vals) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) { const bool pos = labels[i] == 1u; keys[i] = (pos == (posFirst != 0)) ? 0ull : 1ull; vals[i] = (uint32_t)i; }
}
__global__ void sort_rekey_kernel(const double* __restrict__ scores, const uint32_t* __restrict__ valsIn, uint64_t N,
                                  unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < N) { const uint32_t v = valsIn[i]; keys[i] = ~f64_key(scores[v]); vals[i] = v; }
}
// result (permutation) in valsA, sorted keys in keysA
int launch_sort_scores_desc_labelties(const double* scores, const uint32_t* labels, uint64_t N, int posFirst, unsigned long long* keysA,
                                      unsigned long long* keysB, uint32_t* valsA, uint32_t* valsB, uint32_t* hist, cudaStream_t s) {
	if (N == 0 || N >= (1ull << 32)) return -1;
	const uint32_t nBlocks = (uint32_t)((N + RCHUNK - 1) / RCHUNK);
	const unsigned g = (unsigned)((N + 255) / 256);
	sort_init_label_kernel<<<g, 256, 0, s>>>(labels, N, posFirst, keysA, valsA);
	radix_hist_kernel<<<nBlocks, RB, 0, s>>>(keysA, N, 0, nBlocks, hist);
	radix_scan_digits_kernel<<<32, 256, 0, s>>>(hist, nBlocks);
	radix_scatter_kernel<<<nBlocks, RB, 0, s>>>(keysA, valsA, N, 0, nBlocks, hist, keysB, valsB);
	sort_rekey_kernel<<<g, 256, 0, s>>>(scores, valsB, N, keysA, valsA);
	unsigned long long *kin = keysA, *kout = keysB;
	uint32_t *vin = valsA, *vout = valsB;
	for (int pass = 0; pass < 8; pass++) {
		radix_hist_kernel<<<nBlocks, RB, 0, s>>>(kin, N, pass * 8, nBlocks, hist);
		radix_scan_digits_kernel<<<32, 256, 0, s>>>(hist, nBlocks);
		radix_scatter_kernel<<<nBlocks, RB, 0, s>>>(kin, vin, N, pass * 8, nBlocks, hist, kout, vout);
		unsigned long long* tk = kin; kin = kout; kout = tk;
		uint32_t* tv = vin; vin = vout; vout = tv;
	}
	return 0;
}

// (class, fold) key of every example for the device-built fold lists (psm_folds_build_device); the per-key counts go through
// shared memory (2 * nFolds <= FK_BINS bins)
constexpr uint32_t FK_BINS = 2048;
__global__ void __launch_bounds__(256) fold_keys_kernel(const uint32_t* __restrict__ labels, const uint32_t* __restrict__ foldIds, uint32_t n, uint32_t nFolds,
                                                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ counts,
                                                         uint32_t* __restrict__ bad) {
	__shared__ uint32_t h[FK_BINS];
	const uint32_t K = 2 * nFolds;
	for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) h[k] = 0;
	__syncthreads();
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t f = foldIds[i];
		if (f >= nFolds) { atomicOr(bad, 1u); f = 0; }
		const uint32_t key = (labels[i] > 0 ? 0u : nFolds) + f;
		keys[i] = key; vals[i] = i;
		atomicAdd(&h[key], 1u);
	}
	__syncthreads();
	for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) if (h[k]) atomicAdd(&counts[k], h[k]);
}
void launch_fold_keys(const uint32_t* labels, const uint32_t* foldIds, uint32_t n, uint32_t nFolds, unsigned long long* keys, uint32_t* vals, uint32_t* counts,
                      uint32_t* bad, cudaStream_t s) {
	if (!n) return;
	const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)n + 4095) / 4096, 148 * 8);
	fold_keys_kernel<<<blocks, 256, 0, s>>>(labels, foldIds, n, nFolds, keys, vals, counts, bad);
}

// generic entry: ascending (stable) LSD sort of prefilled (keysA, valsA) on the key bits [bit0, bit0 + 8*passes); passes must be even
// so that the result is back in keysA / valsA.  hist: ((N + 4095) / 4096) * 256 + 256 uint32.
int launch_radix_sort_u64(unsigned long long* keysA, unsigned long long* keysB, uint32_t* valsA, uint32_t* valsB, uint64_t N, int passes,
                          uint32_t* hist, cudaStream_t s, int bit0) {
	if (N == 0) return 0;
	if (N >= (1ull << 32) || passes <= 0 || passes > 8 || (passes & 1) || bit0 < 0 || bit0 + 8 * passes > 64) return -1;
	const uint32_t nBlocks = (uint32_t)((N + RCHUNK - 1) / RCHUNK);
	unsigned long long *kin = keysA, *kout = keysB;
	uint32_t *vin = valsA, *vout = valsB;
	for (int pass = 0; pass < passes; pass++) {
		radix_hist_kernel<<<nBlocks, RB, 0, s>>>(kin, N, bit0 + pass * 8, nBlocks, hist);
		radix_scan_digits_kernel<<<32, 256, 0, s>>>(hist, nBlocks);
		radix_scatter_kernel<<<nBlocks, RB, 0, s>>>(kin, vin, N, bit0 + pass * 8, nBlocks, hist, kout, vout);
		unsigned long long* tk = kin; kin = kout; kout = tk;
		uint32_t* tv = vin; vin = vout; vout = tv;
	}
	return 0;
}

}  // namespace psm
