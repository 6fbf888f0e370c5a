// K1b knn_fold on the tensor cores -- the same neighbour table as knn.cu (bit for bit), for folds whose P x P x m distance
// work is large enough to matter (BASELINE configs[3]: P = 8000 positives, m = 1000 features, 1.9e11 fp64 operations per
// fold in the exact kernel).
//
// Replaces ANNkd_tree + annkSearch (reference src/samplerKNN.cpp:60-93) in three steps:
//   1. candidates: d~(i,j) = |x_i|^2 + |x_j|^2 - 2 x_i.x_j with the cross term on the 5th-generation tensor cores as a
//      3xTF32 product (x = hi + lo, both TF32; hi.hi + hi.lo + lo.hi accumulated in fp32 in TMEM): tcgen05.mma
//      kind::tf32, M = N = 128, K = 8 per instruction, issued by one thread; the operand tiles are pre-split, pre-laid-out
//      16 KB images that one thread streams into a 3-stage shared-memory ring with cp.async.bulk (TMA) on mbarriers; the
//      epilogue reads the 128 x 128 fp32 accumulator out of TMEM with tcgen05.ld and writes d~ as fp32;
//   2. per query the C smallest d~ (C = 32 or 64 > k) are kept as candidates, together with the C-th smallest value a_C;
//   3. rerank: the candidates' distances are recomputed in fp64 in ANN's own summation order (annDist: ascending feature,
//      separately rounded subtract / multiply / add -- the bits knn.cu produces), zero distances are dropped, the k smallest
//      by (distance, index) are the result.  It is PROVABLY the exact answer when  a_C - E > d_k  (every non-candidate has
//      d~ >= a_C and |d~ - d| <= E, E from the 3xTF32 / fp32 error bound below); a query that fails the test is counted and
//      the caller falls back to the exact kernel for the fold, so the table never depends on tensor-core rounding.
#include "kernels.cuh"

namespace psm {

constexpr int TC_BM = 128;                       // tile edge (rows of A = rows of B = 128 positives)
constexpr int TC_BK = 32;                        // features per pipeline stage
constexpr int TC_STAGES = 3;
constexpr uint32_t TC_IMG_FLOATS = TC_BM * TC_BK;            // one operand image: [chunk 0..7][row 0..127][4 floats]
constexpr uint32_t TC_IMG_BYTES = TC_IMG_FLOATS * 4;         // 16 KB
constexpr uint32_t TC_STAGE_BYTES = 4 * TC_IMG_BYTES;        // A hi, A lo, B hi, B lo
constexpr uint32_t TC_TMEM_COLS = 128;

__device__ __forceinline__ float to_tf32(float v) {
	uint32_t r;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
	return __uint_as_float(r);
}

// ---- operand images ---------------------------------------------------------------------------------------------------------
// images[((b * nKb + kb) * 2 + part) * TC_IMG_FLOATS + (c * 128 + r) * 4 + e] = part (0 hi / 1 lo) of feature kb*32 + c*4 + e of
// positive b*128 + r: exactly the canonical K-major, no-swizzle shared-memory layout of a UMMA operand (8-row x 16-byte core
// matrices, 128 bytes apart along the rows, 2048 bytes apart along K), so a tile is one contiguous bulk copy.
__global__ void __launch_bounds__(256) knn_tc_prep_kernel(const double* __restrict__ Xp, int P, int m, int nKb, float* __restrict__ images) {
	const int kb = blockIdx.x, b = blockIdx.y;
	float* hi = images + ((size_t)(b * nKb + kb) * 2) * TC_IMG_FLOATS;
	float* lo = hi + TC_IMG_FLOATS;
	for (int pair = threadIdx.x; pair < TC_BM * (TC_BK / 4); pair += blockDim.x) {
		const int c = pair & 7, r = pair >> 3;               // consecutive threads: consecutive chunks of one row (coalesced reads)
		const int g = b * TC_BM + r, k0 = kb * TC_BK + c * 4;
		float h[4], l[4];
#pragma unroll
		for (int e = 0; e < 4; e++) {
			const float v = (g < P && k0 + e < m) ? (float)Xp[(size_t)g * m + k0 + e] : 0.0f;
			h[e] = to_tf32(v);
			l[e] = to_tf32(v - h[e]);
		}
		*reinterpret_cast<float4*>(hi + (size_t)(c * TC_BM + r) * 4) = make_float4(h[0], h[1], h[2], h[3]);
		*reinterpret_cast<float4*>(lo + (size_t)(c * TC_BM + r) * 4) = make_float4(l[0], l[1], l[2], l[3]);
	}
}

// |x|^2 of the fp32-rounded rows (one warp per row) and the largest of them (float bits order like integers for values >= 0)
__global__ void __launch_bounds__(128) knn_tc_norms_kernel(const double* __restrict__ Xp, int P, int Ppad, int m, float* __restrict__ norms,
                                                            uint32_t* __restrict__ maxNormBits) {
	const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (row >= Ppad) return;
	float s = 0.0f;
	if (row < P)
		for (int k = lane; k < m; k += 32) { const float v = (float)Xp[(size_t)row * m + k]; s = fmaf(v, v, s); }
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if (lane == 0) { norms[row] = s; atomicMax(maxNormBits, __float_as_uint(s)); }
}

// ---- the distance GEMM --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
	uint32_t ok = 0;
	while (!ok)
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst)), "l"(src),
	             "r"(bytes), "r"(tc_smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE, leading (K) byte offset 2048, stride
// (8-row group) byte offset 128, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
	return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A and B TF32, both K-major, N = 128, M = 128, dense
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void tc_mma(uint32_t tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
	    ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

// grid (Ppad/128, Ppad/128): CTA (bj, bi) writes D[bi*128 .. +128][bj*128 .. +128]; 128 threads: lane 0 of warp 0 is the TMA
// producer, lane 0 of warp 1 issues the MMAs, all four warps run the epilogue (warp w owns TMEM lanes 32w .. 32w+31)
__global__ void __launch_bounds__(128, 1) knn_tc_gemm_kernel(const float* __restrict__ images, const float* __restrict__ norms, int nKb, int Ppad,
                                                              float* __restrict__ D) {
	extern __shared__ __align__(1024) unsigned char tc_smem[];
	__shared__ __align__(8) uint64_t fullBar[TC_STAGES], emptyBar[TC_STAGES], accumBar;
	__shared__ uint32_t tmemHolder;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int bj = blockIdx.x, bi = blockIdx.y;
	if (tid == 0) {
		for (int s = 0; s < TC_STAGES; s++) { tc_mbar_init(&fullBar[s], 1); tc_mbar_init(&emptyBar[s], 1); }
		tc_mbar_init(&accumBar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmemHolder)), "r"(TC_TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmemHolder;

	if (warp == 0 && lane == 0) {
		const float* aImg = images + (size_t)bi * nKb * 2 * TC_IMG_FLOATS;
		const float* bImg = images + (size_t)bj * nKb * 2 * TC_IMG_FLOATS;
		for (int kb = 0; kb < nKb; kb++) {
			const int st = kb % TC_STAGES, use = kb / TC_STAGES;
			if (use > 0) tc_mbar_wait(&emptyBar[st], (uint32_t)(use - 1) & 1u);   // the MMAs that read this stage have completed
			unsigned char* dst = tc_smem + (size_t)st * TC_STAGE_BYTES;
			tc_mbar_expect_tx(&fullBar[st], TC_STAGE_BYTES);
			tc_bulk_g2s(dst, aImg + (size_t)kb * 2 * TC_IMG_FLOATS, 2 * TC_IMG_BYTES, &fullBar[st]);                       // A hi | A lo (adjacent images)
			tc_bulk_g2s(dst + 2 * TC_IMG_BYTES, bImg + (size_t)kb * 2 * TC_IMG_FLOATS, 2 * TC_IMG_BYTES, &fullBar[st]);    // B hi | B lo
		}
	} else if (warp == 1 && lane == 0) {
		for (int kb = 0; kb < nKb; kb++) {
			const int st = kb % TC_STAGES, use = kb / TC_STAGES;
			tc_mbar_wait(&fullBar[st], (uint32_t)use & 1u);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			const uint32_t base = tc_smem_u32(tc_smem + (size_t)st * TC_STAGE_BYTES);
#pragma unroll
			for (int ks = 0; ks < TC_BK / 8; ks++) {                       // K = 8 per instruction = two 16-byte chunks = 4096 bytes of the image
				const uint64_t aHi = tc_desc(base + ks * 4096u), aLo = tc_desc(base + TC_IMG_BYTES + ks * 4096u);
				const uint64_t bHi = tc_desc(base + 2 * TC_IMG_BYTES + ks * 4096u), bLo = tc_desc(base + 3 * TC_IMG_BYTES + ks * 4096u);
				tc_mma(tmem, aHi, bHi, (kb | ks) ? 1u : 0u);               // 3xTF32: hi.hi + hi.lo + lo.hi (lo.lo is below fp32 resolution)
				tc_mma(tmem, aHi, bLo, 1u);
				tc_mma(tmem, aLo, bHi, 1u);
			}
			tc_commit(&emptyBar[st]);                                      // arrives when the MMAs above have read the stage
		}
		tc_commit(&accumBar);                                              // arrives when the accumulator is complete
	}
	__syncwarp();
	tc_mbar_wait(&accumBar, 0);
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	{
		const int gi = bi * TC_BM + warp * 32 + lane;
		const float ni = norms[gi];
		float* out = D + (size_t)gi * Ppad + (size_t)bj * TC_BM;
		const float* nj = norms + (size_t)bj * TC_BM;
#pragma unroll 1
		for (int c0 = 0; c0 < TC_BM; c0 += 32) {
			uint32_t v[32];
			const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
			asm volatile(
			    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
			    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
			    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
			    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
			      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
			      "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
			      "=r"(v[31])
			    : "r"(taddr));
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
			for (int q = 0; q < 32; q += 4) {
				float4 o;
				o.x = (ni + __ldg(nj + c0 + q)) - 2.0f * __uint_as_float(v[q]);
				o.y = (ni + __ldg(nj + c0 + q + 1)) - 2.0f * __uint_as_float(v[q + 1]);
				o.z = (ni + __ldg(nj + c0 + q + 2)) - 2.0f * __uint_as_float(v[q + 2]);
				o.w = (ni + __ldg(nj + c0 + q + 3)) - 2.0f * __uint_as_float(v[q + 3]);
				*reinterpret_cast<float4*>(out + c0 + q) = o;
			}
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
}

// ---- candidates: the C smallest approximate distances of every query (one warp per query) -------------------------------------
// every lane keeps the C smallest values of its stride-32 share of the row in a sorted register list (any number of the winners
// may fall into one lane), then the 32 lists are merged by C rounds of a warp arg-min over the list heads -- the scheme of
// knn_select_kernel (knn.cu) on fp32 values
template <int C>
__global__ void __launch_bounds__(128) knn_tc_select_kernel(const float* __restrict__ D, int P, int Ppad, int32_t* __restrict__ cand,
                                                             float* __restrict__ aC) {
	const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (q >= P) return;
	const float* row = D + (size_t)q * Ppad;
	const float INF = __int_as_float(0x7f800000);
	float lv[C];
	int lj[C];
#pragma unroll
	for (int i = 0; i < C; i++) { lv[i] = INF; lj[i] = -1; }
	for (int j = lane; j < P; j += 32) {
		const float v = row[j];
		if (v < lv[C - 1]) {
			lv[C - 1] = v; lj[C - 1] = j;
#pragma unroll
			for (int i = C - 1; i > 0; i--) {
				if (lv[i] < lv[i - 1]) {
					const float tv = lv[i]; lv[i] = lv[i - 1]; lv[i - 1] = tv;
					const int tj = lj[i]; lj[i] = lj[i - 1]; lj[i - 1] = tj;
				}
			}
		}
	}
	float last = INF;
	for (int r = 0; r < C; r++) {
		float bv = lv[0];
		unsigned bj = lj[0] < 0 ? 0xFFFFFFFFu : (unsigned)lj[0];
		const unsigned mine = bj;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
			const unsigned oj = __shfl_xor_sync(0xffffffffu, bj, o);
			if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
		}
		if (bj != 0xFFFFFFFFu && bj == mine) {                                  // this lane owns the winner: pop its head
#pragma unroll
			for (int i = 0; i < C - 1; i++) { lv[i] = lv[i + 1]; lj[i] = lj[i + 1]; }
			lv[C - 1] = INF; lj[C - 1] = -1;
		}
		if (lane == 0) cand[(size_t)q * C + r] = bj == 0xFFFFFFFFu ? -1 : (int32_t)bj;
		last = bj == 0xFFFFFFFFu ? INF : bv;                                    // fewer than C points: every point is a candidate
	}
	if (lane == 0) aC[q] = last;
}

// ---- rerank: exact fp64 distances of the candidates, ANN's order; completeness test ------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) knn_tc_rerank_kernel(const double* __restrict__ Xp, int P, int m, int localk, const int32_t* __restrict__ cand,
                                                             const float* __restrict__ aC, const float* __restrict__ norms,
                                                             const uint32_t* __restrict__ maxNormBits, int32_t* __restrict__ nnIdx,
                                                             double* __restrict__ nnDist, uint32_t* __restrict__ nIncomplete) {
	const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (q >= P) return;
	constexpr int PER = C / 32;
	const double INF = __longlong_as_double(0x7FF0000000000000ll);
	const double* xq = Xp + (size_t)q * m;
	double d[PER];
	int id[PER];
#pragma unroll
	for (int s = 0; s < PER; s++) {
		const int j = cand[(size_t)q * C + s * 32 + lane];
		id[s] = j; d[s] = INF;
		if (j >= 0) {
			const double* xj = Xp + (size_t)j * m;
			double acc = 0.0;
			for (int k = 0; k < m; k++) {                                      // annDist: ascending feature, no FMA (kd_util / brute.cpp)
				const double t = __dsub_rn(xq[k], xj[k]);
				acc = __dadd_rn(acc, __dmul_rn(t, t));
			}
			d[s] = acc;
			if (acc == 0.0) { d[s] = INF; id[s] = -1; }                        // self and exact duplicates are never neighbours
		}
	}
	const size_t o = (size_t)q * localk;
	double dk = 0.0;
	bool shortOf = false;
	for (int r = 0; r < localk; r++) {
		// this lane's best remaining candidate
		double bd = INF; unsigned bj = 0xFFFFFFFFu; int bs = -1;
#pragma unroll
		for (int s = 0; s < PER; s++) {
			const unsigned jj = id[s] < 0 ? 0xFFFFFFFFu : (unsigned)id[s];
			if (d[s] < bd || (d[s] == bd && jj < bj)) { bd = d[s]; bj = jj; bs = s; }
		}
		unsigned long long key = (unsigned long long)__double_as_longlong(bd);
		unsigned wj = bj;
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) {
			const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, off);
			const unsigned oj = __shfl_xor_sync(0xffffffffu, wj, off);
			if (ok < key || (ok == key && oj < wj)) { key = ok; wj = oj; }
		}
		if (wj != 0xFFFFFFFFu && wj == bj && bs >= 0) {                        // this lane owns the winner (indices are unique): retire it
#pragma unroll
			for (int s = 0; s < PER; s++) if (s == bs) { d[s] = INF; id[s] = -1; }
		}
		if (wj == 0xFFFFFFFFu) shortOf = true; else dk = __longlong_as_double((long long)key);
		if (lane == 0) {
			nnIdx[o + r] = wj == 0xFFFFFFFFu ? -1 : (int32_t)wj;
			if (nnDist) nnDist[o + r] = wj == 0xFFFFFFFFu ? INF : __longlong_as_double((long long)key);
		}
	}
	if (lane == 0) {
		// every non-candidate j has d~_j >= a_C; |d~_j - d_j| <= E with E = (|x_q|^2 + max|x|^2) * (2m + 32) * 2^-23 (input rounding to
		// fp32, fp32 norms, dropped lo.lo terms, fp32 accumulation of m products): complete iff a_C - E > d_k, or all points were candidates
		const float a = aC[q];
		const double E = ((double)norms[q] + (double)__uint_as_float(*maxNormBits)) * (double)(2 * m + 32) * (1.0 / 8388608.0);
		const bool all = !(a < __int_as_float(0x7f800000));
		if (!all && (shortOf || !((double)a - E > dk))) atomicAdd(nIncomplete, 1u);
	}
}

size_t knn_tensor_workspace_bytes(int P, int m, int C) {
	const size_t Ppad = ((size_t)P + TC_BM - 1) / TC_BM * TC_BM, nKb = ((size_t)m + TC_BK - 1) / TC_BK;
	return Ppad / TC_BM * nKb * 2 * TC_IMG_BYTES + Ppad * Ppad * 4 + Ppad * 4 + (size_t)P * C * 4 + (size_t)P * 4 + 256 + 6 * 256;
}

// Candidate count for `localk` neighbours, 0 if the tensor path does not apply.
int knn_tensor_candidates(int localk) { return localk <= 12 ? 32 : localk <= 28 ? 64 : 0; }

// *nIncompleteDev (device u32, zeroed here) counts the queries whose result is not provably exact; the caller reads it back and
// runs launch_knn for the fold if it is not zero.
int launch_knn_tensor(const double* Xp, int P, int m, int localk, void* workspace, int32_t* nnIdx, double* nnDist, uint32_t* nIncompleteDev,
                      cudaStream_t s) {
	const int C = knn_tensor_candidates(localk);
	if (C == 0 || P < 2) return -1;
	const int Ppad = (P + TC_BM - 1) / TC_BM * TC_BM, nKb = (m + TC_BK - 1) / TC_BK, nBlk = Ppad / TC_BM;
	unsigned char* w = (unsigned char*)workspace;
	auto take = [&](size_t bytes) { unsigned char* p = w; w += (bytes + 255) & ~(size_t)255; return p; };
	float* images = (float*)take((size_t)nBlk * nKb * 2 * TC_IMG_BYTES);
	float* D = (float*)take((size_t)Ppad * Ppad * 4);
	float* norms = (float*)take((size_t)Ppad * 4);
	int32_t* cand = (int32_t*)take((size_t)P * C * 4);
	float* aC = (float*)take((size_t)P * 4);
	uint32_t* maxNorm = (uint32_t*)take(4);
	if (cudaMemsetAsync(maxNorm, 0, 4, s) != cudaSuccess || cudaMemsetAsync(nIncompleteDev, 0, 4, s) != cudaSuccess) return -1;
	knn_tc_prep_kernel<<<dim3(nKb, nBlk), 256, 0, s>>>(Xp, P, m, nKb, images);
	knn_tc_norms_kernel<<<(Ppad * 32 + 127) / 128, 128, 0, s>>>(Xp, P, Ppad, m, norms, maxNorm);
	const size_t smem = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024;
	if (allow_max_dyn_smem(knn_tc_gemm_kernel) != cudaSuccess) return -1;
	knn_tc_gemm_kernel<<<dim3(nBlk, nBlk), 128, smem, s>>>(images, norms, nKb, Ppad, D);
	const int blocks = (P * 32 + 127) / 128;
	if (C == 32) {
		knn_tc_select_kernel<32><<<blocks, 128, 0, s>>>(D, P, Ppad, cand, aC);
		knn_tc_rerank_kernel<32><<<blocks, 128, 0, s>>>(Xp, P, m, localk, cand, aC, norms, maxNorm, nnIdx, nnDist, nIncompleteDev);
	} else {
		knn_tc_select_kernel<64><<<blocks, 128, 0, s>>>(D, P, Ppad, cand, aC);
		knn_tc_rerank_kernel<64><<<blocks, 128, 0, s>>>(Xp, P, m, localk, cand, aC, norms, maxNorm, nnIdx, nnDist, nIncompleteDev);
	}
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace psm
