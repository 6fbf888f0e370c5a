// K6 predict_accumulate -- drop the held-out fold through every forest of a partition batch and add the
// partition-averaged class fractions into the fold score buffer.
//
// Replaces, per partition of the reference:
//   Sampler::copyTestSet + transposeMatrix + DataDouble     src/sampler.cpp:81-98, src/hyperSMURF.cpp:316-320
//   rfRanger predict ctor (test-fold Data::sort + deep copy of the forest -- pure waste)  src/rfRanger.cpp:56-118
//   Tree::predict                                            src/ranger/Tree/Tree.cpp:137-193
//   ForestProbability::predictInternal                       src/ranger/Forest/ForestProbability.cpp:124-150
//   Sampler::accumulateAndDivideResInProbVect                src/sampler.cpp:116-133
//
// Each CTA keeps a tile of 256 test rows resident in shared memory (feature-major, so a lane reads
// tile[feature][lane]: conflict-free whatever feature each lane is at) and streams 16-byte packed nodes
// through the read-only path; rows are gathered from the resident dataset by index, never materialised.
// Per (partition, row) forest means go to a scratch plane and a second kernel adds them in ascending
// partition order -- the reference's fp64 summation order at ensThrd = 1, so scores are bit-identical.
#include "kernels.cuh"

#include <cstdlib>

namespace psm {

constexpr int PT = 256;   // rows per CTA

template <bool SMEM>
__global__ void __launch_bounds__(PT) predict_kernel(const double* __restrict__ x, uint32_t m, const uint32_t* __restrict__ testIdx,
                                                      uint32_t Nte, const Node16* __restrict__ nodes, uint32_t NC, uint32_t nPartsBatch,
                                                      uint32_t T, uint32_t pch, double* __restrict__ scratch) {
	extern __shared__ double tile[];   // [m][PT]
	const uint32_t ld = m + 1;
	const uint32_t r0 = blockIdx.x * PT;
	const int tid = threadIdx.x;
	if (SMEM) {
		const int lane = tid & 31, w = tid >> 5;
		for (int rr = w; rr < PT; rr += PT / 32) {
			uint32_t r = r0 + rr;
			if (r < Nte) {
				const double* src = x + (size_t)testIdx[r] * ld;
				for (uint32_t c = lane; c < m; c += 32) tile[(size_t)c * PT + rr] = src[c];
			}
		}
		__syncthreads();
	}
	const uint32_t row = r0 + tid;
	if (row >= Nte) return;
	const double* xr = x + (size_t)testIdx[row] * ld;
	const uint32_t p0 = blockIdx.y * pch;
	const uint32_t p1 = min(p0 + pch, nPartsBatch);
	const double Td = (double)T;
	for (uint32_t pl = p0; pl < p1; pl++) {
		double s0 = 0.0, s1 = 0.0;
		for (uint32_t t = 0; t < T; t++) {
			const Node16* base = nodes + ((size_t)pl * T + t) * NC;
			uint32_t idx = 0;
			double a;
			unsigned long long b;
			while (true) {
				const uint4 raw = __ldg(reinterpret_cast<const uint4*>(base + idx));
				a = __hiloint2double((int)raw.y, (int)raw.x);
				b = ((unsigned long long)raw.w << 32) | raw.z;
				if ((int)raw.w < 0) break;                                   // leaf: sign bit of the second word
				const double v = SMEM ? tile[(size_t)raw.z * PT + tid] : xr[raw.z];
				idx = raw.w + (v <= a ? 0u : 1u);                            // Tree.cpp:163-169; right = left + 1
			}
			s0 = __dadd_rn(s0, a);                                           // ForestProbability.cpp:136-141
			s1 = __dadd_rn(s1, __longlong_as_double((long long)(b & 0x7FFFFFFFFFFFFFFFull)));
		}
		scratch[((size_t)pl * 2 + 0) * Nte + row] = __ddiv_rn(s0, Td);       // :145-149
		scratch[((size_t)pl * 2 + 1) * Nte + row] = __ddiv_rn(s1, Td);
	}
}

// ---- sm_100a path: forests staged through shared memory by the TMA (cp.async.bulk + mbarrier), double-buffered ----
// The L1 path above is bound by divergent 16-byte node loads (ncu r01: 7 sectors/request, long-scoreboard stalls);
// here a whole tree (nNodes*16 B, contiguous) lands in shared memory with one bulk copy issued by one thread while the
// CTA is still traversing the previous tree, and node fetches become LDS.128.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "WAIT_LOOP:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra DONE;\n\t"
	    "bra WAIT_LOOP;\n\t"
	    "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
	             "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

__global__ void __launch_bounds__(PT) predict_tma_kernel(const double* __restrict__ x, uint32_t m, const uint32_t* __restrict__ testIdx,
                                                          uint32_t Nte, const Node16* __restrict__ nodes, uint32_t NC,
                                                          const TreeState* __restrict__ ts, uint32_t nPartsBatch, uint32_t T, uint32_t pch,
                                                          uint32_t maxNodes, double* __restrict__ scratch) {
	extern __shared__ __align__(128) unsigned char smraw[];
	double* tile = reinterpret_cast<double*>(smraw);                                   // [m][PT]
	Node16* tbuf = reinterpret_cast<Node16*>(smraw + (size_t)m * PT * sizeof(double)); // [2][maxNodes]
	uint64_t* bars = reinterpret_cast<uint64_t*>(tbuf + 2 * (size_t)maxNodes);         // [2]
	const uint32_t ld = m + 1;
	const uint32_t r0 = blockIdx.x * PT;
	const int tid = threadIdx.x;
	const uint32_t p0 = blockIdx.y * pch;
	const uint32_t p1 = min(p0 + pch, nPartsBatch);
	const uint32_t g0 = p0 * T, nT = (p1 - p0) * T;
	if (tid == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0 && nT > 0) {
		const uint32_t bytes = ts[g0].nNodes * (uint32_t)sizeof(Node16);
		mbar_expect_tx(&bars[0], bytes);
		tma_bulk_g2s(tbuf, nodes + (size_t)g0 * NC, bytes, &bars[0]);
	}
	{
		const int lane = tid & 31, w = tid >> 5;
		for (int rr = w; rr < PT; rr += PT / 32) {
			uint32_t r = r0 + rr;
			if (r < Nte) {
				const double* src = x + (size_t)testIdx[r] * ld;
				for (uint32_t c = lane; c < m; c += 32) tile[(size_t)c * PT + rr] = src[c];
			}
		}
	}
	__syncthreads();
	const uint32_t row = r0 + tid;
	const bool valid = row < Nte;
	const double Td = (double)T;
	double s0 = 0.0, s1 = 0.0;
	for (uint32_t i = 0; i < nT; i++) {
		const uint32_t b = i & 1;
		if (tid == 0 && i + 1 < nT) {                     // buffer b^1 was released by the __syncthreads that ended iteration i-1
			const uint32_t bytes = ts[g0 + i + 1].nNodes * (uint32_t)sizeof(Node16);
			mbar_expect_tx(&bars[b ^ 1], bytes);
			tma_bulk_g2s(tbuf + (size_t)(b ^ 1) * maxNodes, nodes + (size_t)(g0 + i + 1) * NC, bytes, &bars[b ^ 1]);
		}
		mbar_wait(&bars[b], (i >> 1) & 1);
		if (valid) {
			const uint4* base = reinterpret_cast<const uint4*>(tbuf + (size_t)b * maxNodes);
			uint32_t idx = 0;
			uint4 raw;
			while (true) {
				raw = base[idx];
				if ((int)raw.w < 0) break;
				const double a = __hiloint2double((int)raw.y, (int)raw.x);
				const double v = tile[(size_t)raw.z * PT + tid];
				idx = raw.w + (v <= a ? 0u : 1u);
			}
			s0 = __dadd_rn(s0, __hiloint2double((int)raw.y, (int)raw.x));
			s1 = __dadd_rn(s1, __hiloint2double((int)(raw.w & 0x7FFFFFFFu), (int)raw.z));
			if ((i + 1) % T == 0) {
				const uint32_t pl = p0 + i / T;
				scratch[((size_t)pl * 2 + 0) * Nte + row] = __ddiv_rn(s0, Td);
				scratch[((size_t)pl * 2 + 1) * Nte + row] = __ddiv_rn(s1, Td);
				s0 = 0.0; s1 = 0.0;
			}
		}
		__syncthreads();
	}
}

// ---- compact path: 8-byte nodes {float(threshold), feature:12 | left child:20} and an fp32 row tile ----------------------
// double -> float rounding is monotone, so  float(v) < float(thr)  implies  v < thr  and  float(v) > float(thr)  implies
// v > thr; only float(v) == float(thr) is undecided and is resolved with the exact fp64 values from global memory (about one
// visit in 10^7).  Decisions are therefore identical to ranger's fp64 `value <= split_value` (Tree.cpp:163), while shared
// memory per CTA drops from 88 KB to 44 KB (5 CTAs/SM instead of 2) and a node visit costs LDS.64 + LDS.32 instead of
// LDS.128 + LDS.64 (ncu r01: the 16-byte-node kernel is bound by shared-memory wavefronts and LDS latency).
// Half of the lane-cycles of the walk are lost to depth divergence (a warp is busy for the deepest of its 32 paths).  Visiting
// the test rows in an order that groups similar rows was tried and does not help: keyed by the total path length in the
// first 8 trees of the fold, or by the leaf reached in the first tree (probe kernel + counting sort, scores unchanged), the
// kernel ran 5-8% SLOWER at BASELINE configs[1] and configs[2] -- a row's depth in one tree says little about its depth in
// another (the trees split on random candidates among 21 noise features), so the warp maximum stays where it is.
// Leaf payloads are the exact fp64 fractions count_c / n (TreeProbability.cpp:47-63).  Almost every leaf is pure or has
// n <= 10 samples (min_node_size), so its pair of fractions is one of the 66 values p/n, n <= 10: the leaf stores the code
// n*11+p and the kernel looks the two doubles up in a shared-memory table built with the same IEEE division; the rare other
// leaf (a larger node without a valid split) stores code 0xFFFFFFFF and reads its fractions from the 16-byte node array.
// A leaf is a fixed point of the walk step: its "threshold" is a NaN (every ordered compare is false, so the step neither
// goes left nor records an equality) whose payload carries the fraction code, and its child field is its own index -- a walk
// that has arrived can be stepped again without moving, which is what lets two walks share one loop (walk_two).
struct __align__(8) Node8 {
	float thr;       // internal: float(threshold); leaf: LEAF_NAN | fraction code
	uint32_t meta;   // internal: RIGHT child = left + 1 (bits 0..19) | feature << 20 (top bit clear); leaf: 0x80000000 | own index
};
constexpr int LEAF_TAB = 121;
constexpr uint32_t LEAF_NAN = 0x7FC00000u, LEAF_UNCODED = 0x3FFFFFu;   // quiet NaN; payload = code n*11+p, or "read the 16-byte node"
// The table of leaf fractions lives in global memory (one 16-byte L1-resident load per row and tree).  As a __shared__ array its
// 1.9 KB of static shared memory were what kept BASELINE configs[1] at four CTAs per SM instead of five (ncu r02:
// launch__occupancy_limit_shared_mem 4 with 44.9 KB of dynamic shared memory per CTA; the walk is latency-bound -- short
// scoreboard 3.9 and fixed-latency waits 3.2 cycles per issue at 8 warps per scheduler).
#ifndef PSM_PREDICT_LTAB_GLOBAL
#define PSM_PREDICT_LTAB_GLOBAL 1
#endif
__device__ double2 g_leaf_tab[LEAF_TAB];
__global__ void leaf_tab_init_kernel() {
	const int tid = threadIdx.x;
	if (tid >= LEAF_TAB) return;
	const int n = tid / 11, p2 = tid % 11;
	const bool ok = n >= 1 && p2 <= n;
	g_leaf_tab[tid] = make_double2(ok ? __ddiv_rn((double)p2, (double)n) : 0.0, ok ? __ddiv_rn((double)(n - p2), (double)n) : 0.0);
}

// the one-in-10^7 visit where float(value) == float(threshold): decide with the exact fp64 values.  Kept out of line on
// purpose -- inlined, the compiler if-converts it and its ~12 predicated instructions are issued on EVERY node visit
// (ncu r01: 35 instructions per visit, issue-bound).
__device__ __noinline__ uint32_t predict_exact_leaf(const double* __restrict__ xr, const Node16* __restrict__ gtree) {
	uint32_t idx = 0;
	while (true) {
		const Node16 nd = gtree[idx];
		if (node_is_leaf(nd)) return idx;
		idx = (uint32_t)(nd.b >> 32) + (xr[(uint32_t)nd.b] <= nd.a ? 0u : 1u);       // Tree.cpp:163-169
	}
}

__global__ void __launch_bounds__(256) pack_forest_kernel(const Node16* __restrict__ nodes, const TreeState* __restrict__ ts, uint32_t NC,
                                                           uint32_t nTrees, Node8* __restrict__ out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (size_t)nTrees * NC) return;
	const uint32_t tree = (uint32_t)(i / NC), node = (uint32_t)(i % NC);
	Node8 o; o.thr = __uint_as_float(LEAF_NAN | LEAF_UNCODED); o.meta = 0x80000000u | node;
	if (node < ts[tree].nNodes) {
		const Node16 nd = nodes[i];
		if (!node_is_leaf(nd)) { o.thr = __double2float_rn(nd.a); o.meta = ((uint32_t)(nd.b >> 32) + 1u) | ((uint32_t)(nd.b & 0x7FFu) << 20); }
		else {
			const double f0 = nd.a, f1 = __longlong_as_double((long long)(nd.b & 0x7FFFFFFFFFFFFFFFull));
			uint32_t code = LEAF_UNCODED;
			for (int n = 1; n <= 10 && code == LEAF_UNCODED; n++)
				for (int p2 = 0; p2 <= n; p2++)
					if (__ddiv_rn((double)p2, (double)n) == f0 && __ddiv_rn((double)(n - p2), (double)n) == f1) { code = (uint32_t)(n * 11 + p2); break; }
			o.thr = __uint_as_float(LEAF_NAN | code);
		}
	}
	out[i] = o;
}

#ifndef PSM_CPT
#define PSM_CPT 256
#endif
constexpr int CPT = PSM_CPT;          // rows per CTA of the compact kernel
#ifndef PSM_NBUF
#define PSM_NBUF 2
#endif
constexpr uint32_t NBUF = PSM_NBUF;   // pipeline depth: the TMA runs NBUF-1 stages ahead of the traversal (deeper pipelines cost occupancy and measured slower)

// Hand-written traversal, 11 instructions per node visit (nvcc's version of the same loop: 17; the kernel is issue- and
// latency-bound, ncu r01 / r02).  Node8: .x = float(threshold), .y = right child (bits 0..19) | feature << 20; leaves: .x = NaN
// with the fraction code, .y = sign bit | own index.  Go left iff value < threshold (Tree.cpp:163: left iff value <= split_value;
// the float compare only differs from the double one when float(value) == float(threshold), tracked in `und`).
// (all working registers are PTX-local and the outputs are written once at the end: an output operand may share its register
// with an input, so it must not be written while an input is still needed)
__device__ __forceinline__ void walk_one(uint32_t tileLane, uint32_t nbase, uint32_t& lx, uint32_t& na, uint32_t& und) {
	asm volatile(
	    "{\n\t"
	    ".reg .pred pl, pg, pu;\n\t"
	    ".reg .b32 nx, ny, a, f, fa, c;\n\t"
	    ".reg .f32 v, t;\n\t"
	    "setp.ne.u32 pu, 0, 0;\n\t"
	    "mov.u32 a, %4;\n\t"
	    "ld.shared.v2.u32 {nx, ny}, [a];\n\t"
	    "setp.lt.s32 pl, ny, 0;\n\t"
	    "@pl bra PSM_WALK_DONE;\n"
	    "PSM_WALK:\n\t"
	    "shr.u32 f, ny, 20;\n\t"
	    "mad.lo.u32 fa, f, %5, %3;\n\t"
	    "ld.shared.f32 v, [fa];\n\t"
	    "and.b32 c, ny, 0xFFFFF;\n\t"
	    "mad.lo.u32 a, c, 8, %4;\n\t"
	    "mov.b32 t, nx;\n\t"
	    "setp.lt.f32 pg, v, t;\n\t"
	    "setp.eq.or.f32 pu, v, t, pu;\n\t"
	    "@pg sub.u32 a, a, 8;\n\t"
	    "ld.shared.v2.u32 {nx, ny}, [a];\n\t"
	    "setp.ge.s32 pl, ny, 0;\n\t"
	    "@pl bra PSM_WALK;\n"
	    "PSM_WALK_DONE:\n\t"
	    "mov.u32 %0, nx;\n\t"
	    "mov.u32 %1, a;\n\t"
	    "selp.u32 %2, 1, 0, pu;\n\t"
	    "}"
	    : "=r"(lx), "=r"(na), "=r"(und)
	    : "r"(tileLane), "r"(nbase), "n"(CPT * 4)
	    : "memory");
}
// Two trees at once for the same row: the walk is a chain of dependent shared-memory loads (ncu r02: short-scoreboard and
// fixed-latency stalls are 7 of the 12 stall cycles per issued instruction at 10 warps per scheduler), so a second,
// independent chain in the same warp doubles the loads in flight at the same instruction count -- the warp already runs
// to the deepest of its 32 walks, now to the deepest of 64.  A walk that has reached its leaf keeps being stepped: the leaf
// is a fixed point (Node8), only the two loads are switched off by the walk's predicate (plA / plB), so a finished walk costs
// no shared-memory wavefronts and nothing has to be selected or restored.
__device__ __forceinline__ void walk_two(uint32_t tileLane, uint32_t nbaseA, uint32_t nbaseB, uint32_t& lxA, uint32_t& naA, uint32_t& undA,
                                         uint32_t& lxB, uint32_t& naB, uint32_t& undB) {
	asm volatile(
	    "{\n\t"
	    ".reg .pred plA, pgA, puA, plB, pgB, puB, pany;\n\t"
	    ".reg .b32 nxA, nyA, aA, fA, faA, cA, nxB, nyB, aB, fB, faB, cB;\n\t"
	    ".reg .f32 vA, tA, vB, tB;\n\t"
	    "setp.ne.u32 puA, 0, 0;\n\t"
	    "setp.ne.u32 puB, 0, 0;\n\t"
	    "mov.f32 vA, 0f00000000;\n\t"
	    "mov.f32 vB, 0f00000000;\n\t"
	    "mov.u32 aA, %7;\n\t"
	    "mov.u32 aB, %8;\n\t"
	    "ld.shared.v2.u32 {nxA, nyA}, [aA];\n\t"
	    "ld.shared.v2.u32 {nxB, nyB}, [aB];\n\t"
	    "setp.ge.s32 plA, nyA, 0;\n\t"
	    "setp.ge.s32 plB, nyB, 0;\n\t"
	    "or.pred pany, plA, plB;\n\t"
	    "@!pany bra PSM_WALK2_DONE;\n"
	    "PSM_WALK2:\n\t"
	    "shr.u32 fA, nyA, 20;\n\t"
	    "shr.u32 fB, nyB, 20;\n\t"
	    "mad.lo.u32 faA, fA, %9, %6;\n\t"
	    "mad.lo.u32 faB, fB, %9, %6;\n\t"
	    "@plA ld.shared.f32 vA, [faA];\n\t"
	    "@plB ld.shared.f32 vB, [faB];\n\t"
	    "and.b32 cA, nyA, 0xFFFFF;\n\t"
	    "and.b32 cB, nyB, 0xFFFFF;\n\t"
	    "mad.lo.u32 aA, cA, 8, %7;\n\t"
	    "mad.lo.u32 aB, cB, 8, %8;\n\t"
	    "mov.b32 tA, nxA;\n\t"
	    "mov.b32 tB, nxB;\n\t"
	    "setp.lt.f32 pgA, vA, tA;\n\t"
	    "setp.lt.f32 pgB, vB, tB;\n\t"
	    "setp.eq.or.f32 puA, vA, tA, puA;\n\t"
	    "setp.eq.or.f32 puB, vB, tB, puB;\n\t"
	    "@pgA sub.u32 aA, aA, 8;\n\t"
	    "@pgB sub.u32 aB, aB, 8;\n\t"
	    "@plA ld.shared.v2.u32 {nxA, nyA}, [aA];\n\t"
	    "@plB ld.shared.v2.u32 {nxB, nyB}, [aB];\n\t"
	    "setp.ge.s32 plA, nyA, 0;\n\t"
	    "setp.ge.s32 plB, nyB, 0;\n\t"
	    "or.pred pany, plA, plB;\n\t"
	    "@pany bra PSM_WALK2;\n"
	    "PSM_WALK2_DONE:\n\t"
	    "mov.u32 %0, nxA;\n\t"
	    "mov.u32 %1, aA;\n\t"
	    "selp.u32 %2, 1, 0, puA;\n\t"
	    "mov.u32 %3, nxB;\n\t"
	    "mov.u32 %4, aB;\n\t"
	    "selp.u32 %5, 1, 0, puB;\n\t"
	    "}"
	    : "=r"(lxA), "=r"(naA), "=r"(undA), "=r"(lxB), "=r"(naB), "=r"(undB)
	    : "r"(tileLane), "r"(nbaseA), "r"(nbaseB), "n"(CPT * 4)
	    : "memory");
}

// TPS = trees per pipeline stage (PSM_PREDICT_TPS=1|2|4|8 overrides the default of 1, see launch_predict).
// DIRECT (one partition chunk: the CTA walks every forest of the batch for its rows): the partition means are added into the
// fold score buffer right here, in ascending partition order with the arithmetic of accumulate_kernel (sampler.cpp:121-132) --
// no [partition][row] scratch plane is written and read back (2 x 22 GB per BASELINE configs[2] fold).
template <uint32_t TPS, bool DIRECT>
__global__ void __launch_bounds__(CPT, TPS == 2 ? 5 : 6) predict_compact_kernel(const double* __restrict__ x, uint32_t m, const uint32_t* __restrict__ testIdx,
                                                              uint32_t Nte, const Node16* __restrict__ nodes, const Node8* __restrict__ nodes8,
                                                              uint32_t NC, const TreeState* __restrict__ ts, uint32_t nPartsBatch, uint32_t T,
                                                              uint32_t pch, uint32_t maxNodes, double* __restrict__ scratch, double divider,
                                                              double* __restrict__ class12) {
	extern __shared__ __align__(128) unsigned char smraw[];
	float* tile = reinterpret_cast<float*>(smraw);                                     // [m][CPT]
	Node8* tbuf = reinterpret_cast<Node8*>(smraw + (size_t)m * CPT * sizeof(float));    // [NBUF][TPS][maxNodes]
	uint64_t* bars = reinterpret_cast<uint64_t*>(tbuf + (size_t)NBUF * TPS * maxNodes);   // [NBUF]
#if !PSM_PREDICT_LTAB_GLOBAL
	__shared__ double ltab[2][LEAF_TAB];
#endif
	const uint32_t ld = m + 1;
	const uint32_t r0 = blockIdx.x * CPT;
	const int tid = threadIdx.x;
	const uint32_t p0 = blockIdx.y * pch;
	const uint32_t p1 = min(p0 + pch, nPartsBatch);
	const uint32_t g0 = p0 * T, nT = (p1 - p0) * T;
	const uint32_t nStages = (nT + TPS - 1) / TPS;
#if !PSM_PREDICT_LTAB_GLOBAL
	if (tid < LEAF_TAB) {
		const int n = tid / 11, p2 = tid % 11;
		const bool ok = n >= 1 && p2 <= n;
		ltab[0][tid] = ok ? __ddiv_rn((double)p2, (double)n) : 0.0;
		ltab[1][tid] = ok ? __ddiv_rn((double)(n - p2), (double)n) : 0.0;
	}
#endif
	if (tid == 0) {
		for (uint32_t q = 0; q < NBUF; q++) mbar_init(&bars[q], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	// thread 0 is the TMA producer; the tree sizes of the stage after next are fetched one stage early so that the global
	// load latency never sits on the critical path of the block
	uint32_t szNext[TPS];
	auto load_sizes = [&](uint32_t stage, uint32_t (&sz)[TPS]) {
#pragma unroll
		for (uint32_t q = 0; q < TPS; q++) {
			const uint32_t t = stage * TPS + q;
			sz[q] = (t < nT) ? ((ts[g0 + t].nNodes + 1u) & ~1u) * (uint32_t)sizeof(Node8) : 0u;
		}
	};
	auto issue_stage = [&](uint32_t stage, const uint32_t (&sz)[TPS]) {
		const uint32_t bb = stage % NBUF;
		uint32_t total = 0;
#pragma unroll
		for (uint32_t q = 0; q < TPS; q++) total += sz[q];
		mbar_expect_tx(&bars[bb], total);
#pragma unroll
		for (uint32_t q = 0; q < TPS; q++)
			if (sz[q]) tma_bulk_g2s(tbuf + ((size_t)bb * TPS + q) * maxNodes, nodes8 + (size_t)(g0 + stage * TPS + q) * NC, sz[q], &bars[bb]);
	};
	if (tid == 0) {
		for (uint32_t q = 0; q + 1 < NBUF && q < nStages; q++) {
			uint32_t sz0[TPS];
			load_sizes(q, sz0);
			issue_stage(q, sz0);
		}
		load_sizes(NBUF - 1, szNext);
	}
	{
		const int lane = tid & 31, w = tid >> 5;
		for (int rr = w; rr < CPT; rr += CPT / 32) {
			uint32_t r = r0 + rr;
			if (r < Nte) {
				const double* src = x + (size_t)testIdx[r] * ld;
				for (uint32_t c = lane; c < m; c += 32) tile[(size_t)c * CPT + rr] = __double2float_rn(src[c]);
			}
		}
	}
	__syncthreads();
	const uint32_t row = r0 + tid;
	const bool valid = row < Nte;
	const double Td = (double)T;
	const uint32_t tileLane = smem_u32(tile) + (uint32_t)tid * 4u;   // shared address of tile[0][tid]; CPT*4 bytes per feature
	const uint32_t tbuf32 = smem_u32(tbuf);
	const uint32_t treeBytes = maxNodes * (uint32_t)sizeof(Node8);
	double s0 = 0.0, s1 = 0.0;
	uint32_t tInPart = 0;                                 // trees of the current partition already summed
	double* outp = scratch + (size_t)p0 * 2 * Nte + row;  // scratch[(pl*2 + c)*Nte + row] of the current partition
	double c1 = 0.0, c2 = 0.0;                            // DIRECT: this row's fold scores
	if (DIRECT && valid) { c1 = class12[row]; c2 = class12[(size_t)Nte + row]; }
	for (uint32_t st = 0; st < nStages; st++) {
		const uint32_t b = st % NBUF;
		if (tid == 0 && st + NBUF - 1 < nStages) {        // its buffer was released by the barrier that ended stage st-1
			issue_stage(st + NBUF - 1, szNext);
			load_sizes(st + NBUF, szNext);
		}
		mbar_wait(&bars[b], (st / NBUF) & 1);
		if (valid) {
			const uint32_t nq = min(TPS, nT - st * TPS);
			uint32_t nbase = tbuf32 + b * TPS * treeBytes;
			// leaf payload of one finished walk -> the partition sums, in tree order (ForestProbability.cpp:136-149)
			auto finish = [&](uint32_t lx, uint32_t na, uint32_t und, uint32_t nb, uint32_t q) {
				double f0, f1;
				lx &= LEAF_UNCODED;                                          // the payload of the leaf's NaN
				if (und | (lx >= (uint32_t)LEAF_TAB)) {                       // rare: exact redo and/or an uncoded leaf
					const Node16* gtree = nodes + (size_t)(g0 + st * TPS + q) * NC;
					uint32_t idx = (na - nb) >> 3;
					if (und) idx = predict_exact_leaf(x + (size_t)testIdx[row] * ld, gtree);   // about one row-tree in 10^6
					const uint4 raw = __ldg(reinterpret_cast<const uint4*>(gtree + idx));
					f0 = __hiloint2double((int)raw.y, (int)raw.x);
					f1 = __hiloint2double((int)(raw.w & 0x7FFFFFFFu), (int)raw.z);
				} else {
#if PSM_PREDICT_LTAB_GLOBAL
					const double2 lf = __ldg(&g_leaf_tab[lx]);
					f0 = lf.x; f1 = lf.y;
#else
					f0 = ltab[0][lx]; f1 = ltab[1][lx];
#endif
				}
				s0 = __dadd_rn(s0, f0);                                      // ForestProbability.cpp:136-141
				s1 = __dadd_rn(s1, f1);
				if (++tInPart == T) {
					if (DIRECT) {
						c1 = __dadd_rn(c1, __dmul_rn(__ddiv_rn(s0, Td), divider));
						c2 = __dadd_rn(c2, __dmul_rn(__ddiv_rn(s1, Td), divider));
					} else {
						outp[0] = __ddiv_rn(s0, Td);                             // :145-149
						outp[Nte] = __ddiv_rn(s1, Td);
						outp += (size_t)2 * Nte;
					}
					s0 = 0.0; s1 = 0.0; tInPart = 0;
				}
			};
			if (TPS == 2 && nq == 2) {
				// two trees per stage: the two walks are interleaved in one loop (walk_two), tree order kept in the sums
				uint32_t lxA, naA, undA, lxB, naB, undB;
				walk_two(tileLane, nbase, nbase + treeBytes, lxA, naA, undA, lxB, naB, undB);
				finish(lxA, naA, undA, nbase, 0);
				finish(lxB, naB, undB, nbase + treeBytes, 1);
			} else {
#pragma unroll 1
				for (uint32_t q = 0; q < nq; q++, nbase += treeBytes) {
					uint32_t lx, na, und;
					walk_one(tileLane, nbase, lx, na, und);
					finish(lx, na, und, nbase, q);
				}
			}
		}
		// (r02, measured and not kept: no CTA barrier per tree -- every warp arrives on a "buffer free" mbarrier when it leaves a tree
		// and the producer waits on it before reusing the buffer, so the warps drift up to a tree apart: predict 13.0 ms per
		// BASELINE configs[1] CV against 12.4 with the barrier, 14.3 with three buffers)
		__syncthreads();
	}
	if (DIRECT && valid) { class12[row] = c1; class12[(size_t)Nte + row] = c2; }
}

__global__ void __launch_bounds__(256) accumulate_kernel(const double* __restrict__ scratch, uint32_t Nte, uint32_t nPartsBatch,
                                                          double divider, double* __restrict__ class12) {
	const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= Nte) return;
	double c1 = class12[row], c2 = class12[(size_t)Nte + row];
	for (uint32_t pl = 0; pl < nPartsBatch; pl++) {                          // sampler.cpp:121-132, ascending partition order
		c1 = __dadd_rn(c1, __dmul_rn(scratch[((size_t)pl * 2 + 0) * Nte + row], divider));
		c2 = __dadd_rn(c2, __dmul_rn(scratch[((size_t)pl * 2 + 1) * Nte + row], divider));
	}
	class12[row] = c1;
	class12[(size_t)Nte + row] = c2;
}

void launch_pack_forest(const Node16* nodes, const TreeState* ts, uint32_t NC, uint32_t nTrees, void* nodes8, cudaStream_t s) {
	const size_t total = (size_t)nTrees * NC;
	pack_forest_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(nodes, ts, NC, nTrees, reinterpret_cast<Node8*>(nodes8));
}

// returns 1 when the fold scores were updated directly (no scratch plane: the caller skips launch_accumulate), 0 when the
// per-partition means are in `scratch`, -1 on a configuration error
int launch_predict(const double* x, uint32_t m, const uint32_t* testIdx, uint32_t Nte, const Node16* nodes, const void* nodes8, uint32_t NC,
                   const TreeState* ts, uint32_t maxNodes, uint32_t nPartsBatch, uint32_t T, double* scratch, double divider, double* class12,
                   cudaStream_t s) {
	if (Nte == 0 || nPartsBatch == 0) return 0;
	{
		// compact path: fp32 tile + 8-byte nodes (feature < 2047, child id < 2^20)
		const uint32_t mn = (maxNodes + 1u) & ~1u;
		// trees per stage: two, walked side by side (walk_two), when four tree buffers still leave the five CTAs per SM its 48
		// registers allow -- small trees, e.g. BASELINE configs[2]; otherwise one.  Measured on B200 (predict ms per CV, r02): configs[2]
		// at a tenth of the rows 14.9 (1) / 14.1 (2); configs[1], whose ~9 KB trees drop the kernel from five CTAs per SM to three
		// with four buffers, 12.3 (1) / 13.5 (2).  (r01, two trees walked one after the other: 13.3 / 16.6 and 14.6 / 16.6.)
		uint32_t tps = 1;
		{
			const size_t smem2 = (size_t)m * CPT * sizeof(float) + (size_t)NBUF * 2 * mn * sizeof(Node8) + 128;
			if ((227 * 1024) / (smem2 + 1024) >= 5) tps = 2;
		}
		if (const char* e = getenv("PSM_PREDICT_TPS")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) tps = (uint32_t)v; }
		const size_t smemC = (size_t)m * CPT * sizeof(float) + (size_t)NBUF * tps * mn * sizeof(Node8) + 128;
		if (nodes8 && maxNodes > 0 && m <= 2047 && NC < (1u << 20) && smemC <= 100 * 1024 && !getenv("PSM_PREDICT_WIDE")) {
			const uint32_t rowTiles = (Nte + CPT - 1) / CPT;
			uint32_t ctasPerSm = (uint32_t)((227 * 1024) / (smemC + 1024));
			if (ctasPerSm < 1) ctasPerSm = 1;
			if (ctasPerSm > 8) ctasPerSm = 8;
			const uint32_t slots = 148 * ctasPerSm;
			uint32_t chunks = (6 * slots + rowTiles - 1) / rowTiles;
			if (chunks < 1) chunks = 1;
			if (chunks > nPartsBatch) chunks = nPartsBatch;
			const uint32_t pch = (nPartsBatch + chunks - 1) / chunks;
			chunks = (nPartsBatch + pch - 1) / pch;
			const dim3 grid(rowTiles, chunks);
			const Node8* n8 = reinterpret_cast<const Node8*>(nodes8);
			const bool direct = chunks == 1 && class12 != nullptr && !getenv("PSM_PREDICT_SCRATCH");
#if PSM_PREDICT_LTAB_GLOBAL
			leaf_tab_init_kernel<<<1, 128, 0, s>>>();     // idempotent (every launch writes the same values), so fold lanes need no lock
#endif
#define PSM_LAUNCH_COMPACT(K)                                                                                                              \
	do {                                                                                                                                   \
		if (direct) {                                                                                                                      \
			if (allow_max_dyn_smem(predict_compact_kernel<K, true>) != cudaSuccess) return -1;                                             \
			predict_compact_kernel<K, true><<<grid, CPT, smemC, s>>>(x, m, testIdx, Nte, nodes, n8, NC, ts, nPartsBatch, T, pch, mn, scratch, divider, class12);  \
		} else {                                                                                                                           \
			if (allow_max_dyn_smem(predict_compact_kernel<K, false>) != cudaSuccess) return -1;                                            \
			predict_compact_kernel<K, false><<<grid, CPT, smemC, s>>>(x, m, testIdx, Nte, nodes, n8, NC, ts, nPartsBatch, T, pch, mn, scratch, divider, class12); \
		}                                                                                                                                  \
	} while (0)
			if (tps == 8) PSM_LAUNCH_COMPACT(8);
			else if (tps == 4) PSM_LAUNCH_COMPACT(4);
			else if (tps == 2) PSM_LAUNCH_COMPACT(2);
			else PSM_LAUNCH_COMPACT(1);
#undef PSM_LAUNCH_COMPACT
			return direct ? 1 : 0;
		}
	}
	{
		// TMA-staged path when the row tile plus two tree buffers fit in shared memory
		const size_t tileB = (size_t)m * PT * sizeof(double);
		const size_t smemT = tileB + 2 * (size_t)maxNodes * sizeof(Node16) + 64;
		if (maxNodes > 0 && smemT <= 110 * 1024) {
			const uint32_t rowTiles = (Nte + PT - 1) / PT;
			uint32_t ctasPerSm = (uint32_t)((227 * 1024) / (smemT + 1024));
			if (ctasPerSm < 1) ctasPerSm = 1;
			if (ctasPerSm > 8) ctasPerSm = 8;
			const uint32_t slots = 148 * ctasPerSm;
			uint32_t chunks = (6 * slots + rowTiles - 1) / rowTiles;
			if (chunks < 1) chunks = 1;
			if (chunks > nPartsBatch) chunks = nPartsBatch;
			const uint32_t pch = (nPartsBatch + chunks - 1) / chunks;
			chunks = (nPartsBatch + pch - 1) / pch;
			if (allow_max_dyn_smem(predict_tma_kernel) != cudaSuccess) return -1;
			predict_tma_kernel<<<dim3(rowTiles, chunks), PT, smemT, s>>>(x, m, testIdx, Nte, nodes, NC, ts, nPartsBatch, T, pch, maxNodes, scratch);
			return 0;
		}
	}
	const size_t smem = (size_t)m * PT * sizeof(double);
	const bool useSmem = smem <= 200 * 1024;
	const uint32_t rowTiles = (Nte + PT - 1) / PT;
	uint32_t ctasPerSm = useSmem ? (uint32_t)((227 * 1024) / (smem + 1024)) : 8u;
	if (ctasPerSm < 1) ctasPerSm = 1;
	if (ctasPerSm > 8) ctasPerSm = 8;
	const uint32_t slots = 148 * ctasPerSm;
	uint32_t chunks = (6 * slots + rowTiles - 1) / rowTiles;
	if (chunks < 1) chunks = 1;
	if (chunks > nPartsBatch) chunks = nPartsBatch;
	const uint32_t pch = (nPartsBatch + chunks - 1) / chunks;
	chunks = (nPartsBatch + pch - 1) / pch;
	dim3 grid(rowTiles, chunks);
	if (useSmem) {
		if (smem > 48 * 1024)
			if (allow_max_dyn_smem(predict_kernel<true>) != cudaSuccess) return -1;
		predict_kernel<true><<<grid, PT, smem, s>>>(x, m, testIdx, Nte, nodes, NC, nPartsBatch, T, pch, scratch);
	} else {
		predict_kernel<false><<<grid, PT, 0, s>>>(x, m, testIdx, Nte, nodes, NC, nPartsBatch, T, pch, scratch);
	}
	return 0;
}

void launch_accumulate(const double* scratch, uint32_t Nte, uint32_t nPartsBatch, double divider, double* class12, cudaStream_t s) {
	if (Nte == 0 || nPartsBatch == 0) return;
	accumulate_kernel<<<(Nte + 255) / 256, 256, 0, s>>>(scratch, Nte, nPartsBatch, divider, class12);
}

}  // namespace psm
