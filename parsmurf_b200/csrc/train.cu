// K4/K5 forest training -- level-wise, node-parallel growth of every tree of a partition batch at once.
//
// Replaces ranger's per-tree recursive growth as rfRanger drives it (reference src/rfRanger.cpp:6-54,177-183):
//   Tree::bootstrap                         src/ranger/Tree/Tree.cpp:413-446      -> bootstrap_kernel
//   Tree::createPossibleSplitVarSubset      src/ranger/Tree/Tree.cpp:245-272      -> gen_candidates (rng.cuh)
//   TreeProbability::splitNodeInternal      src/ranger/Tree/TreeProbability.cpp:81-121
//   TreeProbability::findBestSplit{,ValueSmallQ,ValueLargeQ}   :141-358           -> split_search_kernel / split_select_kernel
//   Tree::splitNode (child ids, partition)  src/ranger/Tree/Tree.cpp:274-346      -> level_finalize_kernel / partition_lists_kernel
//   TreeProbability::addToTerminalNodes     src/ranger/Tree/TreeProbability.cpp:47-63
//
// Data layout (HBM, all per batch): for every tree and every feature a list of 8-byte entries
// (label | multiplicity | dense rank | row), ascending by feature value, containing the tree's in-bag rows.
// Every open node owns the same contiguous segment [start, start+cnt) in all m lists of its tree, so
//   * split search for (node, candidate feature) is ONE coalesced stream over cnt entries -- no gathers;
//     class counts are prefix-summed with warp shuffles, the Gini proxy sum_c L_c^2/n_L + sum_c R_c^2/n_R is
//     evaluated in fp64 exactly as ranger rounds it, arg-max reduced with shuffles (first candidate in draw
//     order, then lowest value, strict '>': TreeProbability.cpp:274,340);
//   * after the split decision every list segment is stably partitioned into the two child segments
//     (ping-pong buffers), which keeps all lists sorted without ever sorting again.
// Node ids are assigned breadth-first in parent-id order exactly like Tree.cpp:294-302, which is what makes
// the i-th mtry draw of the device stream belong to node i.
#include <cstdlib>
#include <type_traits>
#include "kernels.cuh"
#include "rng.cuh"

namespace psm {

// ------------------------------------------------------------------------------------------ bootstrap
// one warp per tree: replay mt19937_64((uint32)((t+1)*forestSeed)) and histogram Ntr accepted draws.
__global__ void __launch_bounds__(32) bootstrap_kernel(TrainArgs A, const uint32_t* __restrict__ forestSeeds) {
	__shared__ unsigned long long mt[MT_N];
	const uint32_t tree = blockIdx.x;
	const uint32_t pl = tree / A.T, t = tree % A.T;
	const uint32_t N = A.parts[pl].Ntr;
	const int lane = lane_id();
	const uint32_t treeSeed = (uint32_t)((unsigned long long)(t + 1) * forestSeeds[pl]);   // Forest.cpp:438
	mt_seed(mt, (unsigned long long)treeSeed);
	uint32_t* cnt = A.inbag + (size_t)tree * A.stride;
	uint32_t taken = 0, pos = MT_N;
	while (taken < N) {
		mt_twist(mt);
		pos = 0;
		for (int b = 0; b < MT_N && taken < N; b += 32) {
			int i = b + lane;
			unsigned long long v = 0;
			bool ok = false;
			if (i < MT_N) ok = lemire_accept(mt_temper(mt[i]), (unsigned long long)N, &v);
			unsigned acc = __ballot_sync(0xffffffffu, ok);
			uint32_t seq = taken + __popc(acc & lanemask_lt());
			if (ok && seq < N) atomicAdd(&cnt[(uint32_t)v], 1u);
			uint32_t nacc = __popc(acc);
			if (taken + nacc >= N) {
				// position just after the N-th accepted output
				uint32_t need = N - taken;   // 1..nacc
				unsigned m2 = acc;
				for (uint32_t k = 1; k < need; k++) m2 &= m2 - 1;
				pos = b + __ffs(m2);         // index of need-th set bit + 1
				taken = N;
			} else {
				taken += nacc;
				pos = (b + 32 < MT_N) ? b + 32 : MT_N;
			}
		}
	}
	unsigned long long* g = A.mt + (size_t)tree * MT_N;
	for (int i = lane; i < MT_N; i += 32) g[i] = mt[i];
	if (lane == 0) A.ts[tree].mtPos = pos;
}

// injected bootstrap: histogram of the reference's sampleIDs
__global__ void inbag_from_ids_kernel(TrainArgs A, const unsigned long long* __restrict__ ids, const uint64_t* __restrict__ idOffsets) {
	const uint32_t tree = blockIdx.y;
	const uint32_t pl = tree / A.T, t = tree % A.T;
	const uint32_t N = A.parts[pl].Ntr;
	const unsigned long long* src = ids + idOffsets[pl] + (size_t)t * N;
	uint32_t* cnt = A.inbag + (size_t)tree * A.stride;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) atomicAdd(&cnt[(uint32_t)src[i]], 1u);
}

// ------------------------------------------------------------------------------------------ per-tree lists
// one warp per (tree, feature): filter the partition's sorted (rank,row) list by in-bag count.
template <typename E>
__global__ void __launch_bounds__(128) build_lists_kernel(TrainArgs A) {
	const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (wid >= A.nTrees * A.m) return;
	const uint32_t tree = wid / A.m, f = wid % A.m;
	const uint32_t pl = tree / A.T;
	const PartDesc pd = A.parts[pl];
	const int lane = lane_id();
	const entry_t* S = A.Sall + pd.sOffset + (size_t)f * pd.Ntr;
	const uint32_t* cnt = A.inbag + (size_t)tree * A.stride;
	E* out = reinterpret_cast<E*>(A.lists[0]) + ((size_t)tree * A.m + f) * A.stride;
	uint32_t off = 0;
	bool bad = false;
	// four 32-entry groups per trip: the in-bag lookups are dependent gathers (ncu r01: 36% issue-active, L1-bound), so the loads
	// of all four groups are issued before the first compaction
	constexpr int BL_U = 4;
	const unsigned lt = lanemask_lt();
	for (uint32_t b = 0; b < pd.Ntr; b += 32 * BL_U) {
		entry_t e[BL_U];
		uint32_t c[BL_U];
#pragma unroll
		for (int u = 0; u < BL_U; u++) { const uint32_t i = b + u * 32 + lane; e[u] = (i < pd.Ntr) ? S[i] : 0; }
#pragma unroll
		for (int u = 0; u < BL_U; u++) { const uint32_t i = b + u * 32 + lane; c[u] = (i < pd.Ntr) ? cnt[e_row(e[u])] : 0; }
#pragma unroll
		for (int u = 0; u < BL_U; u++) {
			const uint32_t row = e_row(e[u]);
			const bool keep = c[u] > 0;
			if (c[u] > Ent<E>::kMaxCountE) bad = true;
			const unsigned mk = __ballot_sync(0xffffffffu, keep);
			if (keep) out[off + __popc(mk & lt)] = Ent<E>::make(row, e_rank(e[u]), c[u], row >= pd.nPosOut ? 1u : 0u);
			off += __popc(mk);
		}
	}
	if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&A.counters[CNT_ERROR], ERR_COUNT_LIMIT);
	if (f == 0 && lane == 0) A.ts[tree].nEntries = off;
}

// The same for all the trees of a partition at once (compact entries, T <= BLM_T): the in-bag counts are first transposed to
// [partition][row][BLM_T] bytes, so ONE 16-byte gather per list entry serves every tree, and one warp per (partition, feature)
// reads that feature's sorted list once and feeds T output lists.  The per-tree kernel above looks every row up once per tree
// and feature -- 250 M dependent 4-byte gathers per BASELINE configs[1] fold (ncu r01: 36% issue-active, L1-bound) -- and reads
// every S column T times.
constexpr uint32_t BLM_T = 16;
__global__ void __launch_bounds__(256) inbag_transpose_kernel(TrainArgs A, unsigned char* __restrict__ inbagT) {
	const uint32_t pl = blockIdx.y;
	const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= A.parts[pl].Ntr) return;
	uint32_t w[BLM_T / 4] = {0, 0, 0, 0};
	bool bad = false;
	for (uint32_t t = 0; t < A.T; t++) {
		const uint32_t c = A.inbag[((size_t)pl * A.T + t) * A.stride + row];
		if (c > Ent<uint32_t>::kMaxCountE) bad = true;
		w[t >> 2] |= min(c, 255u) << (8 * (t & 3));
	}
	if (bad) atomicOr(&A.counters[CNT_ERROR], ERR_COUNT_LIMIT);
	reinterpret_cast<uint4*>(inbagT)[(size_t)pl * A.stride + row] = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(128) build_lists_multi_kernel(TrainArgs A, const unsigned char* __restrict__ inbagT) {
	const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t nParts = A.nTrees / A.T;
	if (wid >= nParts * A.m) return;
	const uint32_t pl = wid / A.m, f = wid % A.m;
	const PartDesc pd = A.parts[pl];
	const int lane = lane_id();
	const unsigned lt = lanemask_lt();
	const entry_t* S = A.Sall + pd.sOffset + (size_t)f * pd.Ntr;
	const uint4* cnt = reinterpret_cast<const uint4*>(inbagT) + (size_t)pl * A.stride;
	uint32_t* out0 = reinterpret_cast<uint32_t*>(A.lists[0]) + ((size_t)pl * A.T * A.m + f) * A.stride;   // tree t: + t * m * stride
	const size_t treeStep = (size_t)A.m * A.stride;
	uint32_t off[BLM_T];
#pragma unroll
	for (uint32_t t = 0; t < BLM_T; t++) off[t] = 0;
	constexpr int U = 2;
	for (uint32_t b = 0; b < pd.Ntr; b += 32 * U) {
		entry_t e[U];
		uint4 c[U];
#pragma unroll
		for (int u = 0; u < U; u++) { const uint32_t i = b + u * 32 + lane; e[u] = (i < pd.Ntr) ? S[i] : 0; }
#pragma unroll
		for (int u = 0; u < U; u++) { const uint32_t i = b + u * 32 + lane; c[u] = (i < pd.Ntr) ? __ldg(cnt + e_row(e[u])) : make_uint4(0, 0, 0, 0); }
#pragma unroll
		for (int u = 0; u < U; u++) {
			const uint32_t row = e_row(e[u]), rank = e_rank(e[u]);
			const uint32_t w[4] = {c[u].x, c[u].y, c[u].z, c[u].w};
#pragma unroll
			for (uint32_t t = 0; t < BLM_T; t++) {
				if (t < A.T) {
					const uint32_t k = (w[t >> 2] >> (8 * (t & 3))) & 0xFFu;
					const unsigned mk = __ballot_sync(0xffffffffu, k > 0);
					if (k > 0) out0[t * treeStep + off[t] + __popc(mk & lt)] = Ent<uint32_t>::make(row, rank, k, 0);
					off[t] += __popc(mk);
				}
			}
		}
	}
	if (f == 0 && lane == 0) {
#pragma unroll
		for (uint32_t t = 0; t < BLM_T; t++) if (t < A.T) A.ts[pl * A.T + t].nEntries = off[t];
	}
}

// ------------------------------------------------------------------------------------------ candidates
// Draw the mtry split-variable candidates of ONE node from the tree's stream (warp-cooperative).
// out == nullptr: consume the draws without storing (trivially terminal node).  perm[] is iota(m) on entry and
// exit in Fisher-Yates mode, all-zero in rejection mode.
__device__ void gen_candidates(unsigned long long* mt, uint32_t& pos, uint32_t m, uint32_t mtry, uint32_t* perm, uint32_t* jbuf,
                               uint32_t* out) {
	const int lane = lane_id();
	const bool fisher = !(mtry < (m + 1) / 10);   // utility.cpp:86
	if (fisher) {
		if (out == nullptr) {
			uint32_t rem = mtry;
			while (rem) {
				if (pos == MT_N) { mt_twist(mt); pos = 0; }
				uint32_t take = min(rem, (uint32_t)MT_N - pos);
				pos += take; rem -= take;
			}
			return;
		}
		uint32_t q0 = 0;
		while (q0 < mtry) {
			if (pos == MT_N) { mt_twist(mt); pos = 0; }
			uint32_t take = min(min(32u, mtry - q0), (uint32_t)MT_N - pos);
			if ((uint32_t)lane < take) {
				uint32_t q = q0 + lane;
				double u = canonical01(mt_temper(mt[pos + lane]));
				double jd = __dadd_rn((double)q, __dmul_rn(u, (double)(m - q)));     // utility.cpp:133
				uint32_t j = (uint32_t)__double2ull_rz(jd);
				jbuf[q] = j < m ? j : m - 1;
			}
			pos += take; q0 += take;
		}
		__syncwarp();
		if (lane == 0)
			for (uint32_t q = 0; q < mtry; q++) { uint32_t j = jbuf[q]; uint32_t a = perm[q]; perm[q] = perm[j]; perm[j] = a; }
		__syncwarp();
		for (uint32_t q = lane; q < mtry; q += 32) out[q] = perm[q];
		__syncwarp();
		if (lane == 0)
			for (int q = (int)mtry - 1; q >= 0; q--) { uint32_t j = jbuf[q]; uint32_t a = perm[q]; perm[q] = perm[j]; perm[j] = a; }
		__syncwarp();
	} else {
		// drawWithoutReplacementSimple (utility.cpp:94-117), 32 stream outputs per step: every lane maps one output with the
		// Lemire step and tests it against the selection flags; among equal draws inside the step only the first lane can be
		// accepted; the step is cut right after the output that completes the mtry-th acceptance, exactly where the
		// sequential loop of the reference stops consuming the stream.
		uint32_t count = 0;
		while (count < mtry) {
			if (pos == MT_N) { mt_twist(mt); pos = 0; }
			const uint32_t take = min(32u, (uint32_t)MT_N - pos);
			unsigned long long v = 0;
			bool ok = false;
			if ((uint32_t)lane < take) ok = lemire_accept(mt_temper(mt[pos + lane]), (unsigned long long)m, &v);
			const uint32_t draw = (uint32_t)v;            // range [0, m-1]; the skipped column m is never reached (utility.cpp:107-111)
			const bool fresh = ok && perm[draw] == 0;     // not selected by an earlier step (utility.cpp:112)
			const unsigned peers = __match_any_sync(0xffffffffu, fresh ? draw : (0x80000000u | (uint32_t)lane));
			const bool cand = fresh && ((__ffs(peers) - 1) == lane);
			const unsigned acc = __ballot_sync(0xffffffffu, cand);
			const uint32_t need = mtry - count, nacc = __popc(acc);
			uint32_t last = take - 1;                     // last consumed lane of this step
			if (nacc >= need) {
				unsigned m2 = acc;
				for (uint32_t k = 1; k < need; k++) m2 &= m2 - 1;
				last = __ffs(m2) - 1;
			}
			if (cand && (uint32_t)lane <= last) {
				jbuf[count + __popc(acc & lanemask_lt())] = draw;
				perm[draw] = 1;
			}
			__syncwarp();
			count += min(nacc, need);
			pos += last + 1;
		}
		__syncwarp();
		if (out) for (uint32_t q = lane; q < mtry; q += 32) out[q] = jbuf[q];
		for (uint32_t q = lane; q < mtry; q += 32) perm[jbuf[q]] = 0;
		__syncwarp();
	}
}

// Fisher-Yates draw of up to 32 consecutive nodes at once, one node per lane (mtry <= FY_MAX).  Node c of the group uses the
// mtry stream outputs [c*mtry, (c+1)*mtry) staged in ubuf; the permutation is the identity with at most 2*mtry overrides,
// kept as a tiny per-lane (position -> value) map.  out[c] == nullptr: the node's draws are consumed but not stored.
constexpr uint32_t FY_MAX = 8;
__device__ __forceinline__ void fy_lane(const unsigned long long* ubuf, uint32_t m, uint32_t mtry, uint32_t* out) {
	uint32_t kpos[2 * FY_MAX], kval[2 * FY_MAX];
	uint32_t nk = 0;
	uint32_t res[FY_MAX];
#pragma unroll
	for (uint32_t q = 0; q < FY_MAX; q++) {
		if (q < mtry) {
			const double u = canonical01(ubuf[q]);
			const double jd = __dadd_rn((double)q, __dmul_rn(u, (double)(m - q)));     // utility.cpp:133
			uint32_t j = (uint32_t)__double2ull_rz(jd);
			if (j >= m) j = m - 1;
			// value currently at positions q and j
			uint32_t vq = q, vj = j, iq = 0xFFFFFFFFu, ij = 0xFFFFFFFFu;
			for (uint32_t t = 0; t < nk; t++) { if (kpos[t] == q) { vq = kval[t]; iq = t; } if (kpos[t] == j) { vj = kval[t]; ij = t; } }
			res[q] = vj;                                   // perm[q] after the swap
			if (j != q) {                                  // perm[j] = old perm[q]
				if (ij != 0xFFFFFFFFu) kval[ij] = vq; else { kpos[nk] = j; kval[nk] = vq; nk++; }
			}
			(void)iq;                                      // position q is never read again (later swaps touch positions > q)
		}
	}
	if (out) {
#pragma unroll
		for (uint32_t q = 0; q < FY_MAX; q++) if (q < mtry) out[q] = res[q];
	}
}

__device__ __forceinline__ void write_leaf(Node16* nd, uint32_t npos, uint32_t nneg) {
	double n = (double)(npos + nneg);                                     // TreeProbability.cpp:49-62
	Node16 v;
	v.a = __ddiv_rn((double)npos, n);
	v.b = (unsigned long long)__double_as_longlong(__ddiv_rn((double)nneg, n)) | 0x8000000000000000ull;
	*nd = v;
}

__device__ __forceinline__ bool node_live(uint32_t npos, uint32_t nneg) {
	return (npos + nneg > 10u) && npos > 0 && nneg > 0;                    // TreeProbability.cpp:84-105
}

// shared-memory carve-up of the one-warp-per-tree kernels
struct TreeSmem {
	unsigned long long* mt;
	unsigned long long* ubuf;   // 32 * FY_MAX staged stream outputs
	uint32_t* perm;
	uint32_t* jbuf;
	uint32_t* child;   // 64 entries
};
__device__ __forceinline__ TreeSmem carve(unsigned char* raw, uint32_t m, uint32_t mtry) {
	TreeSmem s;
	s.mt = (unsigned long long*)raw;
	s.ubuf = s.mt + MT_N;
	s.perm = (uint32_t*)(raw + (MT_N + 32 * FY_MAX) * 8);
	s.jbuf = s.perm + m;
	s.child = s.jbuf + mtry;
	return s;
}
size_t tree_smem_bytes(uint32_t m, uint32_t mtry) { return (MT_N + 32 * FY_MAX) * 8 + (size_t)(m + mtry + 64) * 4; }

__device__ __forceinline__ void init_perm(uint32_t* perm, uint32_t m, uint32_t mtry) {
	const bool fisher = !(mtry < (m + 1) / 10);
	for (uint32_t i = lane_id(); i < m; i += 32) perm[i] = fisher ? i : 0u;
	__syncwarp();
}

// root node of every tree: class totals, terminal test, first candidate draw, first work item
__global__ void __launch_bounds__(32) init_trees_kernel(TrainArgs A) {
	extern __shared__ __align__(16) unsigned char raw[];
	TreeSmem sm = carve(raw, A.m, A.mtry);
	const uint32_t tree = blockIdx.x;
	const uint32_t pl = tree / A.T;
	const PartDesc pd = A.parts[pl];
	const int lane = lane_id();
	const uint32_t* cnt = A.inbag + (size_t)tree * A.stride;
	uint32_t npos = 0, nneg = 0;
	for (uint32_t r = lane; r < pd.Ntr; r += 32) { uint32_t c = cnt[r]; if (r < pd.nPosOut) npos += c; else nneg += c; }
	for (int o = 16; o > 0; o >>= 1) { npos += __shfl_xor_sync(0xffffffffu, npos, o); nneg += __shfl_xor_sync(0xffffffffu, nneg, o); }
	TreeState ts = A.ts[tree];
	Node16* nodes = A.nodes + (size_t)tree * A.NC;
	const bool live = node_live(npos, nneg);
	uint32_t item = 0xFFFFFFFFu;
	uint32_t* candBase = A.cand;
	if (live) {
		if (lane == 0) item = atomicAdd(&A.counters[CNT_NEXT_ITEMS], 1u);
		item = __shfl_sync(0xffffffffu, item, 0);
		if (lane == 0) {
			Item it; it.tree = tree; it.node = 0; it.start = 0; it.cnt = ts.nEntries; it.npos = npos; it.nneg = nneg; it.nPosOut = pd.nPosOut; it.pl = pl;
			A.items[0][item] = it;
			const uint32_t k = A.useClasses ? size_class(ts.nEntries) : (uint32_t)(NCLS - 1);
			A.clsItems[0][k][atomicAdd(&A.counters[CNT_CLS0 + k], 1u)] = item;
		}
	} else if (lane == 0) {
		write_leaf(&nodes[0], npos, nneg);
	}
	if (A.injCand) {
		if (live) for (uint32_t q = lane; q < A.mtry; q += 32) candBase[(size_t)item * A.mtry + q] = A.injCand[((size_t)tree * A.injCandNodes + 0) * A.mtry + q];
	} else {
		unsigned long long* g = A.mt + (size_t)tree * MT_N;
		for (int i = lane; i < MT_N; i += 32) sm.mt[i] = g[i];
		init_perm(sm.perm, A.m, A.mtry);
		uint32_t pos = ts.mtPos;
		gen_candidates(sm.mt, pos, A.m, A.mtry, sm.perm, sm.jbuf, live ? candBase + (size_t)item * A.mtry : nullptr);
		for (int i = lane; i < MT_N; i += 32) g[i] = sm.mt[i];
		ts.mtPos = pos;
	}
	if (lane == 0) {
		ts.nNodes = 1; ts.itemBase = live ? item : 0; ts.itemCount = live ? 1 : 0; ts.overflow = 0;
		A.ts[tree] = ts;
	}
}

// ------------------------------------------------------------------------------------------ split search
// one warp per (open node, candidate feature): stream the node's segment of that feature's list.
// Each lane owns SS_IPL consecutive entries of a 32*SS_IPL-entry chunk (two 16-byte loads from a 16-byte aligned
// base: odd segment starts are handled by shifting the chunk grid one entry to the left and masking), so the class
// counts need ONE warp scan per chunk; the next chunk's loads are issued before the current chunk is evaluated.
// The exact fp64 score (two IEEE divisions) is only evaluated for positions whose fp32 estimate could still beat the
// lane's running best: estimate error < 1e-6 relative, filter margin 5e-6, so no winner is ever skipped.
#ifndef SS_MINB
#define SS_MINB 8
#endif
#ifndef SS_MINB_MULTI
#define SS_MINB_MULTI 7      // the chunk loop keeps ~75 registers live (two chunks of entries in flight); 7 CTAs x 128 threads x 73 registers fill the file
#endif

// entries per lane and chunk: the compact 4-byte entries take 8 (two 16-byte loads, 256-entry chunks) in the chunk loop of the
// long segments and 4 in the single-chunk classes; the 8-byte entries always 4
#ifndef PSM_SS_IPL_MULTI
#define PSM_SS_IPL_MULTI 8
#endif
#ifndef SS_BLOCK_DEFAULT
#define SS_BLOCK_DEFAULT 64     // threads per CTA of the merged launch (PSM_SS_BLOCK=32|64|128 overrides)
#endif
template <typename E, bool MULTI> struct SsIpl { static constexpr int v = (MULTI && sizeof(E) == 4) ? PSM_SS_IPL_MULTI : 4; };

template <int IPL> __device__ __forceinline__ void ss_vec(const unsigned long long* p, unsigned long long (&e)[IPL]) {
#pragma unroll
	for (int q = 0; q < IPL / 2; q++) {
		const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(p + 2 * q));
		e[2 * q] = a.x; e[2 * q + 1] = a.y;
	}
}
template <int IPL> __device__ __forceinline__ void ss_vec(const uint32_t* p, uint32_t (&e)[IPL]) {
#pragma unroll
	for (int q = 0; q < IPL / 4; q++) {
		const uint4 a = __ldg(reinterpret_cast<const uint4*>(p + 4 * q));
		e[4 * q] = a.x; e[4 * q + 1] = a.y; e[4 * q + 2] = a.z; e[4 * q + 3] = a.w;
	}
}
template <typename E, int IPL>
__device__ __forceinline__ void ss_load(const E* __restrict__ base, uint32_t v, uint32_t lo, uint32_t hi, E (&e)[IPL]) {
	// entries v..v+IPL-1 of the aligned grid; valid iff lo <= index < hi.  A vector that straddles a segment end is loaded whole
	// (16-byte pieces that hold at least one valid entry: the neighbouring entries belong to the same list or its padding, see
	// ss_load_single) and masked afterwards -- no per-entry loads.
	if (v >= lo && v + IPL <= hi) {
		ss_vec<IPL>(base + v, e);
	} else {
		constexpr int APE = 16 / (int)sizeof(E);
#pragma unroll
		for (int q = 0; q < IPL / APE; q++) {
			E t[APE];
#pragma unroll
			for (int j = 0; j < APE; j++) t[j] = (E)0;
			if (v + q * APE < hi && v + (q + 1) * APE > lo) ss_vec<APE>(base + v + q * APE, t);
#pragma unroll
			for (int j = 0; j < APE; j++) { const uint32_t i = v + q * APE + j; e[q * APE + j] = (i >= lo && i < hi) ? t[j] : (E)0; }
		}
	}
}

// single-chunk variant: the whole aligned vector is loaded whenever any of its entries is valid (the surrounding entries
// belong to neighbouring segments of the same list, or to the list's padding: start - off is a multiple of the vector
// length and the stride is a multiple of 4, so the load never leaves the allocation) and the invalid slots are zeroed
// afterwards -- no divergent per-entry path for the lanes that straddle the segment ends.
template <typename E, int IPL>
__device__ __forceinline__ void ss_load_single(const E* __restrict__ base, uint32_t v, uint32_t lo, uint32_t hi, E (&e)[IPL]) {
#pragma unroll
	for (int j = 0; j < IPL; j++) e[j] = (E)0;
	if (v < hi && v + IPL > lo) ss_vec<IPL>(base + v, e);
#pragma unroll
	for (int j = 0; j < IPL; j++) e[j] = (v + j >= lo && v + j < hi) ? e[j] : (E)0;
}

// Packed running class counts (positives in the high half, negatives in the low half of one word) and their conversion to
// fp32 WITHOUT the conversion unit: counts are below 2^23, so (count | 0x4B000000) read as a float is 2^23 + count exactly and
// one FADD gives the float.  For the 16+16 packing of the compact entries the two halves are picked with byte permutes (one
// ALU instruction each); I2F runs at a quarter of the ALU rate and the old kernel issued two per entry (ncu r01).
template <typename E> struct PkOf;
template <> struct PkOf<uint32_t> {
	typedef uint32_t type; static constexpr int shift = 16;
	static __device__ __forceinline__ float hi_f(uint32_t pk) { return __uint_as_float(__byte_perm(pk, 0x4B000000u, 0x7432)) - 8388608.0f; }
	static __device__ __forceinline__ float lo_f(uint32_t pk) { return __uint_as_float(__byte_perm(pk, 0x4B000000u, 0x7410)) - 8388608.0f; }
	// the same before the subtraction of 2^23 (the packed estimate subtracts two at a time)
	static __device__ __forceinline__ float hi_raw(uint32_t pk) { return __uint_as_float(__byte_perm(pk, 0x4B000000u, 0x7432)); }
	static __device__ __forceinline__ float lo_raw(uint32_t pk) { return __uint_as_float(__byte_perm(pk, 0x4B000000u, 0x7410)); }
	// "entries j and j+1 hold different values": ranks differ (bits 14..27)
	static __device__ __forceinline__ bool rank_differs(uint32_t a, uint32_t b) { return ((a ^ b) & 0x0FFFC000u) != 0u; }
};
template <> struct PkOf<unsigned long long> {
	typedef unsigned long long type; static constexpr int shift = 32;
	static __device__ __forceinline__ float hi_f(unsigned long long pk) { return __uint_as_float((uint32_t)(pk >> 32) | 0x4B000000u) - 8388608.0f; }
	static __device__ __forceinline__ float lo_f(unsigned long long pk) { return __uint_as_float((uint32_t)pk | 0x4B000000u) - 8388608.0f; }
	static __device__ __forceinline__ float hi_raw(unsigned long long pk) { return __uint_as_float((uint32_t)(pk >> 32) | 0x4B000000u); }
	static __device__ __forceinline__ float lo_raw(unsigned long long pk) { return __uint_as_float((uint32_t)pk | 0x4B000000u); }
	static __device__ __forceinline__ bool rank_differs(unsigned long long a, unsigned long long b) { return ((a ^ b) & 0x0000FFFFFF000000ull) != 0ull; }
};

// Packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): two positions of the split-score estimate per instruction.  Every lane of a
// pair is rounded exactly like the scalar operation, so the packed estimate is bit for bit the scalar one.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// MUFU.RCP, 1 ulp (__frcp_rn is the correctly rounded one: a Newton step and a guarded slow path around the same MUFU)
__device__ __forceinline__ float rcp_approx(float x) {
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// group collectives: all 32 lanes of a warp always take part (an idle group of a sub-warp instantiation recomputes the
// last work item and only skips the store), so the shuffles run with the full member mask and a segment width -- a partial
// runtime mask would make nvcc wrap every collective in a WARPSYNC/COLLECTIVE region
template <int G> __device__ __forceinline__ uint32_t grp_max(uint32_t v) {
	if constexpr (G == 32) return __reduce_max_sync(0xffffffffu, v);
	else {
#pragma unroll
		for (int o = G / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
		return v;
	}
}
template <int G> __device__ __forceinline__ uint32_t grp_min(uint32_t v) {
	if constexpr (G == 32) return __reduce_min_sync(0xffffffffu, v);
	else {
#pragma unroll
		for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
		return v;
	}
}

// G = lanes per (node, candidate) work item, MULTI = the segment may span several chunks of G*IPL entries.
// tree_level_kernel sorts the open nodes of a level into six size classes (kernels.cuh: size_class); every class has its own
// compact item list and the instantiation that fits it, and ONE launch covers them all, longest class first
// (split_search_merged_kernel below):
//   class 0 (<= 32 grid entries): G = 8, four work items per warp      class 1 (<= 64): G = 16, two per warp
//   class 2 (<= 128): G = 32, one chunk, straight-line code             classes 3-5: G = 32, chunk loop with prefetch
// At the mid and deep levels most nodes hold a few dozen rows and the kernel is instruction-issue bound (ncu r01: 557 warp
// instructions per (node, candidate) for an 89-entry segment, 65% issue-active), so the per-warp fixed cost -- work-item
// set-up, the scan, the exact block, the arg-max -- is what the small classes share between two or four items.
// (An earlier attempt launched the three G variants over ALL items and let the non-matching warps exit: the extra item
// loads and empty warps cost more than the packing saved; the compact per-class lists are what makes it pay.)
//
// The score of a split position.  ranger maximises  sum_c L_c^2/n_L + sum_c R_c^2/n_R  (TreeProbability.cpp:337), which for two
// classes equals  n - 2 g  with  g = L0 L1/n_L + R0 R1/n_R  >= 0: the arg-max of the score is the arg-min of g.  Every position gets
// an fp32 ESTIMATE of g in the single-division form  (L0 L1 n_R + R0 R1 n_L) * rcp(n_L n_R)  -- eight roundings and one approximate
// reciprocal, relative error below 1e-6 of the score -- and only the positions whose estimate is within 4e-6 of the best estimate
// of their chunk (or of the best exact score of earlier chunks) are evaluated exactly as ranger does, with its two IEEE fp64
// divisions.  r02: one MUFU per position instead of two, no integer-to-float conversions (PkOf), 256-entry chunks without
// range checks away from the segment ends, the exact block and the pruning bound behind ONE vote per chunk.
// Measured alternatives that did NOT pay at BASELINE configs[1] (r01, split search stayed at ~14.1 ms per CV in every case):
//   * one warp per node walking its mtry candidates (set-up and arg-max paid once, next candidate prefetched, filter bound
//     carried across candidates): 25% fewer instructions per node, 15.5-17.6 ms -- the five-fold loss of independent warps
//     costs more than the instructions saved;
//   * a single shared exact-evaluation block fed by a per-lane pending mask instead of one block per lane slot: same time
//     (only the taken blocks ever executed);
//   * no software pipeline (40 registers, 12 CTAs per SM instead of 8): same time -- what the extra warps hide, the exposed
//     load of every second chunk gives back;
//   * (r02) the four warps of a CTA sharing one long segment (> 1021 entries: a quarter each, class counts of the quarters
//     exchanged through shared memory after a counting pass), to shorten the 24-chunk serial walk of the first levels and the
//     half-empty second wave of 5 000 warps on 4 144 slots: 11.9 ms re-reading the quarter through L1 / L2, 11.9 ms with the
//     quarter staged in shared memory, against 10.9 ms -- ncu shows the one-warp loop ~75% issue-bound while its warps are
//     resident (1.3 warp instructions per entry), so the extra counting pass and the barrier cost more than the tail wave.
// tIdx = the thread's index among the threads of its size class (G threads per work item)
template <typename E, int G, bool MULTI>
__device__ __forceinline__ void split_search_body(const TrainArgs& A, const uint32_t* __restrict__ list, uint32_t nList, int buf, size_t tIdx, uint32_t c) {
	constexpr int IPL = SsIpl<E, MULTI>::v;
	typedef PkOf<E> PK;
	typedef typename PK::type pk_t;
	constexpr int PKS = PK::shift;
	constexpr pk_t LOMASK = ((pk_t)1 << PKS) - 1;
	// grid.y = candidate index, so no integer division is needed to find the work item
	uint32_t li = (uint32_t)(tIdx / G);
	const bool idle = li >= nList;
	if (G == 32) { if (idle) return; }
	else if (idle) li = nList - 1;                            // keep the warp whole (see grp_max)
	const uint32_t item = __ldg(list + li);
	const Item it = A.items[buf][item];
	constexpr uint32_t APE = 16 / sizeof(E);                  // entries per aligned 16-byte vector
	const uint32_t off = it.start % APE;                      // A.stride % 4 == 0, so list + start - off is 16-byte aligned
	const uint32_t lo = off, hi = it.cnt + off;                // valid index range in the aligned grid
	const int lane = lane_id();
	const int gl = lane % G;
	const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - gl));   // this group's lanes (ballot filter only)
	constexpr unsigned FULL = 0xffffffffu;
	const uint32_t f = A.cand[(size_t)item * A.mtry + c];
	const uint32_t nPosOut = it.nPosOut;
	const E* base = reinterpret_cast<const E*>(A.lists[A.listBuf]) + ((size_t)it.tree * A.m + f) * A.stride + (it.start - off);
	const uint32_t n = it.npos + it.nneg;
	const float fP = (float)it.npos, fN = (float)it.nneg, fn = (float)n;
	pk_t absBase = 0;                                         // class counts left of the current chunk, packed
	double best = -1.0;
	float wbound = 0.0f;                                      // group-uniform: (best exact score of earlier chunks) * (1 - 4e-6)
	uint32_t bpos = 0xFFFFFFFFu;
	pk_t babs = 0;
	E e[IPL], nx[IPL];
	if (MULTI) ss_load<E, IPL>(base, gl * IPL, lo, hi, e);
	else ss_load_single<E, IPL>(base, gl * IPL, lo, hi, e);
	uint32_t b = 0;
	do {
		const uint32_t v0 = b + gl * IPL;
		if (MULTI) ss_load<E, IPL>(base, v0 + G * IPL, lo, hi, nx);   // prefetch the next chunk
		else {
#pragma unroll
			for (int j = 0; j < IPL; j++) nx[j] = (E)0;
		}
		// lane-local inclusive prefix of (positives, negatives) packed in one word: 16+16 bits for the compact entries
		// (a 256-entry chunk holds at most 256*15 samples, a whole tree at most 16384), 32+32 bits for the wide ones
		pk_t pk[IPL];
		pk_t run = 0;
#pragma unroll
		for (int j = 0; j < IPL; j++) {
			const pk_t w = (pk_t)Ent<E>::count(e[j]);             // 0 for masked entries
			run += Ent<E>::neg(e[j], nPosOut) ? w : (w << PKS);
			pk[j] = run;
		}
		pk_t inc = run;
#pragma unroll
		for (int o = 1; o < G; o <<= 1) {
			pk_t t = __shfl_up_sync(FULL, inc, o, G);
			if (gl >= o) inc += t;
		}
		const pk_t excl = absBase + inc - run;
		// the entry that follows this lane's last one: first entry of the next lane, or of the next chunk for the last lane
		E enext = __shfl_down_sync(FULL, e[0], 1, G);
		if (MULTI) { const E c0 = __shfl_sync(FULL, nx[0], 0, G); if (gl == G - 1) enext = c0; }
		else if (gl == G - 1) enext = (E)0;
		// chunks that touch a segment end check the position range; the others run without (MULTI only: a single chunk always does)
		const bool edge = !MULTI || (b == 0 && off != 0) || (b + G * IPL >= hi);
		float g[IPL];
		float gmin = __int_as_float(0x7F800000);
		auto estimate = [&](auto edgeTag) {
			constexpr bool EDGE = decltype(edgeTag)::value;
			static_assert(IPL % 2 == 0, "the estimate works on pairs of positions");
			const f32x2_t kM23 = pk2(-8388608.0f, -8388608.0f), kM1 = pk2(-1.0f, -1.0f), fn2 = pk2(fn, fn), fP2 = pk2(fP, fP), fN2 = pk2(fN, fN);
#pragma unroll
			for (int j = 0; j < IPL; j += 2) {
				// g = (L0 L1 n_R + R0 R1 n_L) * rcp(n_L n_R) for positions j and j + 1, the operations of the scalar form
				//   fmaf(fl * fq, nr, ((fP - fl) * (fN - fq)) * nl) * rcp(nl * nr)   in the same order, two at a time
				const pk_t a0 = excl + pk[j], a1 = excl + pk[j + 1];
				const f32x2_t fl = add2(pk2(PK::hi_raw(a0), PK::hi_raw(a1)), kM23), fq = add2(pk2(PK::lo_raw(a0), PK::lo_raw(a1)), kM23);
				const f32x2_t nl = add2(fl, fq), nr = fma2(nl, kM1, fn2);
				const f32x2_t num = fma2(mul2(fl, fq), nr, mul2(mul2(fma2(fl, kM1, fP2), fma2(fq, kM1, fN2)), nl));
				float d0, d1;
				up2(mul2(nl, nr), d0, d1);
				float gp[2];
				up2(mul2(num, pk2(rcp_approx(d0), rcp_approx(d1))), gp[0], gp[1]);
#pragma unroll
				for (int q = 0; q < 2; q++) {
					const int jj = j + q;
					bool ok = PK::rank_differs(e[jj], jj + 1 < IPL ? e[(jj + 1) % IPL] : enext);
					if (EDGE) { const uint32_t v = v0 + jj; ok = ok && v >= lo && v + 1 < hi; }
					const float gj = ok ? gp[q] : __int_as_float(0x7F800000);
					g[jj] = gj;
					gmin = fminf(gmin, gj);
				}
			}
		};
		if (edge) estimate(std::true_type()); else estimate(std::false_type());
		// Only positions whose estimate is within the estimate's error (< 1e-6 relative; margin 4e-6) of the best
		// estimate of this chunk, or of the best exact score of earlier chunks, can be the arg-max: evaluate those
		// exactly (two IEEE fp64 divisions, TreeProbability.cpp:337).  Estimates are >= 0 (or +inf), so their bit patterns
		// order like unsigned integers and one redux gives the group minimum.
		const float cmin = __uint_as_float(grp_min<G>(__float_as_uint(gmin)));
		const float thrScore = fmaxf(fmaf(-2.0f, cmin, fn) * 0.999996f, wbound);
		const float gthr = (fn - thrScore) * 0.5f;                // score >= thrScore  <=>  g <= gthr
		bool cand = false;
#pragma unroll
		for (int j = 0; j < IPL; j++) cand = cand || (g[j] <= gthr);
		if (__ballot_sync(FULL, cand) & gm) {
			if (cand) {
#pragma unroll
				for (int j = 0; j < IPL; j++) {
					if (g[j] <= gthr) {
						const pk_t a = excl + pk[j];
						const uint32_t lpi = (uint32_t)(a >> PKS), lni = (uint32_t)(a & LOMASK);
						const uint32_t nl = lpi + lni, nr = n - nl;
						const uint32_t rp = it.npos - lpi, rn = it.nneg - lni;
						const unsigned long long sli = (unsigned long long)lpi * lpi + (unsigned long long)lni * lni;
						const unsigned long long sri = (unsigned long long)rp * rp + (unsigned long long)rn * rn;
						const double score = __dadd_rn(__ddiv_rn((double)sri, (double)nr), __ddiv_rn((double)sli, (double)nl));
						if (score > best) { best = score; bpos = v0 + j - off; babs = a; }
					}
				}
			}
			if (MULTI) {
				// group-uniform bound for the following chunks: a later position must strictly beat the best exact score so far
				const float bf = best > 0.0 ? __double2float_rd(best) : 0.0f;
				wbound = __uint_as_float(grp_max<G>(__float_as_uint(bf))) * 0.999996f;
			}
		}
		if (MULTI) {
			absBase += __shfl_sync(FULL, inc, G - 1, G);
#pragma unroll
			for (int j = 0; j < IPL; j++) e[j] = nx[j];
		}
		b += G * IPL;
	} while (MULTI && b < hi);
	// arg-max over the group's lanes: highest score, then lowest position (ascending value order, strict '>').
	// Scores are positive doubles, so their two 32-bit halves order like unsigned integers: three redux steps.
	uint32_t blp = (uint32_t)(babs >> PKS), bln = (uint32_t)(babs & LOMASK);
	{
		const unsigned long long sb = best > 0.0 ? (unsigned long long)__double_as_longlong(best) : 0ull;
		const uint32_t mhi = grp_max<G>((uint32_t)(sb >> 32));
		const bool c1 = (uint32_t)(sb >> 32) == mhi && sb != 0ull;
		const uint32_t mlo = grp_max<G>(c1 ? (uint32_t)sb : 0u);
		const bool c2 = c1 && (uint32_t)sb == mlo;
		const uint32_t mpos = grp_min<G>(c2 ? bpos : 0xFFFFFFFFu);
		const unsigned win = __ballot_sync(FULL, c2 && bpos == mpos) & gm;
		// (the shuffles are unconditional: groups of one warp may disagree on `win`)
		const int src = win ? __ffs(win) - 1 : lane;
		const double wbest = __shfl_sync(FULL, best, src);
		const uint32_t wlp = __shfl_sync(FULL, blp, src), wln = __shfl_sync(FULL, bln, src);
		if (win) { best = wbest; blp = wlp; bln = wln; bpos = mpos; }
		else { best = -1.0; bpos = 0xFFFFFFFFu; }
	}
	if (gl == 0 && !idle) {
		SearchRes r; r.score = best; r.pos = bpos; r.lpos = blp; r.lneg = bln; r.pad = 0;
		A.res[(size_t)item * A.mtry + c] = r;
	}
}

// one size class per launch (PSM_SS_SEPARATE=1: the r01 / early r02 form, kept for A/B measurements)
template <typename E, int G, bool MULTI>
__global__ void __launch_bounds__(128, MULTI ? SS_MINB_MULTI : SS_MINB) split_search_kernel(TrainArgs A, const uint32_t* __restrict__ list, uint32_t nList, int buf) {
	split_search_body<E, G, MULTI>(A, list, nList, buf, (size_t)blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y);   // grid.y = candidate
}

// All the size classes of a level in ONE launch.  The CTAs are laid out largest class first (the three chunk-loop classes in
// descending size, then the single-chunk warp class, the 16-lane and the 8-lane groups), so the longest segments start first and
// the short ones fill the slots they leave (longest-processing-time-first over ~10^5 work items whose lengths span 10 .. 6 000
// entries at the mid levels), and a level pays one launch instead of up to four: with one launch per class every class ran its
// own ramp and its own half-empty last wave (ncu r02: warps active 30% of the 44% the registers allow in the chunk-loop class,
// 6-8 us launch floors for each of the small classes on the ~15 deepest levels).  The kernel has no barrier and no shared
// memory, so the CTA size is free: SsPlan.block threads (a multiple of 32) -- a CTA only retires with its slowest warp, so small
// CTAs keep the slots of finished warps from idling next to a long segment.
struct SsPlan {
	uint32_t ctaEnd[NCLS];    // exclusive CTA bounds of the classes in launch order (class NCLS-1 first)
	uint32_t count[NCLS];     // items of the class, same order
};
// Registers: the chunk loop wants 72 (SS_MINB_MULTI), but with every class in one launch the resident warps are what hides the
// latency of the short segments: measured on B200 at BASELINE configs[1] (split search ms per CV, 64-thread CTAs):
// 72 registers 8.97, 64: 8.33, 56: 7.96, 48 (15 spill stores + 15 reloads per chunk): 7.87, 40: 7.76; one launch per class (six
// classes): 11.95; CTAs of 32 / 128 threads at 72 registers: 9.23 / 9.16; class cuts at 254 / 1022 entries instead of 510 / 2046:
// the same; 128-entry chunks at 48 / 40 registers: 8.27 / 7.92; no prefetch of the next chunk at 48 / 40 registers: 7.91 / 7.79.
// Below ~56 registers everything lands within 2%: ncu (profiles/r02_summary.md) shows the merged kernel at 62-71% issue-active
// with 27-35 resident warps per SM -- what is left is the instruction count (~290 warp instructions per 256-entry chunk).
#ifndef SS_MINB_MERGED32
#define SS_MINB_MERGED32 10     // 4-byte entries: 48 registers
#endif
#ifndef SS_MINB_MERGED64
#define SS_MINB_MERGED64 8      // 8-byte entries: 64 registers
#endif
template <typename E>
__global__ void __launch_bounds__(128, sizeof(E) == 4 ? SS_MINB_MERGED32 : SS_MINB_MERGED64) split_search_merged_kernel(TrainArgs A, SsPlan P, int buf) {
	// grid = (candidate, work-item CTAs [, overflow of the 65535 limit]): the candidate is the fastest dimension, so CTAs are handed
	// out in work-item order -- longest class first -- with the mtry candidates of an item side by side
	const uint32_t bx = blockIdx.z * gridDim.y + blockIdx.y, c = blockIdx.x;
	// the per-level counters (next level's item count and class sizes) were read by the host before this launch and are next
	// touched by this level's tree_level_kernel: cleared here instead of by a memset of their own
	static_assert(CNT_NEXT_ITEMS == 0 && CNT_CLS0 == 1, "the per-level counters are contiguous");
	if (bx == 0 && c == 0 && threadIdx.x < 1 + NCLS) A.counters[threadIdx.x] = 0;
	static_assert(NCLS == 6, "launch order below: classes 5, 4, 3 (chunk loop), 2 (one warp, one chunk), 1 (16 lanes), 0 (8 lanes)");
	if (bx < P.ctaEnd[2]) {
		const int q = bx < P.ctaEnd[0] ? 0 : bx < P.ctaEnd[1] ? 1 : 2;
		const uint32_t first = q ? P.ctaEnd[q - 1] : 0u;
		split_search_body<E, 32, true>(A, A.clsItems[buf][NCLS - 1 - q], P.count[q], buf, (size_t)(bx - first) * blockDim.x + threadIdx.x, c);
	} else if (bx < P.ctaEnd[3]) {
		split_search_body<E, 32, false>(A, A.clsItems[buf][2], P.count[3], buf, (size_t)(bx - P.ctaEnd[2]) * blockDim.x + threadIdx.x, c);
	} else if (bx < P.ctaEnd[4]) {
		split_search_body<E, 16, false>(A, A.clsItems[buf][1], P.count[4], buf, (size_t)(bx - P.ctaEnd[3]) * blockDim.x + threadIdx.x, c);
	} else if (bx < P.ctaEnd[5]) {
		split_search_body<E, 8, false>(A, A.clsItems[buf][0], P.count[5], buf, (size_t)(bx - P.ctaEnd[4]) * blockDim.x + threadIdx.x, c);
	}
}

// eight lanes per open node (four nodes per warp): pick the winning candidate, compute the threshold, describe the list partition.
// The work per node is a short chain of dependent loads (results -> list entries -> feature values), so the kernel is bound by
// how many chains are in flight: with one warp per node the 60 000 nodes of a mid level of BASELINE configs[1] needed six waves.
constexpr int SEL_G = 8;
template <typename E>
__global__ void __launch_bounds__(128) split_select_kernel(TrainArgs A, uint32_t nItems, int buf) {
	uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) / SEL_G;
	const bool idle = item >= nItems;
	if (idle) item = nItems - 1;                          // keep the warp whole for the shuffles; an idle group only skips its stores
	const Item it = A.items[buf][item];
	const int gl = lane_id() % SEL_G;
	double best = -1.0;
	uint32_t bc = 0xFFFFFFFFu;
	for (uint32_t c = gl; c < A.mtry; c += SEL_G) {
		double s = A.res[(size_t)item * A.mtry + c].score;
		if (s > best) { best = s; bc = c; }          // ascending c per lane: first maximum stays
	}
#pragma unroll
	for (int o = SEL_G / 2; o > 0; o >>= 1) {
		double os = __shfl_xor_sync(0xffffffffu, best, o, SEL_G);
		uint32_t oc = __shfl_xor_sync(0xffffffffu, bc, o, SEL_G);
		if (os > best || (os == best && oc < bc)) { best = os; bc = oc; }
	}
	if (idle || gl != 0) return;
	Node16* node = A.nodes + (size_t)it.tree * A.NC + it.node;
	SplitDec d; d.split = 0; d.nleft = 0; d.lpos = 0; d.lneg = 0;
	if (!(best >= 0.0)) {                             // TreeProbability.cpp:184-186: no candidate had two distinct values
		write_leaf(node, it.npos, it.nneg); A.dec[item] = d;
		PartTask pt; pt.tree = it.tree; pt.start = it.start; pt.cnt = 0; pt.nleft = 0;
		A.ptask[item] = pt;
		return;
	}
	const SearchRes r = A.res[(size_t)item * A.mtry + bc];
	const uint32_t f = A.cand[(size_t)item * A.mtry + bc];
	const E* list = reinterpret_cast<const E*>(A.lists[A.listBuf]) + ((size_t)it.tree * A.m + f) * A.stride + it.start;
	const PartDesc pd = A.parts[it.pl];
	const double* col = A.Dall + pd.dOffset + (size_t)f * pd.Ntr;
	const E e0 = list[r.pos], e1 = list[r.pos + 1];
	const double lo = col[Ent<E>::row(e0)], hi = col[Ent<E>::row(e1)];
	double thr = __dmul_rn(__dadd_rn(lo, hi), 0.5);     // TreeProbability.cpp:348
	if (thr == hi) thr = lo;                            // :353-355
	Node16 v; v.a = thr; v.b = (unsigned long long)f;   // left child id is filled in by level_finalize
	*node = v;
	d.split = 1; d.nleft = r.pos + 1; d.lpos = r.lpos; d.lneg = r.lneg;
	A.dec[item] = d;
	// list partition: a terminal child's entries are never read again (level_finalize writes its leaf payload from
	// the class counts), so they are not moved; with two terminal children the node is skipped entirely
	const bool deadL = !node_live(d.lpos, d.lneg), deadR = !node_live(it.npos - d.lpos, it.nneg - d.lneg);
	PartTask pt; pt.tree = it.tree; pt.start = it.start; pt.nleft = d.nleft;
	pt.cnt = (deadL && deadR) ? 0u : (it.cnt | (deadL ? PT_DEAD_LEFT : 0u) | (deadR ? PT_DEAD_RIGHT : 0u));
	A.ptask[item] = pt;
	A.splitFeat[item] = f;
}

// ------------------------------------------------------------------------------------------ go-left flags
// One CTA per tree: bit r of the tree's bitmap = "in-bag row r goes to the left child of its node".  Only the left part
// [0, nleft) of every split node's segment in the winning feature's list is read; the bits are set with shared-memory
// atomics and the bitmap leaves the SM as one coalesced store.  (Per-row byte stores straight to global memory were bound
// by scattered-store throughput: ~60 us per level for the 6M rows of a BASELINE configs[1] fold.)
constexpr int TF_THREADS = 128;   // 1 bookkeeping warp + 3 flag warps; 8 CTAs per SM keep the 1000 trees of a BASELINE configs[1] batch in one wave
constexpr uint32_t TF_LONG = 1024;   // segments with at least this many left entries are shared by all the flag threads
// the go-left bitmap of one tree, by threads ftid = 0 .. FT-1 (whole warps) of its CTA, which meet at named barrier 1
template <typename E>
__device__ __forceinline__ void tree_flags_body(const TrainArgs& A, uint32_t tree, const TreeState& ts, uint32_t* bm_s, int bitmapInSmem, int ftid, int FT) {
	uint32_t* out = A.flagBits + (size_t)tree * A.flagWords;
	uint32_t* bm = bitmapInSmem ? bm_s : out;
	const int lane = ftid & 31, warp = ftid >> 5;
	for (uint32_t w = ftid; w < A.flagWords; w += FT) bm[w] = 0;
	asm volatile("bar.sync 1, %0;" ::"r"(FT) : "memory");
	const PartTask* tasks = A.ptask + ts.itemBase;
	const uint32_t* feats = A.splitFeat + ts.itemBase;
	const E* lists = reinterpret_cast<const E*>(A.lists[A.listBuf]) + (size_t)tree * A.m * A.stride;
	auto mark = [&](const E* seg, uint32_t nleft, uint32_t first, uint32_t step) {
		for (uint32_t b = first; b < nleft; b += 4 * step) {
			E e4[4];
#pragma unroll
			for (int k = 0; k < 4; k++) { const uint32_t i = b + k * step; e4[k] = (i < nleft) ? __ldg(seg + i) : (E)0; }
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t i = b + k * step;
				if (i < nleft) { const uint32_t r = Ent<E>::row(e4[k]); atomicOr(&bm[r >> 5], 1u << (r & 31)); }
			}
		}
	};
	// long segments (top levels): all threads; short ones: one warp each
	uint32_t nshort = 0;
	for (uint32_t i = 0; i < ts.itemCount; i++) {
		const uint4 raw = __ldg(reinterpret_cast<const uint4*>(tasks + i));   // {tree, start, cnt, nleft}
		if (raw.z == 0) continue;                                             // node became a leaf
		if (raw.w >= TF_LONG) mark(lists + (size_t)__ldg(feats + i) * A.stride + raw.y, raw.w, ftid, FT);
		else nshort++;
	}
	if (nshort) {
		for (uint32_t i = warp; i < ts.itemCount; i += FT / 32) {
			const uint4 raw = __ldg(reinterpret_cast<const uint4*>(tasks + i));
			if (raw.z == 0 || raw.w >= TF_LONG) continue;
			mark(lists + (size_t)__ldg(feats + i) * A.stride + raw.y, raw.w, lane, 32);
		}
	}
	if (bitmapInSmem) {
		asm volatile("bar.sync 1, %0;" ::"r"(FT) : "memory");
		for (uint32_t w = ftid; w < A.flagWords; w += FT) out[w] = bm_s[w];
	}
}

// ------------------------------------------------------------------------------------------ level finalize
// The mtry draws of the `nchild` nodes created by up to 32 parents, in node-id order (ids childBase .. childBase + nchild - 1);
// sm.child[c] = work-item slot of child c (its candidates go to candBase + slot * mtry) or 0xFFFFFFFF for a terminal child,
// whose draws are consumed but not stored.  Warp-cooperative.
__device__ __forceinline__ void draw_children(const TrainArgs& A, const TreeSmem& sm, uint32_t& pos, uint32_t tree, uint32_t childBase,
                                              uint32_t nchild, uint32_t* candBase, const int lane) {
	const bool laneFY = !A.injCand && A.mtry <= FY_MAX && !(A.mtry < (A.m + 1) / 10);
	if (laneFY) {
		// one node per lane: stage the next 32*mtry stream outputs (tempered) and draw up to 32 nodes at once
		for (uint32_t c0 = 0; c0 < nchild; c0 += 32) {
			const uint32_t nc = min(32u, nchild - c0);
			uint32_t need = nc * A.mtry, got = 0;
			while (got < need) {
				if (pos == MT_N) { mt_twist(sm.mt); pos = 0; }
				const uint32_t take = min(need - got, (uint32_t)MT_N - pos);
				for (uint32_t q = lane; q < take; q += 32) sm.ubuf[got + q] = mt_temper(sm.mt[pos + q]);
				got += take; pos += take;
			}
			__syncwarp();
			if ((uint32_t)lane < nc) {
				const uint32_t ci = sm.child[c0 + lane];
				fy_lane(sm.ubuf + (size_t)lane * A.mtry, A.m, A.mtry, ci != 0xFFFFFFFFu ? candBase + (size_t)ci * A.mtry : nullptr);
			}
			__syncwarp();
		}
	} else {
		for (uint32_t c = 0; c < nchild; c++) {
			const uint32_t ci = sm.child[c];
			if (A.injCand) {
				if (ci != 0xFFFFFFFFu)
					for (uint32_t q = lane; q < A.mtry; q += 32)
						candBase[(size_t)ci * A.mtry + q] = A.injCand[((size_t)tree * A.injCandNodes + (childBase + c)) * A.mtry + q];
			} else {
				gen_candidates(sm.mt, pos, A.m, A.mtry, sm.perm, sm.jbuf, ci != 0xFFFFFFFFu ? candBase + (size_t)ci * A.mtry : nullptr);
			}
		}
	}
	__syncwarp();
}

// one warp per tree: breadth-first child ids in parent-id order, child work items, leaf payloads of trivially
// terminal children, and the mtry draw of every new node in node-id order.
__device__ __forceinline__ void level_finalize_body(const TrainArgs& A, const TreeSmem& sm, const uint32_t tree, TreeState ts, const int buf) {
	const int lane = lane_id();
	const unsigned lt = lanemask_lt();
	const Item* cur = A.items[buf] + ts.itemBase;
	const SplitDec* dec = A.dec + ts.itemBase;
	Node16* nodes = A.nodes + (size_t)tree * A.NC;

	// pass 1: count splits and live children
	uint32_t nsplit = 0, nlive = 0;
	unsigned long long searched = 0;   // samples of the nodes that reached findBestSplit this level (SURVEY 8(d) byte count; one atomic per tree)
	for (uint32_t b = 0; b < ts.itemCount; b += 32) {
		uint32_t i = b + lane;
		uint32_t l = 0, s = 0;
		if (i < ts.itemCount) {
			SplitDec d = dec[i];
			const Item it = cur[i];
			searched += it.npos + it.nneg;
			if (d.split) {
				s = 1;
				l = (node_live(d.lpos, d.lneg) ? 1u : 0u) + (node_live(it.npos - d.lpos, it.nneg - d.lneg) ? 1u : 0u);
			}
		}
		for (int o = 16; o > 0; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); s += __shfl_xor_sync(0xffffffffu, s, o); }
		nlive += l; nsplit += s;
	}
	for (int o = 16; o > 0; o >>= 1) searched += __shfl_xor_sync(0xffffffffu, searched, o);
	if (lane == 0) atomicAdd(&A.stats[STAT_SPLIT_BYTES], searched * ((unsigned long long)A.mtry * 8ull + 8ull));
	if (nsplit == 0) {
		if (lane == 0) { ts.itemCount = 0; A.ts[tree] = ts; }
		return;
	}
	uint32_t nextBase = 0;
	if (lane == 0) nextBase = atomicAdd(&A.counters[CNT_NEXT_ITEMS], nlive);
	nextBase = __shfl_sync(0xffffffffu, nextBase, 0);
	if (nextBase + nlive > A.itemCap || ts.nNodes + 2 * nsplit > A.NC) {
		if (lane == 0) { atomicOr(&A.counters[CNT_ERROR], ERR_POOL_OVERFLOW); ts.itemCount = 0; ts.overflow = 1; A.ts[tree] = ts; }
		return;
	}
	Item* nextItems = A.items[buf ^ 1];
	uint32_t* candBase = A.cand;
	uint32_t pos = ts.mtPos;
	unsigned long long* g = A.mt + (size_t)tree * MT_N;
	if (!A.injCand) {
		for (int i = lane; i < MT_N; i += 32) sm.mt[i] = g[i];
		init_perm(sm.perm, A.m, A.mtry);
	}
	// pass 2
	uint32_t childBase = ts.nNodes, liveOff = 0;
	for (uint32_t b = 0; b < ts.itemCount; b += 32) {
		uint32_t i = b + lane;
		SplitDec d; d.split = 0; d.nleft = 0; d.lpos = 0; d.lneg = 0;
		Item it; it.tree = tree; it.node = 0; it.start = 0; it.cnt = 0; it.npos = 0; it.nneg = 0; it.nPosOut = 0; it.pl = 0;
		if (i < ts.itemCount) { d = dec[i]; it = cur[i]; }
		const unsigned smask = __ballot_sync(0xffffffffu, d.split != 0);
		const uint32_t sp = __popc(smask & lt);
		const uint32_t leftId = childBase + 2 * sp;
		const uint32_t rpos = it.npos - d.lpos, rneg = it.nneg - d.lneg;
		const bool liveL = d.split && node_live(d.lpos, d.lneg);
		const bool liveR = d.split && node_live(rpos, rneg);
		uint32_t l = (liveL ? 1u : 0u) + (liveR ? 1u : 0u);
		uint32_t inc = l;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		const uint32_t myOff = nextBase + liveOff + inc - l;
		if (d.split) {
			nodes[it.node].b = (nodes[it.node].b & 0xFFFFFFFFull) | ((unsigned long long)leftId << 32);   // Tree.cpp:294-302
			uint32_t o = myOff;
			if (liveL) {
				Item c; c.tree = tree; c.node = leftId; c.start = it.start; c.cnt = d.nleft; c.npos = d.lpos; c.nneg = d.lneg; c.nPosOut = it.nPosOut; c.pl = it.pl;
				nextItems[o] = c; sm.child[2 * sp] = o; o++;
			} else { write_leaf(&nodes[leftId], d.lpos, d.lneg); sm.child[2 * sp] = 0xFFFFFFFFu; }
			if (liveR) {
				Item c; c.tree = tree; c.node = leftId + 1; c.start = it.start + d.nleft; c.cnt = it.cnt - d.nleft; c.npos = rpos; c.nneg = rneg; c.nPosOut = it.nPosOut; c.pl = it.pl;
				nextItems[o] = c; sm.child[2 * sp + 1] = o;
			} else { write_leaf(&nodes[leftId + 1], rpos, rneg); sm.child[2 * sp + 1] = 0xFFFFFFFFu; }
		}
		{
			// size-class lists of the next level (one warp-aggregated atomic per class and batch of 32 parents)
			const uint32_t big = (uint32_t)(NCLS - 1);
			const uint32_t kL = liveL ? (A.useClasses ? size_class(d.nleft) : big) : (uint32_t)NCLS;
			const uint32_t kR = liveR ? (A.useClasses ? size_class(it.cnt - d.nleft) : big) : (uint32_t)NCLS;
			const uint32_t oL = myOff, oR = myOff + (liveL ? 1u : 0u);
#pragma unroll
			for (uint32_t k = 0; k < (uint32_t)NCLS; k++) {
				const unsigned mL = __ballot_sync(0xffffffffu, kL == k), mR = __ballot_sync(0xffffffffu, kR == k);
				const uint32_t tot = __popc(mL) + __popc(mR);
				if (tot == 0) continue;
				uint32_t cb = 0;
				if (lane == 0) cb = atomicAdd(&A.counters[CNT_CLS0 + k], tot);
				cb = __shfl_sync(0xffffffffu, cb, 0);
				uint32_t* cl = A.clsItems[buf ^ 1][k];
				if (kL == k) cl[cb + __popc(mL & lt)] = oL;
				if (kR == k) cl[cb + __popc(mL) + __popc(mR & lt)] = oR;
			}
		}
		__syncwarp();
		const uint32_t nchild = 2 * __popc(smask);
		draw_children(A, sm, pos, tree, childBase, nchild, candBase, lane);
		childBase += nchild;
		liveOff += __shfl_sync(0xffffffffu, inc, 31);
	}
	if (!A.injCand) for (int i = lane; i < MT_N; i += 32) g[i] = sm.mt[i];
	if (lane == 0) {
		if (A.injCand && childBase > A.injCandNodes) atomicOr(&A.counters[CNT_ERROR], ERR_INJ_CAND);
		ts.nNodes = childBase; ts.itemBase = nextBase; ts.itemCount = nlive; ts.mtPos = pos;
		A.ts[tree] = ts;
		atomicAdd(&A.stats[STAT_NODES], (unsigned long long)(2 * nsplit));
	}
}

// One CTA per tree and level, after split_select: warp 0 runs the level's bookkeeping (level_finalize_body) while the other
// warps build the go-left bitmap the list partition reads (tree_flags_body).  Both only depend on split_select's output and used
// to be two launches in a row, each one a 20-40 us latency chain per level (ncu / PSM_TRACE_LEVELS r02: together as long as
// split search itself at BASELINE configs[1]); merged, the two chains overlap and a level costs one launch less.
template <typename E>
__global__ void __launch_bounds__(TF_THREADS, 8) tree_level_kernel(TrainArgs A, int buf, int bitmapInSmem, uint32_t bitmapOffset) {
	extern __shared__ __align__(16) unsigned char raw[];
	const uint32_t tree = blockIdx.x;
	const TreeState ts = A.ts[tree];
	if (ts.itemCount == 0) return;
	__syncthreads();                        // every warp holds this level's item range before warp 0 overwrites A.ts[tree]
	if (threadIdx.x < 32) level_finalize_body(A, carve(raw, A.m, A.mtry), tree, ts, buf);
	else tree_flags_body<E>(A, tree, ts, reinterpret_cast<uint32_t*>(raw + bitmapOffset), bitmapInSmem, (int)threadIdx.x - 32, TF_THREADS - 32);
}

// ------------------------------------------------------------------------------------------ partition
// one warp per (split node, feature): stable two-way partition of the node's segment into the other buffer; grid.y = feature.
// Each trip of the chain (entry load -> flag word -> store) carries PL_CH independent 32-entry chunks.  History of the
// measurements at BASELINE configs[1] (ms per CV): with nvcc's first code (53 instructions per chunk, issue-bound) PL_CH
// 1: 39.7, 2: 30.8, 3: 29.6, 4: 31.4, 8: 44.8, and serving several features per warp was slower still (more, smaller warps
// win at every level); after the rewrite for instruction count 3: 25.0, 4: 22.6, 6 on the first levels only: 22.7.
#ifndef PSM_PL_CH
#define PSM_PL_CH 4
#endif
constexpr int PL_CH_DEFAULT = PSM_PL_CH;
// r02, measured and not kept (partition ms per BASELINE configs[1] CV, 20.9 as built): requesting the next group's entries before
// the current group's flag words (two groups of loads in flight per warp) 23.3 with PL_CH 4 (40 registers), 21.5 / 23.4 with
// PL_CH 3 / 2; CTAs of 64 threads 25.3, 256: 21.7, 512: 23.7, 1024: 31.3 (a CTA's warps take neighbouring segments of one list).
#ifndef PSM_PL_BLOCK
#define PSM_PL_BLOCK 128        // threads per CTA (no barrier, no shared memory: any multiple of 32)
#endif

// the last (< NCH * 32) entries of a segment, fully predicated; NCH = chunk slots compiled in.  PRED = some child of the node is
// terminal and its entries are not written (keepL / keepR); without it the two flags are compile-time true.
template <typename E, int NCH, bool PRED>
__device__ __forceinline__ void partition_tail(const E* srcp, E* dstp, const uint32_t* fbits, uint32_t b, uint32_t cnt, uint32_t lane, unsigned lt,
                                               uint32_t offL, uint32_t offR, bool keepL, bool keepR) {
	E e[NCH];
	uint32_t w[NCH];
#pragma unroll
	for (int c = 0; c < NCH; c++) e[c] = (b + c * 32 + lane < cnt) ? __ldg(srcp + c * 32) : (E)0;
#pragma unroll
	for (int c = 0; c < NCH; c++) w[c] = __ldg(fbits + (Ent<E>::row(e[c]) >> 5));
#pragma unroll
	for (int c = 0; c < NCH; c++) {
		const uint32_t first = b + c * 32;                 // uniform
		const bool valid = first + lane < cnt;
		const bool gl = valid && ((w[c] >> (Ent<E>::row(e[c]) & 31u)) & 1u);
		const unsigned mL = __ballot_sync(0xffffffffu, gl);
		const uint32_t pL = __popc(mL & lt), nL = __popc(mL);
		const uint32_t nv = first < cnt ? min(32u, cnt - first) : 0u;
		const uint32_t idx = gl ? offL + pL : offR + lane - pL;
		if (valid && (!PRED || (gl ? keepL : keepR))) dstp[idx] = e[c];
		offL += nL; offR += nv - nL;
	}
}

template <typename E, int PL_CH, bool PRED>
__device__ __forceinline__ void partition_segment(const E* srcp, E* dstp, const uint32_t* fbits, uint32_t cnt, uint32_t nleft, uint32_t lane, unsigned lt,
                                                  bool keepL, bool keepR) {
	// Short segments (the mid and deep levels: ncu r01 shows 180 warp instructions for an 89-entry segment because all
	// PL_CH predicated slots of the tail are issued) take a tail with only as many chunk slots as they need.
	if (cnt <= 32) { partition_tail<E, 1, PRED>(srcp, dstp, fbits, 0, cnt, lane, lt, 0, nleft, keepL, keepR); return; }
	if (cnt <= 64) { partition_tail<E, 2, PRED>(srcp, dstp, fbits, 0, cnt, lane, lt, 0, nleft, keepL, keepR); return; }
	uint32_t offL = 0, offR = nleft;                           // next free slot of the left / right child segment
	uint32_t b = 0;
	// full groups of PL_CH * 32 entries: no bounds logic at all (what the long segments of the first levels run)
	for (; b + 32 * PL_CH <= cnt; b += 32 * PL_CH, srcp += 32 * PL_CH) {
		E e[PL_CH];
		uint32_t w[PL_CH];
#pragma unroll
		for (int c = 0; c < PL_CH; c++) e[c] = __ldg(srcp + c * 32);
#pragma unroll
		for (int c = 0; c < PL_CH; c++) w[c] = __ldg(fbits + (Ent<E>::row(e[c]) >> 5));
#pragma unroll
		for (int c = 0; c < PL_CH; c++) {
			const bool gl = (w[c] >> (Ent<E>::row(e[c]) & 31u)) & 1u;
			const unsigned mL = __ballot_sync(0xffffffffu, gl);
			const uint32_t pL = __popc(mL & lt), nL = __popc(mL);
			if (!PRED || (gl ? keepL : keepR)) dstp[gl ? offL + pL : offR + lane - pL] = e[c];
			offL += nL; offR += 32 - nL;
		}
	}
	if (b < cnt) {
		if (cnt - b <= 64) partition_tail<E, 2, PRED>(srcp, dstp, fbits, b, cnt, lane, lt, offL, offR, keepL, keepR);
		else partition_tail<E, PL_CH, PRED>(srcp, dstp, fbits, b, cnt, lane, lt, offL, offR, keepL, keepR);
	}
}

template <typename E, int PL_CH>
__global__ void __launch_bounds__(PSM_PL_BLOCK) partition_lists_kernel(TrainArgs A, uint32_t nItems, int buf) {
	const uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (item >= nItems) return;
	const uint4 raw = __ldg(reinterpret_cast<const uint4*>(A.ptask + item));
	// PartTask: cnt carries two flags in its top bits -- a child that is terminal (<= 10 samples or pure) is never read
	// again, so its entries are not written; a node with two terminal children is skipped altogether (split_select)
	const uint32_t tree = raw.x, start = raw.y, cnt = raw.z & PT_CNT_MASK, nleft = raw.w;
	if (cnt == 0) return;                                      // node became a leaf, or neither child will be searched
	const bool keepL = !(raw.z & PT_DEAD_LEFT), keepR = !(raw.z & PT_DEAD_RIGHT);
	const uint32_t f = blockIdx.y;
	const uint32_t lane = lane_id();
	const unsigned lt = lanemask_lt();
	const size_t lo = ((size_t)tree * A.m + f) * A.stride + start;
	const E* srcp = reinterpret_cast<const E*>(A.lists[A.listBuf]) + lo + lane;
	E* dstp = reinterpret_cast<E*>(A.lists[A.listBuf ^ 1]) + lo;
	const uint32_t* fbits = A.flagBits + (size_t)tree * A.flagWords;   // go-left flags, bit-packed per tree (tree_flags_kernel)
	// keep both bases in registers: otherwise every store / flag load re-derives its 64-bit address from the kernel arguments
	asm volatile("" : "+l"(dstp));
	asm volatile("" : "+l"(fbits));
	// The kernel is issue-bound (ncu r01: 72% issue-active at 44% of DRAM peak, 53 instructions per 32-entry chunk), so the
	// loop is written for instruction count: 32-bit indices off two base pointers, one ballot per chunk (the valid lanes of a
	// chunk are a prefix, so a lane's rank among the right-goers is lane - rank among the left-goers), one store per entry,
	// and an unpredicated flag load (an invalid lane carries entry 0 -> row 0, a legal address).  The store predicate of the
	// terminal-child case costs ~20% more instructions in the main loop (ncu r01), so nodes whose children both live on --
	// all the long segments of the first levels -- run the copy of the code without it.
	if (keepL && keepR) partition_segment<E, PL_CH, false>(srcp, dstp, fbits, cnt, nleft, lane, lt, true, true);
	else partition_segment<E, PL_CH, true>(srcp, dstp, fbits, cnt, nleft, lane, lt, keepL, keepR);
}

// ------------------------------------------------------------------------------------------ launchers
void launch_bootstrap(const TrainArgs& A, const uint32_t* forestSeeds, cudaStream_t s) {
	bootstrap_kernel<<<A.nTrees, 32, 0, s>>>(A, forestSeeds);
}
void launch_inbag_from_ids(const TrainArgs& A, const unsigned long long* ids, const uint64_t* idOffsets, cudaStream_t s) {
	dim3 grid((A.stride + 255) / 256, A.nTrees);
	inbag_from_ids_kernel<<<grid, 256, 0, s>>>(A, ids, idOffsets);
}
void launch_build_lists(const TrainArgs& A, unsigned char* inbagT, cudaStream_t s) {
	if (A.entry32 && inbagT && A.T <= BLM_T) {
		const uint32_t nParts = A.nTrees / A.T;
		inbag_transpose_kernel<<<dim3((A.stride + 255) / 256, nParts), 256, 0, s>>>(A, inbagT);
		const size_t warps = (size_t)nParts * A.m;
		build_lists_multi_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, s>>>(A, inbagT);
		return;
	}
	size_t warps = (size_t)A.nTrees * A.m;
	if (A.entry32) build_lists_kernel<uint32_t><<<(unsigned)((warps + 3) / 4), 128, 0, s>>>(A);
	else build_lists_kernel<unsigned long long><<<(unsigned)((warps + 3) / 4), 128, 0, s>>>(A);
}
int launch_init_trees(const TrainArgs& A, cudaStream_t s) {
	size_t smem = tree_smem_bytes(A.m, A.mtry);
	if (smem > 48 * 1024) {
		if (allow_max_dyn_smem(init_trees_kernel) != cudaSuccess) return -1;
	}
	init_trees_kernel<<<A.nTrees, 32, smem, s>>>(A);
	return 0;
}
// clsCount[k] = items in clsItems[buf][k] (kernels.cuh: size_class).  Returns the number of launches.
static int ss_env_block() {
	const char* e = getenv("PSM_SS_BLOCK");
	const int b = e ? atoi(e) : SS_BLOCK_DEFAULT;
	return (b == 32 || b == 64 || b == 128) ? b : SS_BLOCK_DEFAULT;
}
int launch_split_search(const TrainArgs& A, const uint32_t* clsCount, int buf, cudaStream_t s) {
	const bool separate = getenv("PSM_SS_SEPARATE") != nullptr;
	if (!separate) {
		const uint32_t block = (uint32_t)ss_env_block();
		SsPlan P;
		uint32_t end = 0;
		for (int q = 0; q < NCLS; q++) {
			const int k = NCLS - 1 - q;
			const uint32_t G = k == 0 ? 8u : k == 1 ? 16u : 32u;
			P.count[q] = clsCount[k];
			end += (uint32_t)(((size_t)clsCount[k] * G + block - 1) / block);
			P.ctaEnd[q] = end;
		}
		if (end == 0) return 0;
		constexpr uint32_t GY = 32768;                                       // grid.y <= 65535
		const dim3 grid(A.mtry, end < GY ? end : GY, (end + GY - 1) / GY);   // x: candidate; (z, y): work-item CTAs, largest class first
		if (A.entry32) split_search_merged_kernel<uint32_t><<<grid, block, 0, s>>>(A, P, buf);
		else split_search_merged_kernel<unsigned long long><<<grid, block, 0, s>>>(A, P, buf);
		return 1;
	}
	int launches = 0;
	for (int k = NCLS - 1; k >= 0; k--) {
		const uint32_t nl = clsCount[k];
		if (nl == 0) continue;
		launches++;
		const uint32_t* list = A.clsItems[buf][k];
		const uint32_t G = k == 0 ? 8u : k == 1 ? 16u : 32u;
		const dim3 grid((unsigned)(((size_t)nl * G + 127) / 128), A.mtry);   // x: nodes (G lanes each), y: candidate
#define PSM_SS(E)                                                                                            \
	do {                                                                                                     \
		if (k == 0) split_search_kernel<E, 8, false><<<grid, 128, 0, s>>>(A, list, nl, buf);                 \
		else if (k == 1) split_search_kernel<E, 16, false><<<grid, 128, 0, s>>>(A, list, nl, buf);           \
		else if (k == 2) split_search_kernel<E, 32, false><<<grid, 128, 0, s>>>(A, list, nl, buf);           \
		else split_search_kernel<E, 32, true><<<grid, 128, 0, s>>>(A, list, nl, buf);                        \
	} while (0)
		if (A.entry32) PSM_SS(uint32_t);
		else PSM_SS(unsigned long long);
#undef PSM_SS
	}
	return launches;
}
void launch_split_select(const TrainArgs& A, uint32_t nItems, int buf, cudaStream_t s) {
	const unsigned grid = (unsigned)(((size_t)nItems * SEL_G + 127) / 128);
	if (A.entry32) split_select_kernel<uint32_t><<<grid, 128, 0, s>>>(A, nItems, buf);
	else split_select_kernel<unsigned long long><<<grid, 128, 0, s>>>(A, nItems, buf);
}
// tree_flags + level_finalize of one level (tree_level_kernel)
int launch_tree_level(const TrainArgs& A, int buf, cudaStream_t s) {
	const size_t fin = (tree_smem_bytes(A.m, A.mtry) + 15) & ~(size_t)15;
	const size_t bytes = (size_t)A.flagWords * 4;
	const int inSmem = fin + bytes <= 200 * 1024;   // ~1.5M rows per tree; beyond that the bitmap is updated in place in global memory
	const size_t smem = fin + (inSmem ? bytes : 0);
	if (A.entry32) {
		if (smem > 48 * 1024 && allow_max_dyn_smem(tree_level_kernel<uint32_t>) != cudaSuccess) return -1;
		tree_level_kernel<uint32_t><<<A.nTrees, TF_THREADS, smem, s>>>(A, buf, inSmem, (uint32_t)fin);
	} else {
		if (smem > 48 * 1024 && allow_max_dyn_smem(tree_level_kernel<unsigned long long>) != cudaSuccess) return -1;
		tree_level_kernel<unsigned long long><<<A.nTrees, TF_THREADS, smem, s>>>(A, buf, inSmem, (uint32_t)fin);
	}
	return 0;
}
// one launch over all work items in item order (neighbouring segments of a list stay neighbours in time: launching the
// partition per size class like split search measured 8% slower -- the short segments lose the sector sharing with their
// neighbours -- so the per-size variants are dispatched inside the kernel instead)
void launch_partition_lists(const TrainArgs& A, uint32_t nItems, int buf, cudaStream_t s) {
	const dim3 grid((unsigned)(((size_t)nItems * 32 + PSM_PL_BLOCK - 1) / PSM_PL_BLOCK), A.m);   // x: nodes (one warp each), y: feature
	if (A.entry32) partition_lists_kernel<uint32_t, PL_CH_DEFAULT><<<grid, PSM_PL_BLOCK, 0, s>>>(A, nItems, buf);
	else partition_lists_kernel<unsigned long long, PL_CH_DEFAULT><<<grid, PSM_PL_BLOCK, 0, s>>>(A, nItems, buf);
}

}  // namespace psm
