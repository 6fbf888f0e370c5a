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

int launch_predict(const double* x, uint32_t m, const uint32_t* testIdx, uint32_t Nte, const Node16* nodes, uint32_t NC,
                   const TreeState* ts, uint32_t maxNodes, uint32_t nPartsBatch, uint32_t T, double* scratch, cudaStream_t s) {
	if (Nte == 0 || nPartsBatch == 0) return 0;
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
			if (cudaFuncSetAttribute(predict_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemT) != cudaSuccess) return -1;
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
			if (cudaFuncSetAttribute(predict_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
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
