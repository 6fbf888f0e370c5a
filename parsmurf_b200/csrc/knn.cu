// K1 knn_fold -- exact brute-force kNN over the fold's positive examples.
//
// Replaces ANNkd_tree + ANNkd_tree::annkSearch as SamplerKNN::minOversample uses them
// (reference src/samplerKNN.cpp:60-93; leaf scan src/ann_1.1.2/src/kd_search.cpp:172-210;
// ANN_ALLOW_SELF_MATCH = false, src/ann_1.1.2/include/ANN/ANN.h:235).
//
// Distances are squared L2 accumulated in fp64 in ascending feature order with separately rounded
// subtract / multiply / add (no FMA), i.e. the bit pattern ANN's annDist loop produces, so the neighbour
// ORDER (which SMOTE indexes by rank, samplerKNN.cpp:95) is the reference's; exact ties are broken by the
// lower index like ANNbruteForce (src/ann_1.1.2/src/brute.cpp:54-78).
// Caveats: (1) among EXACTLY equidistant neighbours the reference's kd-tree keeps the first one its traversal visits
// (ANNmin_k::insert), which depends on the tree layout, not on the index -- with duplicated or discrete feature vectors the
// rank of two equidistant positives (and which of them falls inside k) may therefore differ from the reference's; continuous
// data (every BASELINE config) has no such ties.  (2) The reference searches over m+1 coordinates, the label column included
// (src/samplerKNN.cpp:60-66); every positive carries the same label value 1.0 there (src/fileImport.cpp:88-91), so the extra
// coordinate adds exactly 0.0 to every distance and the m-coordinate distances here are the same doubles.
//
// Two kernels: a 64x64-tiled pairwise distance kernel (fp64 CUDA cores; B200 keeps full-rate FP64) and a
// warp-per-query top-k selection with register-resident sorted lists merged by warp shuffles.
#include "kernels.cuh"

namespace psm {

constexpr int KT = 64;   // tile edge
constexpr int KC = 16;   // feature chunk

__global__ void __launch_bounds__(256) knn_dist_kernel(const double* __restrict__ Xp, int P, int m, int q0, int nq,
                                                        double* __restrict__ D) {
	__shared__ double As[KC][KT];
	__shared__ double Bs[KC][KT];
	const int tid = threadIdx.x;
	const int tx = tid & 15, ty = tid >> 4;
	const int jt = blockIdx.x * KT, qt = blockIdx.y * KT;
	double acc[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) acc[i][j] = 0.0;

	const int lk = tid & 15, lr = tid >> 4;   // loader: 16 features x 16 rows per pass, 4 passes
	for (int k0 = 0; k0 < m; k0 += KC) {
#pragma unroll
		for (int i = 0; i < 4; i++) {
			int r = lr + 16 * i;
			int k = k0 + lk;
			int q = q0 + qt + r, j = jt + r;
			As[lk][r] = (q < q0 + nq && q < P && k < m) ? Xp[(size_t)q * m + k] : 0.0;
			Bs[lk][r] = (j < P && k < m) ? Xp[(size_t)j * m + k] : 0.0;
		}
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < KC; kk++) {
			double a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; i++) a[i] = As[kk][ty + 16 * i];
#pragma unroll
			for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) {
					double t = __dsub_rn(a[i], b[j]);
					acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(t, t));
				}
		}
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		int q = qt + ty + 16 * i;
		if (q >= nq) continue;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			int jj = jt + tx + 16 * j;
			if (jj < P) D[(size_t)q * P + jj] = acc[i][j];
		}
	}
}

template <int K>
__global__ void __launch_bounds__(128) knn_select_kernel(const double* __restrict__ D, int P, int q0, int nq, int localk,
                                                          int32_t* __restrict__ nnIdx, double* __restrict__ nnDist) {
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (warp >= nq) return;
	const double* row = D + (size_t)warp * P;
	const double INF = __longlong_as_double(0x7FF0000000000000ll);
	double ld[K];
	int lj[K];
#pragma unroll
	for (int i = 0; i < K; i++) { ld[i] = INF; lj[i] = -1; }
	for (int j = lane; j < P; j += 32) {
		double d = row[j];
		if (d == 0.0) continue;            // self and exact duplicates are never neighbours
		if (d < ld[K - 1]) {               // strict: an equal distance with a higher index stays behind
			ld[K - 1] = d; lj[K - 1] = j;
#pragma unroll
			for (int i = K - 1; i > 0; i--) {
				if (ld[i] < ld[i - 1]) {
					double td = ld[i]; ld[i] = ld[i - 1]; ld[i - 1] = td;
					int tj = lj[i]; lj[i] = lj[i - 1]; lj[i - 1] = tj;
				}
			}
		}
	}
	// merge the 32 sorted lists: localk rounds of warp arg-min over the heads, order (distance, index)
	const size_t o = (size_t)(q0 + warp) * localk;
	for (int r = 0; r < localk; r++) {
		double bd = ld[0];
		int bj = lj[0];
		unsigned long long key_d = (unsigned long long)__double_as_longlong(bd);   // non-negative doubles order as integers
		unsigned bjj = (bj < 0) ? 0xFFFFFFFFu : (unsigned)bj;
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) {
			unsigned long long od = __shfl_xor_sync(0xffffffffu, key_d, off);
			unsigned oj = __shfl_xor_sync(0xffffffffu, bjj, off);
			if (od < key_d || (od == key_d && oj < bjj)) { key_d = od; bjj = oj; }
		}
		if ((unsigned)((bj < 0) ? 0xFFFFFFFFu : (unsigned)bj) == bjj && bjj != 0xFFFFFFFFu) {
			// this lane owns the winner: pop its head
#pragma unroll
			for (int i = 0; i < K - 1; i++) { ld[i] = ld[i + 1]; lj[i] = lj[i + 1]; }
			ld[K - 1] = INF; lj[K - 1] = -1;
		}
		if (lane == 0) {
			nnIdx[o + r] = (bjj == 0xFFFFFFFFu) ? -1 : (int32_t)bjj;
			if (nnDist) nnDist[o + r] = (bjj == 0xFFFFFFFFu) ? INF : __longlong_as_double((long long)key_d);
		}
	}
}

// gather the fold's positive rows (features only) into a dense [P][m] matrix
__global__ void gather_rows_kernel(const double* __restrict__ x, uint32_t ld, uint32_t m, const uint32_t* __restrict__ idx,
                                   uint32_t count, double* __restrict__ out) {
	size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	size_t total = (size_t)count * m;
	if (t >= total) return;
	uint32_t r = (uint32_t)(t / m), c = (uint32_t)(t % m);
	out[t] = x[(size_t)idx[r] * ld + c];
}

void launch_gather_rows(const double* x, uint32_t ld, uint32_t m, const uint32_t* idx, uint32_t count, double* out, cudaStream_t s) {
	size_t total = (size_t)count * m;
	if (!total) return;
	gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, ld, m, idx, count, out);
}

int launch_knn(const double* Xp, int P, int m, int localk, double* Dscratch, size_t scratchElems, int32_t* nnIdx, double* nnDist,
               cudaStream_t s) {
	if (localk > 64) return -1;
	int qb = (int)(scratchElems / (size_t)P);
	if (qb > P) qb = P;
	if (qb < 1) return -2;
	for (int q0 = 0; q0 < P; q0 += qb) {
		int nq = (P - q0 < qb) ? (P - q0) : qb;
		dim3 grid((P + KT - 1) / KT, (nq + KT - 1) / KT);
		knn_dist_kernel<<<grid, 256, 0, s>>>(Xp, P, m, q0, nq, Dscratch);
		int blocks = (nq * 32 + 127) / 128;
		if (localk <= 8) knn_select_kernel<8><<<blocks, 128, 0, s>>>(Dscratch, P, q0, nq, localk, nnIdx, nnDist);
		else if (localk <= 16) knn_select_kernel<16><<<blocks, 128, 0, s>>>(Dscratch, P, q0, nq, localk, nnIdx, nnDist);
		else if (localk <= 32) knn_select_kernel<32><<<blocks, 128, 0, s>>>(Dscratch, P, q0, nq, localk, nnIdx, nnDist);
		else knn_select_kernel<64><<<blocks, 128, 0, s>>>(Dscratch, P, q0, nq, localk, nnIdx, nnDist);
	}
	return 0;
}

}  // namespace psm
