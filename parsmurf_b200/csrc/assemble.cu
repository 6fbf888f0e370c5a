// K2 assemble_smote -- fused partition assembly: gather positives, generate SMOTE synthetics, gather the
// partition's negative chunk, and write the training matrix directly in ranger's column-major layout.
//
// Replaces, for a whole batch of partitions in one launch:
//   SamplerKNN::minOversample / majUndersample   (reference src/samplerKNN.cpp:32-141)
//   Sampler::copySample / setSample              (src/sampler.cpp:37-49)
//   transposeMatrix + DataDouble copy            (src/hyperSMURF.cpp:289-291, src/HyperSMURFUtils.cpp:115-120)
//
// SMOTE arithmetic is the reference's, rounding step by step (samplerKNN.cpp:98-101):
//   alpha = rand() / (double)RAND_MAX;  out = nbr*alpha + q*(1-alpha)      -- no FMA contraction.
// Draw layout per synthetic row (i,j): 1 draw for the neighbour rank, then m draws for the alphas.
//
// HBM-bound: each 32x32 tile reads 32 row segments of 256 B (coalesced along features) and writes
// 32 column segments of 256 B (coalesced along rows) through a padded shared-memory transpose.
#include "kernels.cuh"

namespace psm {

__global__ void __launch_bounds__(256) assemble_kernel(const double* __restrict__ x, uint32_t m, const uint32_t* __restrict__ posIdx,
                                                        uint32_t P, const uint32_t* __restrict__ negIdx,
                                                        const int32_t* __restrict__ nn, uint32_t localk, uint32_t fp,
                                                        const int32_t* __restrict__ draws, const PartDesc* __restrict__ parts,
                                                        double* __restrict__ Dall, int* __restrict__ err) {
	__shared__ double tile[32][33];
	const PartDesc pd = parts[blockIdx.z];
	const uint32_t ld = m + 1;
	const uint32_t r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
	if (r0 >= pd.Ntr) return;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps, 4 rows each
	const int32_t* dr = draws + pd.drawOffset;
	for (int rr = ty; rr < 32; rr += 8) {
		uint32_t r = r0 + rr, c = c0 + tx;
		double v = 0.0;
		if (r < pd.Ntr && c < ld) {
			if (r < P) {
				v = x[(size_t)posIdx[r] * ld + c];
			} else if (r < pd.nPosOut) {
				if (c == m) {
					v = 1.0;
				} else {
					uint32_t s = r - P;            // synthetic index = i*fp + j
					uint32_t i = s / fp;
					size_t base = (size_t)s * ld;  // 1 + m draws per synthetic row
					uint32_t sel = (uint32_t)dr[base] % (localk - 1);
					int32_t nb = nn[(size_t)i * localk + sel];
					if (nb < 0) { if (tx == 0 || c == 0) atomicExch(err, 1); nb = (int32_t)i; }
					double alpha = __ddiv_rn((double)dr[base + 1 + c], 2147483647.0);
					double tv = x[(size_t)posIdx[nb] * ld + c];
					double qv = x[(size_t)posIdx[i] * ld + c];
					v = __dadd_rn(__dmul_rn(tv, alpha), __dmul_rn(qv, __dsub_rn(1.0, alpha)));
				}
			} else {
				v = x[(size_t)negIdx[pd.negOffset + (r - pd.nPosOut)] * ld + c];
			}
		}
		tile[rr][tx] = v;
	}
	__syncthreads();
	double* D = Dall + pd.dOffset;
	for (int cc = ty; cc < 32; cc += 8) {
		uint32_t c = c0 + cc, r = r0 + tx;
		if (c < ld && r < pd.Ntr) D[(size_t)c * pd.Ntr + r] = tile[tx][cc];
	}
}

void launch_assemble(const double* x, uint32_t m, const uint32_t* posIdx, uint32_t P, const uint32_t* negIdx, const int32_t* nn,
                     uint32_t localk, uint32_t fp, const int32_t* draws, const PartDesc* parts, uint32_t nPartsBatch,
                     uint32_t maxNtr, double* Dall, int* err, cudaStream_t s) {
	dim3 grid((maxNtr + 31) / 32, (m + 1 + 31) / 32, nPartsBatch);
	assemble_kernel<<<grid, 256, 0, s>>>(x, m, posIdx, P, negIdx, nn, localk, fp, draws, parts, Dall, err);
}

}  // namespace psm
