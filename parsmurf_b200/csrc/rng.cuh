// Device replay of the random streams ranger consumes (reference src/ranger/Tree/Tree.cpp:57,413-432,
// 245-258; src/ranger/Utility/utility.cpp:84-139), so that a tree grown on the GPU makes the SAME bootstrap
// and mtry draws as the reference tree with the same seed (the stream is data-independent).
//
//  * std::mt19937_64                      -- published MT19937-64 (Matsumoto & Nishimura); one warp regenerates
//                                            the 312-word state block cooperatively (3 dependency phases).
//  * std::uniform_int_distribution<size_t>(0, N-1) on a 64-bit engine (libstdc++ >= 11, bits/uniform_int_dist.h
//    _S_nd): Lemire's nearly-divisionless method: hi64(r*N), redraw while lo64(r*N) < (2^64 mod N).
//  * std::uniform_real_distribution<double>(0,1) -> generate_canonical<double,53> (bits/random.tcc:3349-3381)
//    with a 64-bit engine: one draw, (double)r (round-to-nearest) * 2^-64, clamped to nextafter(1,0).
#pragma once
#include "common.cuh"

namespace psm {

constexpr int MT_N = 312;
constexpr int MT_M = 156;

__device__ __forceinline__ unsigned long long mt_temper(unsigned long long y) {
	y ^= (y >> 29) & 0x5555555555555555ull;
	y ^= (y << 17) & 0x71D67FFFEDA60000ull;
	y ^= (y << 37) & 0xFFF7EEE000000000ull;
	y ^= (y >> 43);
	return y;
}

// seed by one lane (sequential recurrence); call with the whole warp, followed by __syncwarp()
__device__ __forceinline__ void mt_seed(unsigned long long* mt, unsigned long long seed) {
	if (lane_id() == 0) {
		unsigned long long x = seed;
		mt[0] = x;
		for (int i = 1; i < MT_N; i++) {
			x = 6364136223846793005ull * (x ^ (x >> 62)) + (unsigned long long)i;
			mt[i] = x;
		}
	}
	__syncwarp();
}

__device__ __forceinline__ unsigned long long mt_mix(unsigned long long a, unsigned long long b, unsigned long long c) {
	unsigned long long x = (a & 0xFFFFFFFF80000000ull) | (b & 0x7FFFFFFFull);
	return c ^ (x >> 1) ^ ((x & 1ull) ? 0xB5026F5AA96619E9ull : 0ull);
}

// regenerate the state block; warp-cooperative.  Word i needs old[i], old[i+1] and word i+156 (old for i<156,
// new for i>=156), so: phase 1 = words 0..155 in 32-wide steps, phase 2 = 156..310, phase 3 = word 311.
__device__ __forceinline__ void mt_twist(unsigned long long* mt) {
	const int lane = lane_id();
	__syncwarp();
	for (int i0 = 0; i0 < MT_N - MT_M; i0 += 32) {
		int i = i0 + lane;
		unsigned long long v = 0;
		if (i < MT_N - MT_M) v = mt_mix(mt[i], mt[i + 1], mt[i + MT_M]);
		__syncwarp();
		if (i < MT_N - MT_M) mt[i] = v;
		__syncwarp();
	}
	for (int i0 = MT_N - MT_M; i0 < MT_N - 1; i0 += 32) {
		int i = i0 + lane;
		unsigned long long v = 0;
		if (i < MT_N - 1) v = mt_mix(mt[i], mt[i + 1], mt[i + MT_M - MT_N]);
		__syncwarp();
		if (i < MT_N - 1) mt[i] = v;
		__syncwarp();
	}
	if (lane == 0) mt[MT_N - 1] = mt_mix(mt[MT_N - 1], mt[0], mt[MT_M - 1]);
	__syncwarp();
}

// Lemire step for range N: returns true if the raw output r is accepted, value in *out
__device__ __forceinline__ bool lemire_accept(unsigned long long r, unsigned long long N, unsigned long long* out) {
	unsigned long long lo = r * N;
	unsigned long long hi = __umul64hi(r, N);
	*out = hi;
	if (lo < N) {
		unsigned long long threshold = (0ull - N) % N;
		if (lo < threshold) return false;
	}
	return true;
}

__device__ __forceinline__ double canonical01(unsigned long long r) {
	double u = __dmul_rn(__ull2double_rn(r), __longlong_as_double(0x3BF0000000000000ll));   // * 2^-64, exact scaling
	if (u >= 1.0) u = __longlong_as_double(0x3FEFFFFFFFFFFFFFll);        // nextafter(1.0, 0.0)
	return u;
}

}  // namespace psm
