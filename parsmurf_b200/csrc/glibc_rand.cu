// Device replay of glibc rand() -- the stream SMOTE draws from (reference src/samplerKNN.cpp:95,99; the generator is
// glibc stdlib/random_r.c TYPE_3: r[i] = r[i-3] + r[i-31] (mod 2^32), output r[i] >> 1).
//
// The recurrence is linear over Z/2^32, so the stream can be entered anywhere: x^n mod (x^31 - x^28 - 1) gives the
// coefficients c_j with r[t+n] = sum_j c_j r[t+j].  Each thread jumps to the start of its own GR_CHUNK-draw chunk by
// multiplying precomputed powers x^(GR_CHUNK*2^j), rebuilds the 31-word window there and then steps the plain recurrence.
// This removes the host-side generation of P*fp*(m+1) draws per partition (32M draws per CV at BASELINE configs[1]).
#include "kernels.cuh"

#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace psm {

constexpr int GR_DEG = 31;
constexpr uint32_t GR_ROUNDS = 8;
constexpr uint32_t GR_CHUNK = GR_ROUNDS * 31;   // draws per thread: whole turns of the 31-word window, so the window stays in registers

// ---- polynomial arithmetic mod x^31 - x^28 - 1 over Z/2^32 (host and device) ----
__host__ __device__ inline void gr_polymul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
	uint32_t d[2 * GR_DEG - 1];
	for (int k = 0; k < 2 * GR_DEG - 1; k++) d[k] = 0;
	for (int i = 0; i < GR_DEG; i++) {
		const uint32_t ai = a[i];
		if (ai == 0) continue;
		for (int j = 0; j < GR_DEG; j++) d[i + j] += ai * b[j];
	}
	for (int k = 2 * GR_DEG - 2; k >= GR_DEG; k--) {   // x^k = x^(k-3) + x^(k-31)
		d[k - 3] += d[k];
		d[k - GR_DEG] += d[k];
	}
	for (int k = 0; k < GR_DEG; k++) out[k] = d[k];
}

// x^n mod charpoly
static void gr_xpow(uint64_t n, uint32_t* out) {
	uint32_t result[GR_DEG] = {0}, base[GR_DEG] = {0}, tmp[GR_DEG];
	result[0] = 1;
	base[1] = 1;
	while (n) {
		if (n & 1) { gr_polymul(result, base, tmp); memcpy(result, tmp, sizeof(tmp)); }
		n >>= 1;
		if (n) { gr_polymul(base, base, tmp); memcpy(base, tmp, sizeof(tmp)); }
	}
	memcpy(out, result, sizeof(result));
}

// extend a chronological window w[0..30] (w[30] newest) to 61 values by stepping the recurrence
static void gr_extend(const uint32_t* w31, uint32_t* w61) {
	memcpy(w61, w31, GR_DEG * 4);
	for (int i = GR_DEG; i < 2 * GR_DEG - 1; i++) w61[i] = w61[i - 3] + w61[i - GR_DEG];
}

// advance a chronological window by n draws (host helper, exported through the C ABI as psm_glibc_jump)
void glibc_window_jump(uint32_t* w31, uint64_t n) {
	if (n == 0) return;
	uint32_t c[GR_DEG], w61[2 * GR_DEG - 1], nw[GR_DEG];
	gr_xpow(n, c);
	gr_extend(w31, w61);
	for (int d = 0; d < GR_DEG; d++) {
		uint32_t s = 0;
		for (int j = 0; j < GR_DEG; j++) s += c[j] * w61[d + j];
		nw[d] = s;
	}
	memcpy(w31, nw, sizeof(nw));
}

// Jump tables, radix 256: T[d][v] = x^(GR_CHUNK * v * 256^d) mod charpoly, d < GR_LEVELS.  They depend on nothing but GR_CHUNK,
// so they are built once per device (host, ~1.5 ms) and every thread reaches its chunk with at most GR_LEVELS - 1 polynomial
// products -- none below 256 chunks, one up to 16 M draws.  (r02 before: x^(GR_CHUNK * 2^j) powers recomputed by the host for every
// launch and one product per set bit of the chunk index, ~10 per warp at a fold's 800 k shuffle draws: 138 us per launch on the
// two dozen CTAs such a launch has, 1.4 ms per BASELINE configs[1] CV and on the critical path of every fold's set-up.)
constexpr int GR_LEVELS = 3;
constexpr unsigned long long GR_MAX_DRAWS = (unsigned long long)GR_CHUNK << 24;

__global__ void __launch_bounds__(128) glibc_rand_fill_kernel(const uint32_t* __restrict__ w61, const uint32_t* __restrict__ tab /*[GR_LEVELS][256][31]*/,
                                                               unsigned long long count, int32_t* __restrict__ out) {
	const unsigned long long chunk = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long first = chunk * GR_CHUNK;
	if (first >= count) return;
	uint32_t q[GR_DEG], t[GR_DEG], pj[GR_DEG];
	{
		const uint32_t* t0 = tab + (size_t)(chunk & 255ull) * GR_DEG;
		for (int k = 0; k < GR_DEG; k++) q[k] = t0[k];
	}
	for (int d = 1; d < GR_LEVELS; d++) {                  // (a warp's 32 chunks share the higher digits: uniform branches)
		const uint32_t v = (uint32_t)(chunk >> (8 * d)) & 255u;
		if (v == 0) continue;
		const uint32_t* td = tab + ((size_t)d * 256 + v) * GR_DEG;
		for (int k = 0; k < GR_DEG; k++) pj[k] = td[k];
		gr_polymul(q, pj, t);
		for (int k = 0; k < GR_DEG; k++) q[k] = t[k];
	}
	uint32_t w[GR_DEG];                                    // window at the chunk start: r[first-31 .. first-1] relative to the stream
#pragma unroll
	for (int d = 0; d < GR_DEG; d++) w[d] = 0;
#pragma unroll 1
	for (int j = 0; j < GR_DEG; j++) {                     // (j outside, d unrolled: w[] is only ever indexed by constants)
		const uint32_t qj = q[j];
#pragma unroll
		for (int d = 0; d < GR_DEG; d++) w[d] += qj * w61[d + j];
	}
	// 248 draws = 8 turns of the window with compile-time indices (w[k] += w[k + 28 mod 31]): the window lives in registers.
	// (r02: 1024-draw chunks walked a window in local memory -- a dependent load / store chain of ~100 cycles per draw on the
	// 782 threads of a fold's shuffle draws: 190 us per launch, 1.9 ms per BASELINE configs[1] CV.)
	const unsigned long long n = (count - first < GR_CHUNK) ? (count - first) : GR_CHUNK;
	int32_t* o = out + first;
#pragma unroll 1
	for (uint32_t turn = 0; turn < GR_ROUNDS; turn++) {
		const uint32_t base = turn * GR_DEG;
		if (base >= n) break;
#pragma unroll
		for (int k = 0; k < GR_DEG; k++) {
			const uint32_t v = w[k] + w[(k + GR_DEG - 3) % GR_DEG];
			w[k] = v;
			if (base + k < n) o[base + k] = (int32_t)(v >> 1);
		}
	}
}

// the jump tables of a device: built on first use, shared by every context (fold lane) of the device, never freed
static const uint32_t* glibc_jump_tables(int device) {
	static std::mutex mu;
	static std::map<int, uint32_t*> tabs;
	std::lock_guard<std::mutex> lk(mu);
	auto it = tabs.find(device);
	if (it != tabs.end()) return it->second;
	std::vector<uint32_t> h((size_t)GR_LEVELS * 256 * GR_DEG, 0u);
	uint32_t step[GR_DEG], tmp[GR_DEG];
	uint64_t e = GR_CHUNK;
	for (int d = 0; d < GR_LEVELS; d++, e *= 256) {
		gr_xpow(e, step);                                          // x^(GR_CHUNK * 256^d)
		uint32_t* T = &h[(size_t)d * 256 * GR_DEG];
		T[0] = 1;                                                  // v = 0: the polynomial 1
		for (int v = 1; v < 256; v++) {
			gr_polymul(T + (size_t)(v - 1) * GR_DEG, step, tmp);
			memcpy(T + (size_t)v * GR_DEG, tmp, sizeof(tmp));
		}
	}
	uint32_t* dev = nullptr;
	if (cudaMalloc(&dev, h.size() * 4) != cudaSuccess) return nullptr;
	if (cudaMemcpy(dev, h.data(), h.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return nullptr; }
	tabs[device] = dev;
	return dev;
}

// host side of the launch: fills `out` (device) with the next `count` rand() outputs of the stream whose chronological
// window is w31; scratch = device buffer of at least 61 uint32
int launch_glibc_rand_fill(const uint32_t* w31, unsigned long long count, uint32_t* scratchDev, int32_t* out, cudaStream_t s) {
	if (count == 0) return 0;
	if (count > GR_MAX_DRAWS) return -1;
	int device = 0;
	if (cudaGetDevice(&device) != cudaSuccess) return -1;
	const uint32_t* tab = glibc_jump_tables(device);
	if (!tab) return -1;
	const unsigned long long chunks = (count + GR_CHUNK - 1) / GR_CHUNK;
	uint32_t h[61];
	gr_extend(w31, h);
	if (cudaMemcpyAsync(scratchDev, h, sizeof(h), cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
	if (cudaStreamSynchronize(s) != cudaSuccess) return -1;    // `h` is a local
	glibc_rand_fill_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, s>>>(scratchDev, tab, count, out);
	return 0;
}

}  // namespace psm
