// Device replay of std::random_shuffle(first, last) as libstdc++ runs it on glibc rand() -- the shuffle of a fold's training
// negatives (reference src/testtraindivider.cpp:90; 12.6M elements per fold at BASELINE configs[2], a sequential, cache-missing
// walk of ~150 ms per fold on the host and the start-up tail of every rank of a multi-GPU run).
//
// The algorithm is  for i = 1 .. n-1:  j_i = rand() % (i + 1);  swap(a[i], a[j_i]).  The draws do not depend on the data and
// the generator can be entered anywhere (glibc_rand.cu), so all j_i are produced in parallel; what is sequential is the chain
// of swaps.  It is undone per output slot by walking the steps BACKWARDS: the content of slot s after step i came from slot j_i
// if s == i, from slot i if s == j_i (i > s), and from s otherwise.  Going down from the last step,
//   * the LARGEST step i > s (not yet passed) with j_i == s, if there is one, ends the walk: the content came from slot i, and
//     no earlier step can touch a slot above its own index -- the source is the original a[i];
//   * otherwise step s itself is next: continue from slot j_s with the steps below s (j_s == s: the source is a[s]).
// "The largest step i <= t with j_i == s" is one binary search in the (j, i) pairs sorted by (j, i).  The walk takes O(log n)
// hops in expectation; every slot is independent, so the whole shuffle is: draws, one radix sort of n pairs, one gather.
#include "kernels.cuh"

namespace psm {

// step i = t + 1 uses draw t (libstdc++ random_shuffle: for (i = first + 1; i != last; ++i) iter_swap(i, first + rand() % ((i - first) + 1)))
__global__ void __launch_bounds__(256) shuffle_keys_kernel(const int32_t* __restrict__ draws, uint32_t n, int shift, unsigned long long* __restrict__ keys,
                                                            uint32_t* __restrict__ vals, uint32_t* __restrict__ jArr) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t == 0) jArr[0] = 0;
	if (t + 1 >= n) return;
	const uint32_t i = t + 1;
	const uint32_t j = (uint32_t)draws[t] % (i + 1u);          // rand() < 2^31 and i + 1 <= 2^31: the 32-bit modulus is the reference's value
	keys[t] = ((unsigned long long)j << shift) | i;
	vals[t] = i;
	jArr[i] = j;
}

__global__ void __launch_bounds__(256) shuffle_resolve_kernel(const unsigned long long* __restrict__ keys, uint32_t nk, const uint32_t* __restrict__ jArr,
                                                               const uint32_t* __restrict__ in, uint32_t n, int shift, uint32_t* __restrict__ out) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const unsigned long long mask = (1ull << shift) - 1ull;
	uint32_t s = k, t = n - 1;                                  // slot being traced; steps <= t are still to be undone
	while (true) {
		// last pair <= (s, t) in (j, i) order
		const unsigned long long target = ((unsigned long long)s << shift) | t;
		uint32_t lo = 0, hi = nk;
		while (lo < hi) {
			const uint32_t mid = lo + ((hi - lo) >> 1);
			if (__ldg(keys + mid) <= target) lo = mid + 1; else hi = mid;
		}
		if (lo > 0) {
			const unsigned long long kk = __ldg(keys + lo - 1);
			const uint32_t jj = (uint32_t)(kk >> shift), ii = (uint32_t)(kk & mask);
			if (jj == s && ii > s) { s = ii; break; }           // hit from above: the content is the original a[ii]
		}
		if (s == 0) break;                                      // slot 0 has no step of its own
		const uint32_t js = __ldg(jArr + s);
		if (js == s) break;                                     // step s swapped the slot with itself; nothing below can reach it
		t = s - 1;
		s = js;
	}
	out[k] = in[s];
}

int shuffle_sort_passes(uint32_t n) {
	const int shift = n <= (1u << 24) ? 24 : 32;
	int jbits = 1;
	while (jbits < 32 && (1ull << jbits) < n) jbits++;
	int passes = (jbits + 7) / 8;
	passes += passes & 1;
	if (shift + 8 * passes > 64) passes = (64 - shift) / 8;
	return passes;
}

// keysA/keysB [n] u64, valsA/valsB [n] u32, hist as launch_radix_sort_u64 needs, draws [n] i32 (already filled with the n - 1 draws
// of the shuffle), jArr [n] u32.  out may not alias in.
int launch_device_shuffle(const int32_t* draws, const uint32_t* in, uint32_t n, unsigned long long* keysA, unsigned long long* keysB, uint32_t* valsA,
                          uint32_t* valsB, uint32_t* hist, uint32_t* jArr, uint32_t* out, cudaStream_t s) {
	if (n == 0) return 0;
	if (n == 1) return cudaMemcpyAsync(out, in, 4, cudaMemcpyDeviceToDevice, s) == cudaSuccess ? 0 : -1;
	const int shift = n <= (1u << 24) ? 24 : 32;
	shuffle_keys_kernel<<<(n + 255) / 256, 256, 0, s>>>(draws, n, shift, keysA, valsA, jArr);
	// The pairs are produced in ascending i and the radix sort is stable, so sorting on the bits of j alone -- the key bits from
	// `shift` up, as many 8-bit passes as n has bits, rounded up to an even count (the result lands back in keysA) -- leaves them in
	// (j, i) order: 4 passes for the 800 k / 12.6 M training negatives of a BASELINE configs[1] / [2] fold instead of 6.
	const int passes = shuffle_sort_passes(n);
	if (launch_radix_sort_u64(keysA, keysB, valsA, valsB, (uint64_t)n - 1, passes, hist, s, shift)) return -1;
	shuffle_resolve_kernel<<<(n + 255) / 256, 256, 0, s>>>(keysA, n - 1, jArr, in, n, shift, out);
	return 0;
}

}  // namespace psm
