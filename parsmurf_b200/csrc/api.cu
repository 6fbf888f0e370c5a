// parsmurf_b200 -- C-ABI implementation (include/parsmurf_b200.h): context, device buffers, batch orchestration.
// No CPU fallback: every compute entry needs a CUDA device and fails with PSM_E_CUDA otherwise.
#include "../../include/parsmurf_b200.h"
#include "kernels.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

using namespace psm;

namespace {

struct DevBuf {
	void* p = nullptr;
	size_t cap = 0;
	cudaError_t ensure(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) { cudaFree(p); p = nullptr; cap = 0; }
		size_t want = bytes + bytes / 8 + 256;     // slack: growing means cudaFree + cudaMalloc, which stall every stream of the device
		cudaError_t e = cudaMalloc(&p, want);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <typename T> T* as() const { return (T*)p; }
};

struct ProfPair { cudaEvent_t a, b; int cls; };

}  // namespace

struct psm_ctx {
	int device = 0;
	cudaStream_t stream = 0;
	cudaStream_t ownStream = nullptr;   // psm_ctx_own_stream
	uint64_t budget = 24ull << 30;
	std::string err;

	// dataset
	DevBuf xOwned;
	const double* x = nullptr;
	uint32_t n = 0, m = 0;

	// fold
	bool foldOpen = false;
	uint32_t P = 0, Nneg = 0, nTP = 0, nTN = 0, Nte = 0, localk = 0;
	bool haveKnn = false;
	std::vector<uint32_t> hTestIdx;
	DevBuf dPos, dNeg, dTest, dXp, dNN, dNNd, dKnnScratch, dKnnTc, dClass12;
	int knnPath = 0;
	uint32_t knnIncomplete = 0;

	// folds on the device (psm_folds_upload): examples of every fold by class; shuffle scratch
	DevBuf dPosList, dNegList, sKeysA, sKeysB, sValsA, sValsB, sHist, sJ, sDraws, sIn;
	DevBuf fbKeysA, fbKeysB, fbValsA, fbValsB, fbHist, fbCounts, dFoldIds;   // psm_folds_build_device
	std::vector<uint32_t> hFoldIds;    // the fold ids resident in dFoldIds
	const uint32_t* fPos = nullptr;    // device lists (own buffers or another context's, psm_folds_attach)
	const uint32_t* fNeg = nullptr;
	std::vector<uint32_t> fPosOff, fNegOff;
	bool foldFromLists = false;        // the open fold was built by psm_fold_begin_device (no host copy of the test ids)

	// batch
	bool batchOpen = false, assembled = false, trained = false, inbagValid = false;
	uint32_t nParts = 0, fp = 0, ratio = 0, p0 = 0, p1 = 0, nB = 0, maxNtr = 0;
	std::vector<PartDesc> hParts;
	uint64_t totalD = 0, totalS = 0, totalDraws = 0;
	DevBuf dParts, dD, dS, dKeysA, dKeysB, dValsA, dValsB, dDraws, dErr, dRandScratch;

	// train
	uint32_t T = 0, mtry = 0;
	TrainArgs A;
	DevBuf dLists0, dLists1, dInbag, dInbagT, dMt, dTs, dNodes, dNodes8, dSplitFeat, dFlagBits, dItems0, dItems1, dCls, dCand, dRes, dDec, dPtask, dCounters, dStats, dSeeds,
	    dInjBoot, dInjBootOff, dInjCand, dSortFlags;
	uint32_t* hCounters = nullptr;   // pinned
	cudaEvent_t evLevel = nullptr;
	cudaStream_t copyStream = nullptr;   // score fetch (scores_fetch)
	cudaEvent_t evCopy = nullptr;
	void* hStage = nullptr;          // pinned D2H staging
	size_t hStageCap = 0;
	uint64_t statSplitBytes = 0, statNodes = 0, statLevels = 0;
	uint32_t maxNodes = 0;

	// predict
	DevBuf dPredScratch;

	// device-resident labels / global scores
	DevBuf dLabels, dFoldLabels, dGlob;
	std::vector<uint32_t> hLabels;
	uint32_t labelsPos = 0, labelsNeg = 0;   // class totals of hLabels
	bool haveLabels = false, foldLabelsReady = false;

	// curves
	uint64_t curvesMixedPairs = 0;
	int curvesUsedHostSort = 0;
	double curvesSpread[2] = {0, 0};
	DevBuf cIdx, cFoldScores, cScratch, cLabels, cPerm, cSorted, cSorted2, cTp, cPartials, cOut, cScores, cKeysA, cKeysB, cValsA, cValsB, cHist;

	// cross-rank communicator (psm_comm_init)
	void* comm = nullptr;            // ncclComm_t
	uint32_t commRank = 0, commWorld = 1;

	// profiling
	bool profOn = false;
	uint32_t profMask = 0xFFFFFFFFu;
	std::vector<ProfPair> profPending;
	std::vector<cudaEvent_t> evPool;
	double profMs[PSM_NUM_KERNEL_CLASSES] = {0};
	uint64_t profLaunches[PSM_NUM_KERNEL_CLASSES] = {0};
};

namespace {

int fail(psm_ctx* c, int code, const std::string& msg) {
	if (c) c->err = msg;
	return code;
}
int cuda_fail(psm_ctx* c, cudaError_t e, const char* where) {
	return fail(c, PSM_E_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                        \
	do {                                                                \
		cudaError_t _e = (call);                                        \
		if (_e != cudaSuccess) return cuda_fail(ctx, _e, #call);        \
	} while (0)
#define ENSURE(buf, bytes) CU((buf).ensure(bytes))

cudaEvent_t get_event(psm_ctx* c) {
	if (!c->evPool.empty()) { cudaEvent_t e = c->evPool.back(); c->evPool.pop_back(); return e; }
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}

struct ProfScope {
	psm_ctx* c; int cls; cudaEvent_t a = nullptr;
	ProfScope(psm_ctx* c_, int cls_, int launches) : c(c_), cls(cls_) {
		c->profLaunches[cls] += launches;
		if (c->profOn && ((c->profMask >> cls) & 1u)) { a = get_event(c); cudaEventRecord(a, c->stream); }
	}
	void add_launches(int n) { c->profLaunches[cls] += n; }
	~ProfScope() {
		if (a) { cudaEvent_t b = get_event(c); cudaEventRecord(b, c->stream); c->profPending.push_back({a, b, cls}); }
	}
};

void prof_resolve(psm_ctx* c) {
	if (c->profPending.empty()) return;
	cudaStreamSynchronize(c->stream);
	for (auto& p : c->profPending) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) c->profMs[p.cls] += ms;
		c->evPool.push_back(p.a); c->evPool.push_back(p.b);
	}
	c->profPending.clear();
}

enum { K_KNN = 0, K_ASSEMBLE, K_SORT, K_BOOT, K_LISTS, K_SEARCH, K_SELECT, K_FINALIZE, K_PARTITION, K_PREDICT, K_CURVES, K_GATHER };

int partition_sizes(uint32_t P, uint32_t Nneg, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t p, PartDesc* d) {
	// Sampler::setPartition (src/sampler.cpp:62-71) + SamplerKNN::setup (src/samplerKNN.cpp:15-30)
	uint32_t chunk = (uint32_t)std::ceil(Nneg / (double)nParts);
	uint64_t before = (uint64_t)chunk * (nParts - 1);
	if (p == nParts - 1 && before > Nneg) return PSM_E_CHUNK;
	uint32_t negAvail = (p != nParts - 1) ? chunk : (uint32_t)(Nneg - before);
	uint32_t nPosOut = (fp + 1) * P;
	uint32_t nNegOut = nPosOut * ratio;
	if (nNegOut == 0) nNegOut = negAvail;
	if (nNegOut > negAvail) nNegOut = negAvail;
	d->negOffset = p * chunk;
	d->nPosOut = nPosOut;
	d->nNegOut = nNegOut;
	d->Ntr = nPosOut + nNegOut;
	return PSM_OK;
}

}  // namespace

extern "C" {

const char* psm_version(void) { return "parsmurf_b200 0.1 (sm_100a)"; }

int psm_ctx_create(int device, psm_ctx** out) {
	if (!out) return PSM_E_ARG;
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || device < 0 || device >= count) return PSM_E_CUDA;
	if (cudaSetDevice(device) != cudaSuccess) return PSM_E_CUDA;
	psm_ctx* c = new psm_ctx;
	c->device = device;
	if (cudaMallocHost((void**)&c->hCounters, CNT_COUNT * sizeof(uint32_t)) != cudaSuccess) { delete c; return PSM_E_CUDA; }
	memset(&c->A, 0, sizeof(c->A));
	*out = c;
	return PSM_OK;
}

int psm_ctx_destroy(psm_ctx* ctx) {
	if (!ctx) return PSM_E_ARG;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	DevBuf* bufs[] = {&ctx->fbKeysA, &ctx->fbKeysB, &ctx->fbValsA, &ctx->fbValsB, &ctx->fbHist, &ctx->fbCounts, &ctx->dFoldIds, &ctx->dPosList, &ctx->dNegList, &ctx->sKeysA, &ctx->sKeysB, &ctx->sValsA, &ctx->sValsB, &ctx->sHist, &ctx->sJ, &ctx->sDraws, &ctx->sIn, &ctx->xOwned, &ctx->dPos, &ctx->dNeg, &ctx->dTest, &ctx->dXp, &ctx->dNN, &ctx->dNNd, &ctx->dKnnScratch, &ctx->dKnnTc, &ctx->dClass12,
	                  &ctx->dParts, &ctx->dD, &ctx->dS, &ctx->dKeysA, &ctx->dKeysB, &ctx->dValsA, &ctx->dValsB, &ctx->dDraws, &ctx->dErr, &ctx->dRandScratch,
	                  &ctx->dLists0, &ctx->dLists1, &ctx->dInbag, &ctx->dInbagT, &ctx->dMt, &ctx->dTs, &ctx->dNodes, &ctx->dNodes8, &ctx->dSplitFeat, &ctx->dFlagBits, &ctx->dItems0, &ctx->dItems1, &ctx->dCls,
	                  &ctx->dCand, &ctx->dRes, &ctx->dDec, &ctx->dPtask, &ctx->dCounters, &ctx->dStats, &ctx->dSeeds, &ctx->dInjBoot, &ctx->dInjBootOff,
	                  &ctx->dInjCand, &ctx->dSortFlags, &ctx->dPredScratch, &ctx->dLabels, &ctx->dFoldLabels, &ctx->dGlob, &ctx->cIdx, &ctx->cFoldScores, &ctx->cScratch, &ctx->cLabels, &ctx->cPerm, &ctx->cSorted, &ctx->cSorted2, &ctx->cTp, &ctx->cPartials, &ctx->cOut,
	                  &ctx->cScores, &ctx->cKeysA, &ctx->cKeysB, &ctx->cValsA, &ctx->cValsB, &ctx->cHist};
	for (DevBuf* b : bufs) b->release();
	for (auto& p : ctx->profPending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
	for (auto e : ctx->evPool) cudaEventDestroy(e);
	if (ctx->hCounters) cudaFreeHost(ctx->hCounters);
	if (ctx->evLevel) cudaEventDestroy(ctx->evLevel);
	if (ctx->hStage) cudaFreeHost(ctx->hStage);
	if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
	if (ctx->evCopy) cudaEventDestroy(ctx->evCopy);
	if (ctx->comm) psm_comm_destroy(ctx);
	if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
	delete ctx;
	return PSM_OK;
}

const char* psm_last_error(psm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int psm_set_stream(psm_ctx* ctx, void* s) {
	if (!ctx) return PSM_E_ARG;
	ctx->stream = (cudaStream_t)s;
	return PSM_OK;
}
int psm_ctx_own_stream(psm_ctx* ctx) {
	if (!ctx) return PSM_E_ARG;
	CU(cudaSetDevice(ctx->device));
	if (!ctx->ownStream) CU(cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking));
	ctx->stream = ctx->ownStream;
	return PSM_OK;
}
int psm_ctx_bind_thread(psm_ctx* ctx) {
	if (!ctx) return PSM_E_ARG;
	CU(cudaSetDevice(ctx->device));
	return PSM_OK;
}
int psm_dataset_devptr(psm_ctx* ctx, const double** x_dev) {
	if (!ctx || !x_dev || !ctx->x) return fail(ctx, PSM_E_ARG, "dataset_devptr: no dataset");
	*x_dev = ctx->x;
	return PSM_OK;
}
int psm_set_batch_budget(psm_ctx* ctx, uint64_t bytes) {
	if (!ctx) return PSM_E_ARG;
	ctx->budget = bytes ? bytes : (24ull << 30);
	return PSM_OK;
}
int psm_sync(psm_ctx* ctx) {
	if (!ctx) return PSM_E_ARG;
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}

// ------------------------------------------------------------------------------------------ dataset
int psm_dataset_upload(psm_ctx* ctx, const double* x_host, uint32_t n, uint32_t m) {
	if (!ctx || !x_host || n == 0 || m == 0) return fail(ctx, PSM_E_ARG, "dataset_upload: bad arguments");
	CU(cudaSetDevice(ctx->device));
	size_t bytes = (size_t)n * (m + 1) * sizeof(double);
	ENSURE(ctx->xOwned, bytes);
	CU(cudaMemcpyAsync(ctx->xOwned.p, x_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));   // x_host is only borrowed for the call
	ctx->x = ctx->xOwned.as<double>();
	ctx->n = n; ctx->m = m;
	ctx->foldOpen = ctx->batchOpen = false;
	ctx->haveLabels = false;   // a new dataset copy: its labels and fold ids are uploaded again too (kept resident otherwise)
	ctx->hFoldIds.clear();
	return PSM_OK;
}

int psm_dataset_attach_device(psm_ctx* ctx, const double* x_dev, uint32_t n, uint32_t m) {
	if (!ctx || !x_dev || n == 0 || m == 0) return fail(ctx, PSM_E_ARG, "dataset_attach_device: bad arguments");
	ctx->x = x_dev; ctx->n = n; ctx->m = m;
	ctx->foldOpen = ctx->batchOpen = false;
	return PSM_OK;
}

// ------------------------------------------------------------------------------------------ fold
int psm_fold_begin(psm_ctx* ctx, const uint32_t* trngPosIdx, uint32_t P, const uint32_t* trngNegIdx, uint32_t Nneg,
                   const uint32_t* testPosIdx, uint32_t nTP, const uint32_t* testNegIdx, uint32_t nTN) {
	if (!ctx || !ctx->x) return fail(ctx, PSM_E_ARG, "fold_begin: no dataset");
	if ((P && !trngPosIdx) || (Nneg && !trngNegIdx) || (nTP && !testPosIdx) || (nTN && !testNegIdx)) return fail(ctx, PSM_E_ARG, "fold_begin: null index array");
	CU(cudaSetDevice(ctx->device));
	ctx->P = P; ctx->Nneg = Nneg; ctx->nTP = nTP; ctx->nTN = nTN; ctx->Nte = nTP + nTN;
	ctx->haveKnn = false; ctx->batchOpen = false; ctx->localk = 0; ctx->foldLabelsReady = false;
	ENSURE(ctx->dPos, std::max<size_t>(4, (size_t)P * 4));
	ENSURE(ctx->dNeg, std::max<size_t>(4, (size_t)Nneg * 4));
	ENSURE(ctx->dTest, std::max<size_t>(4, (size_t)ctx->Nte * 4));
	ENSURE(ctx->dClass12, std::max<size_t>(16, (size_t)ctx->Nte * 16));
	ctx->hTestIdx.resize(ctx->Nte);
	if (nTP) memcpy(ctx->hTestIdx.data(), testPosIdx, (size_t)nTP * 4);       // positives first, then negatives (sampler.cpp:90-95)
	if (nTN) memcpy(ctx->hTestIdx.data() + nTP, testNegIdx, (size_t)nTN * 4);
	// synchronous copies: the host arrays are only borrowed for the duration of the call
	if (P) CU(cudaMemcpyAsync(ctx->dPos.p, trngPosIdx, (size_t)P * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (Nneg) CU(cudaMemcpyAsync(ctx->dNeg.p, trngNegIdx, (size_t)Nneg * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (ctx->Nte) CU(cudaMemcpyAsync(ctx->dTest.p, ctx->hTestIdx.data(), (size_t)ctx->Nte * 4, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemsetAsync(ctx->dClass12.p, 0, std::max<size_t>(16, (size_t)ctx->Nte * 16), ctx->stream));
	// every id of all four arrays must name an example: they index x in the assemble / predict kernels and the score arrays
	ENSURE(ctx->dErr, 16);
	uint32_t bad = 0;
	CU(cudaMemsetAsync(ctx->dErr.as<uint32_t>() + 3, 0, 4, ctx->stream));
	launch_check_ids(ctx->dPos.as<uint32_t>(), P, ctx->n, ctx->dErr.as<uint32_t>() + 3, ctx->stream);
	launch_check_ids(ctx->dNeg.as<uint32_t>(), Nneg, ctx->n, ctx->dErr.as<uint32_t>() + 3, ctx->stream);
	launch_check_ids(ctx->dTest.as<uint32_t>(), ctx->Nte, ctx->n, ctx->dErr.as<uint32_t>() + 3, ctx->stream);
	CU(cudaMemcpyAsync(&bad, ctx->dErr.as<uint32_t>() + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	if (bad) { ctx->foldOpen = false; return fail(ctx, PSM_E_ARG, "fold_begin: index out of range"); }
	ctx->foldOpen = true; ctx->foldFromLists = false;
	return PSM_OK;
}

int psm_folds_upload(psm_ctx* ctx, uint32_t nFolds, const uint32_t* posList, const uint32_t* posOff, const uint32_t* negList, const uint32_t* negOff) {
	if (!ctx || !ctx->x || nFolds == 0 || !posOff || !negOff) return fail(ctx, PSM_E_ARG, "folds_upload: bad arguments");
	CU(cudaSetDevice(ctx->device));
	const uint32_t totPos = posOff[nFolds], totNeg = negOff[nFolds];
	if ((totPos && !posList) || (totNeg && !negList) || (uint64_t)totPos + totNeg > ctx->n) return fail(ctx, PSM_E_ARG, "folds_upload: lists do not match the dataset");
	for (uint32_t f = 0; f < nFolds; f++) if (posOff[f] > posOff[f + 1] || negOff[f] > negOff[f + 1]) return fail(ctx, PSM_E_ARG, "folds_upload: offsets not ascending");
	ENSURE(ctx->dPosList, std::max<size_t>(4, (size_t)totPos * 4));
	ENSURE(ctx->dNegList, std::max<size_t>(4, (size_t)totNeg * 4));
	if (totPos) CU(cudaMemcpyAsync(ctx->dPosList.p, posList, (size_t)totPos * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (totNeg) CU(cudaMemcpyAsync(ctx->dNegList.p, negList, (size_t)totNeg * 4, cudaMemcpyHostToDevice, ctx->stream));
	ENSURE(ctx->dErr, 16);
	uint32_t bad = 0;
	CU(cudaMemsetAsync(ctx->dErr.as<uint32_t>() + 3, 0, 4, ctx->stream));
	launch_check_ids(ctx->dPosList.as<uint32_t>(), totPos, ctx->n, ctx->dErr.as<uint32_t>() + 3, ctx->stream);
	launch_check_ids(ctx->dNegList.as<uint32_t>(), totNeg, ctx->n, ctx->dErr.as<uint32_t>() + 3, ctx->stream);
	CU(cudaMemcpyAsync(&bad, ctx->dErr.as<uint32_t>() + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));   // the host lists are only borrowed
	if (bad) { ctx->fPos = ctx->fNeg = nullptr; return fail(ctx, PSM_E_ARG, "folds_upload: example id out of range"); }
	ctx->fPos = ctx->dPosList.as<uint32_t>(); ctx->fNeg = ctx->dNegList.as<uint32_t>();
	ctx->fPosOff.assign(posOff, posOff + nFolds + 1); ctx->fNegOff.assign(negOff, negOff + nFolds + 1);
	ctx->foldOpen = false;
	return PSM_OK;
}

// equality of two u32 arrays, a few host threads on large inputs (14M ids at BASELINE configs[2]: one thread needs ~5 ms)
static bool same_u32(const uint32_t* a, const uint32_t* b, size_t n) {
	if (n < (1u << 20)) return memcmp(a, b, n * 4) == 0;
	const unsigned nT = 4;
	std::atomic<bool> same(true);
	std::vector<std::thread> th;
	for (unsigned t = 0; t < nT; t++)
		th.emplace_back([&, t]() { const size_t lo = n * t / nT, hi = n * (t + 1) / nT; if (memcmp(a + lo, b + lo, (hi - lo) * 4) != 0) same = false; });
	for (auto& x : th) x.join();
	return same;
}

int psm_folds_build_device(psm_ctx* ctx, const uint32_t* fold_ids, uint32_t nFolds, uint32_t* nPos, uint32_t* nNeg) {
	if (!ctx || !ctx->x || !ctx->haveLabels || !fold_ids || nFolds == 0 || !nPos || !nNeg) return fail(ctx, PSM_E_ARG, "folds_build_device: need the dataset, the labels and the fold ids");
	if (2ull * nFolds > 2048) return fail(ctx, PSM_E_LIMIT, "folds_build_device: more than 1024 folds");
	CU(cudaSetDevice(ctx->device));
	const uint32_t n = ctx->n;
	cudaStream_t s = ctx->stream;
	if (!(ctx->hFoldIds.size() == n && same_u32(ctx->hFoldIds.data(), fold_ids, n))) {
		ctx->hFoldIds.assign(fold_ids, fold_ids + n);
		ENSURE(ctx->dFoldIds, (size_t)n * 4);
		CU(cudaMemcpyAsync(ctx->dFoldIds.p, ctx->hFoldIds.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
	}
	const uint64_t rBlocks = ((uint64_t)n + 4095) / 4096;
	ENSURE(ctx->fbKeysA, (size_t)n * 8); ENSURE(ctx->fbKeysB, (size_t)n * 8);
	ENSURE(ctx->fbValsA, (size_t)n * 4); ENSURE(ctx->fbValsB, (size_t)n * 4);
	ENSURE(ctx->fbHist, (rBlocks + 1) * 256 * 4);
	ENSURE(ctx->fbCounts, (size_t)(2 * nFolds + 1) * 4);
	uint32_t* counts = ctx->fbCounts.as<uint32_t>();
	CU(cudaMemsetAsync(counts, 0, (size_t)(2 * nFolds + 1) * 4, s));
	{
		ProfScope ps(ctx, K_GATHER, 1 + 3 * 2);
		launch_fold_keys(ctx->dLabels.as<uint32_t>(), ctx->dFoldIds.as<uint32_t>(), n, nFolds, ctx->fbKeysA.as<unsigned long long>(), ctx->fbValsA.as<uint32_t>(), counts,
		                 counts + 2 * nFolds, s);
		if (launch_radix_sort_u64(ctx->fbKeysA.as<unsigned long long>(), ctx->fbKeysB.as<unsigned long long>(), ctx->fbValsA.as<uint32_t>(), ctx->fbValsB.as<uint32_t>(), n, 2,
		                          ctx->fbHist.as<uint32_t>(), s))
			return fail(ctx, PSM_E_CUDA, "folds_build_device: sort launch failed");
	}
	std::vector<uint32_t> hc(2 * nFolds + 1);
	CU(cudaMemcpyAsync(hc.data(), counts, hc.size() * 4, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	if (hc[2 * nFolds]) { ctx->fPos = ctx->fNeg = nullptr; return fail(ctx, PSM_E_ARG, "folds_build_device: fold id out of range"); }
	ctx->fPosOff.assign(nFolds + 1, 0); ctx->fNegOff.assign(nFolds + 1, 0);
	for (uint32_t f = 0; f < nFolds; f++) {
		nPos[f] = hc[f]; nNeg[f] = hc[nFolds + f];
		ctx->fPosOff[f + 1] = ctx->fPosOff[f] + nPos[f];
		ctx->fNegOff[f + 1] = ctx->fNegOff[f] + nNeg[f];
	}
	// sorted by (class, fold), example ids ascending inside every key: the positives' lists, then the negatives'
	ctx->fPos = ctx->fbValsA.as<uint32_t>();
	ctx->fNeg = ctx->fbValsA.as<uint32_t>() + ctx->fPosOff[nFolds];
	ctx->foldOpen = false;
	return PSM_OK;
}

int psm_folds_download(psm_ctx* ctx, uint32_t* posList, uint32_t* negList) {
	if (!ctx || !ctx->fPos || ctx->fPosOff.empty()) return fail(ctx, PSM_E_ARG, "folds_download: no device fold lists");
	CU(cudaSetDevice(ctx->device));
	const size_t tp = ctx->fPosOff.back(), tn = ctx->fNegOff.back();
	if (posList && tp) CU(cudaMemcpyAsync(posList, ctx->fPos, tp * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (negList && tn) CU(cudaMemcpyAsync(negList, ctx->fNeg, tn * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}

int psm_folds_attach(psm_ctx* ctx, psm_ctx* src) {
	if (!ctx || !src || !src->fPos || ctx->device != src->device || ctx->n != src->n) return fail(ctx, PSM_E_ARG, "folds_attach: source has no fold lists for this dataset");
	ctx->fPos = src->fPos; ctx->fNeg = src->fNeg; ctx->fPosOff = src->fPosOff; ctx->fNegOff = src->fNegOff;
	ctx->foldOpen = false;
	return PSM_OK;
}

int psm_fold_begin_device(psm_ctx* ctx, uint32_t fold, const uint32_t* window31) {
	if (!ctx || !ctx->x || !ctx->fPos || fold + 1 >= ctx->fPosOff.size()) return fail(ctx, PSM_E_ARG, "fold_begin_device: upload the fold lists first");
	CU(cudaSetDevice(ctx->device));
	const std::vector<uint32_t>&po = ctx->fPosOff, &no = ctx->fNegOff;
	const uint32_t nF = (uint32_t)po.size() - 1;
	const uint32_t nTP = po[fold + 1] - po[fold], nTN = no[fold + 1] - no[fold];
	const uint32_t P = po[nF] - nTP, Nneg = no[nF] - nTN;
	ctx->P = P; ctx->Nneg = Nneg; ctx->nTP = nTP; ctx->nTN = nTN; ctx->Nte = nTP + nTN;
	ctx->haveKnn = false; ctx->batchOpen = false; ctx->localk = 0; ctx->foldLabelsReady = false;
	ctx->hTestIdx.clear();
	ENSURE(ctx->dPos, std::max<size_t>(4, (size_t)P * 4));
	ENSURE(ctx->dNeg, std::max<size_t>(4, (size_t)Nneg * 4));
	ENSURE(ctx->dTest, std::max<size_t>(4, (size_t)ctx->Nte * 4));
	ENSURE(ctx->dClass12, std::max<size_t>(16, (size_t)ctx->Nte * 16));
	const bool shuffle = window31 != nullptr && Nneg > 1;
	if (shuffle) ENSURE(ctx->sIn, (size_t)Nneg * 4);
	cudaStream_t s = ctx->stream;
	// training arrays: the other folds' blocks in ascending fold order (testtraindivider.cpp:57-80)
	uint32_t* tp = ctx->dPos.as<uint32_t>();
	uint32_t* tn = shuffle ? ctx->sIn.as<uint32_t>() : ctx->dNeg.as<uint32_t>();
	size_t op = 0, on = 0;
	for (uint32_t kf = 0; kf < nF; kf++) {
		if (kf == fold) continue;
		const size_t cp = po[kf + 1] - po[kf], cn = no[kf + 1] - no[kf];
		if (cp) CU(cudaMemcpyAsync(tp + op, ctx->fPos + po[kf], cp * 4, cudaMemcpyDeviceToDevice, s));
		if (cn) CU(cudaMemcpyAsync(tn + on, ctx->fNeg + no[kf], cn * 4, cudaMemcpyDeviceToDevice, s));
		op += cp; on += cn;
	}
	// test rows: positives first, then negatives (sampler.cpp:90-95)
	if (nTP) CU(cudaMemcpyAsync(ctx->dTest.p, ctx->fPos + po[fold], (size_t)nTP * 4, cudaMemcpyDeviceToDevice, s));
	if (nTN) CU(cudaMemcpyAsync(ctx->dTest.as<uint32_t>() + nTP, ctx->fNeg + no[fold], (size_t)nTN * 4, cudaMemcpyDeviceToDevice, s));
	if (shuffle) {                                                              // testtraindivider.cpp:90 on the device (shuffle.cu)
		const uint64_t rBlocks = ((uint64_t)Nneg + 4095) / 4096;
		ENSURE(ctx->sKeysA, (size_t)Nneg * 8); ENSURE(ctx->sKeysB, (size_t)Nneg * 8);
		ENSURE(ctx->sValsA, (size_t)Nneg * 4); ENSURE(ctx->sValsB, (size_t)Nneg * 4);
		ENSURE(ctx->sHist, (rBlocks + 1) * 256 * 4); ENSURE(ctx->sJ, (size_t)Nneg * 4); ENSURE(ctx->sDraws, (size_t)Nneg * 4);
		ENSURE(ctx->dRandScratch, (61 + 64 * 31) * 4);
		ProfScope ps(ctx, K_GATHER, 3 + 3 * shuffle_sort_passes(Nneg));
		if (launch_glibc_rand_fill(window31, (unsigned long long)Nneg - 1, ctx->dRandScratch.as<uint32_t>(), ctx->sDraws.as<int32_t>(), s))
			return fail(ctx, PSM_E_CUDA, "fold_begin_device: rand replay launch failed");
		if (launch_device_shuffle(ctx->sDraws.as<int32_t>(), ctx->sIn.as<uint32_t>(), Nneg, ctx->sKeysA.as<unsigned long long>(), ctx->sKeysB.as<unsigned long long>(),
		                          ctx->sValsA.as<uint32_t>(), ctx->sValsB.as<uint32_t>(), ctx->sHist.as<uint32_t>(), ctx->sJ.as<uint32_t>(), ctx->dNeg.as<uint32_t>(), s))
			return fail(ctx, PSM_E_CUDA, "fold_begin_device: shuffle launch failed");
	}
	CU(cudaMemsetAsync(ctx->dClass12.p, 0, std::max<size_t>(16, (size_t)ctx->Nte * 16), s));
	CU(cudaGetLastError());
	ctx->foldOpen = true; ctx->foldFromLists = true;
	return PSM_OK;
}

// host copy of the open fold's test ids (psm_fold_begin keeps one; psm_fold_begin_device fetches it only when somebody asks)
static int ensure_host_test_ids(psm_ctx* ctx) {
	if (ctx->hTestIdx.size() == ctx->Nte) return PSM_OK;
	ctx->hTestIdx.resize(ctx->Nte);
	if (ctx->Nte) {
		CU(cudaMemcpyAsync(ctx->hTestIdx.data(), ctx->dTest.p, (size_t)ctx->Nte * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
	}
	return PSM_OK;
}

int psm_fold_indices_get(psm_ctx* ctx, uint32_t* sizes4, uint32_t* trngPos, uint32_t* trngNeg, uint32_t* testPos, uint32_t* testNeg) {
	if (!ctx || !ctx->foldOpen) return fail(ctx, PSM_E_ARG, "fold_indices_get: no open fold");
	if (sizes4) { sizes4[0] = ctx->P; sizes4[1] = ctx->Nneg; sizes4[2] = ctx->nTP; sizes4[3] = ctx->nTN; }
	cudaStream_t s = ctx->stream;
	if (trngPos && ctx->P) CU(cudaMemcpyAsync(trngPos, ctx->dPos.p, (size_t)ctx->P * 4, cudaMemcpyDeviceToHost, s));
	if (trngNeg && ctx->Nneg) CU(cudaMemcpyAsync(trngNeg, ctx->dNeg.p, (size_t)ctx->Nneg * 4, cudaMemcpyDeviceToHost, s));
	if (testPos && ctx->nTP) CU(cudaMemcpyAsync(testPos, ctx->dTest.p, (size_t)ctx->nTP * 4, cudaMemcpyDeviceToHost, s));
	if (testNeg && ctx->nTN) CU(cudaMemcpyAsync(testNeg, ctx->dTest.as<uint32_t>() + ctx->nTP, (size_t)ctx->nTN * 4, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	return PSM_OK;
}

int psm_fold_knn(psm_ctx* ctx, uint32_t k) {
	if (!ctx || !ctx->foldOpen) return fail(ctx, PSM_E_ARG, "fold_knn: no open fold");
	if (ctx->P < 2) return fail(ctx, PSM_E_TOOFEW_POS, "fold_knn: fewer than two training positives");
	CU(cudaSetDevice(ctx->device));
	const uint32_t localk = std::min(k, ctx->P);                                // samplerKNN.cpp:52
	if (localk == 0 || localk > 64) return fail(ctx, PSM_E_LIMIT, "fold_knn: k must be in 1..64");
	const uint32_t P = ctx->P, m = ctx->m;
	ENSURE(ctx->dXp, (size_t)P * m * 8);
	ENSURE(ctx->dNN, (size_t)P * localk * 4);
	ENSURE(ctx->dNNd, (size_t)P * localk * 8);
	{
		ProfScope ps(ctx, K_GATHER, 1);
		launch_gather_rows(ctx->x, m + 1, m, ctx->dPos.as<uint32_t>(), P, ctx->dXp.as<double>(), ctx->stream);
	}
	// large folds: candidates from the tensor cores, exact fp64 rerank (knn_tc.cu); the exact kernel below is the fallback
	ctx->knnPath = 0; ctx->knnIncomplete = 0;
	const int C = knn_tensor_candidates((int)localk);
	const char* envT = getenv("PSM_KNN_TENSOR");
	const bool tensor = C > 0 && (envT ? atoi(envT) != 0 : ((uint64_t)P * P * m >= (1ull << 33) && m >= 64));
	if (tensor) {
		ENSURE(ctx->dKnnTc, knn_tensor_workspace_bytes((int)P, (int)m, C) + 256);
		ENSURE(ctx->dErr, 16);
		{
			ProfScope ps(ctx, K_KNN, 5);
			if (launch_knn_tensor(ctx->dXp.as<double>(), (int)P, (int)m, (int)localk, ctx->dKnnTc.p, ctx->dNN.as<int32_t>(), ctx->dNNd.as<double>(),
			                      ctx->dErr.as<uint32_t>() + 2, ctx->stream))
				return fail(ctx, PSM_E_CUDA, "fold_knn: tensor-core launch failed");
		}
		uint32_t bad = 0;
		CU(cudaMemcpyAsync(&bad, ctx->dErr.as<uint32_t>() + 2, 4, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
		ctx->knnIncomplete = bad;
		ctx->knnPath = bad ? 2 : 1;
		if (!bad) { ctx->localk = localk; ctx->haveKnn = true; return PSM_OK; }
	}
	size_t scratchElems = std::min<size_t>((size_t)P * P, std::max<size_t>((size_t)P * 64, (size_t)1 << 27));
	ENSURE(ctx->dKnnScratch, scratchElems * 8);
	{
		size_t qb = std::min<size_t>(scratchElems / P, P);
		ProfScope ps(ctx, K_KNN, (int)(2 * ((P + qb - 1) / qb)));
		int rc = launch_knn(ctx->dXp.as<double>(), (int)P, (int)m, (int)localk, ctx->dKnnScratch.as<double>(), scratchElems,
		                    ctx->dNN.as<int32_t>(), ctx->dNNd.as<double>(), ctx->stream);
		if (rc) return fail(ctx, PSM_E_LIMIT, "fold_knn: launch failed");
	}
	CU(cudaGetLastError());
	ctx->localk = localk; ctx->haveKnn = true;
	return PSM_OK;
}

int psm_fold_knn_path(psm_ctx* ctx, int* path, uint32_t* incomplete) {
	if (!ctx || !ctx->haveKnn) return fail(ctx, PSM_E_ARG, "fold_knn_path: no neighbour table");
	if (path) *path = ctx->knnPath;
	if (incomplete) *incomplete = ctx->knnIncomplete;
	return PSM_OK;
}

int psm_fold_knn_get(psm_ctx* ctx, int32_t* nnIdx, double* nnDist) {
	if (!ctx || !ctx->haveKnn || !nnIdx) return fail(ctx, PSM_E_ARG, "fold_knn_get: no neighbour table");
	size_t cnt = (size_t)ctx->P * ctx->localk;
	CU(cudaMemcpyAsync(nnIdx, ctx->dNN.p, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (nnDist) {
		if (!ctx->dNNd.p) return fail(ctx, PSM_E_ARG, "fold_knn_get: distances not available for an injected table");
		CU(cudaMemcpyAsync(nnDist, ctx->dNNd.p, cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}

int psm_fold_knn_set(psm_ctx* ctx, const int32_t* nnIdx, uint32_t localk) {
	if (!ctx || !ctx->foldOpen || !nnIdx || localk == 0) return fail(ctx, PSM_E_ARG, "fold_knn_set: bad arguments");
	size_t cnt = (size_t)ctx->P * localk;
	ENSURE(ctx->dNN, cnt * 4);
	CU(cudaMemcpyAsync(ctx->dNN.p, nnIdx, cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	ctx->localk = localk; ctx->haveKnn = true;
	return PSM_OK;
}

// ------------------------------------------------------------------------------------------ batch
int psm_batch_begin(psm_ctx* ctx, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t p0, uint32_t p1) {
	if (!ctx || !ctx->foldOpen) return fail(ctx, PSM_E_ARG, "batch_begin: no open fold");
	if (nParts == 0 || p0 >= p1 || p1 > nParts) return fail(ctx, PSM_E_ARG, "batch_begin: bad partition range");
	if (ctx->P < 2) return fail(ctx, PSM_E_TOOFEW_POS, "batch_begin: fewer than two training positives");
	CU(cudaSetDevice(ctx->device));
	ctx->nParts = nParts; ctx->fp = fp; ctx->ratio = ratio; ctx->p0 = p0; ctx->p1 = p1; ctx->nB = p1 - p0;
	ctx->hParts.resize(ctx->nB);
	uint64_t dOff = 0, sOff = 0, drOff = 0;
	ctx->maxNtr = 0;
	for (uint32_t i = 0; i < ctx->nB; i++) {
		PartDesc& d = ctx->hParts[i];
		int rc = partition_sizes(ctx->P, ctx->Nneg, nParts, fp, ratio, p0 + i, &d);
		if (rc) return fail(ctx, rc, "batch_begin: last negative chunk would underflow (trngNeg < ceil(trngNeg/nParts)*(nParts-1))");
		if (d.Ntr >= kMaxRows) return fail(ctx, PSM_E_LIMIT, "batch_begin: training matrix has >= 2^24 rows");
		if ((uint64_t)d.negOffset + d.nNegOut > ctx->Nneg) return fail(ctx, PSM_E_CHUNK, "batch_begin: negative chunk out of range");
		d.dOffset = dOff; d.sOffset = sOff; d.drawOffset = drOff;
		dOff += (uint64_t)d.Ntr * (ctx->m + 1);
		sOff += (uint64_t)d.Ntr * ctx->m;
		drOff += (uint64_t)ctx->P * fp * (ctx->m + 1);
		ctx->maxNtr = std::max(ctx->maxNtr, d.Ntr);
	}
	ctx->totalD = dOff; ctx->totalS = sOff; ctx->totalDraws = drOff;
	ENSURE(ctx->dParts, ctx->nB * sizeof(PartDesc));
	CU(cudaMemcpyAsync(ctx->dParts.p, ctx->hParts.data(), ctx->nB * sizeof(PartDesc), cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	ctx->batchOpen = true; ctx->assembled = false; ctx->trained = false;
	return PSM_OK;
}

static int assemble_common(psm_ctx* ctx, const int32_t* draws_host, uint64_t ndraws, const uint32_t* window31) {
	if (!ctx || !ctx->batchOpen) return fail(ctx, PSM_E_ARG, "batch_assemble: no open batch");
	if (ctx->fp > 0) {
		if (!ctx->haveKnn) return fail(ctx, PSM_E_ARG, "batch_assemble: run psm_fold_knn first");
		if (ctx->localk < 2) return fail(ctx, PSM_E_LOCALK, "batch_assemble: min(k,P) < 2");
		if (!window31 && (!draws_host || ndraws < ctx->totalDraws)) return fail(ctx, PSM_E_ARG, "batch_assemble: too few SMOTE draws");
	}
	CU(cudaSetDevice(ctx->device));
	ENSURE(ctx->dD, std::max<uint64_t>(8, ctx->totalD * 8));
	ENSURE(ctx->dErr, 16);
	ENSURE(ctx->dDraws, std::max<uint64_t>(4, ctx->totalDraws * 4));
	if (ctx->totalDraws) {
		if (window31) {
			ENSURE(ctx->dRandScratch, (61 + 64 * 31) * 4);
			ProfScope ps(ctx, K_GATHER, 1);
			if (launch_glibc_rand_fill(window31, ctx->totalDraws, ctx->dRandScratch.as<uint32_t>(), ctx->dDraws.as<int32_t>(), ctx->stream))
				return fail(ctx, PSM_E_CUDA, "batch_assemble: rand replay launch failed");
		} else {
			CU(cudaMemcpyAsync(ctx->dDraws.p, draws_host, ctx->totalDraws * 4, cudaMemcpyHostToDevice, ctx->stream));
		}
	}
	CU(cudaMemsetAsync(ctx->dErr.p, 0, 16, ctx->stream));
	{
		ProfScope ps(ctx, K_ASSEMBLE, 1);
		launch_assemble(ctx->x, ctx->m, ctx->dPos.as<uint32_t>(), ctx->P, ctx->dNeg.as<uint32_t>(), ctx->dNN.as<int32_t>(),
		                std::max(ctx->localk, 2u), ctx->fp, ctx->dDraws.as<int32_t>(), ctx->dParts.as<PartDesc>(), ctx->nB, ctx->maxNtr,
		                ctx->dD.as<double>(), ctx->dErr.as<int>(), ctx->stream);
	}
	CU(cudaGetLastError());
	int herr = 0;
	CU(cudaMemcpyAsync(&herr, ctx->dErr.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));   // also keeps `draws_host` borrowed only for the call
	if (herr) return fail(ctx, PSM_E_NEIGHBOUR, "batch_assemble: SMOTE needs a neighbour that does not exist (duplicate positives)");
	ctx->assembled = true; ctx->trained = false;
	return PSM_OK;
}

int psm_batch_assemble(psm_ctx* ctx, const int32_t* draws_host, uint64_t ndraws) { return assemble_common(ctx, draws_host, ndraws, nullptr); }

int psm_batch_assemble_glibc(psm_ctx* ctx, const uint32_t* window31) {
	if (!window31) return fail(ctx, PSM_E_ARG, "batch_assemble_glibc: null window");
	return assemble_common(ctx, nullptr, 0, window31);
}

int psm_glibc_jump(uint32_t* window31, uint64_t n) {
	if (!window31) return PSM_E_ARG;
	glibc_window_jump(window31, n);
	return PSM_OK;
}

int psm_glibc_fill_device(psm_ctx* ctx, const uint32_t* window31, uint64_t count, uint64_t skip, int32_t* out, uint64_t nOut) {
	if (!ctx || !window31 || count == 0 || skip + nOut > count || (nOut && !out)) return fail(ctx, PSM_E_ARG, "glibc_fill_device: bad arguments");
	CU(cudaSetDevice(ctx->device));
	ENSURE(ctx->dDraws, (size_t)count * 4);
	ENSURE(ctx->dRandScratch, (61 + 64 * 31) * 4);
	if (launch_glibc_rand_fill(window31, count, ctx->dRandScratch.as<uint32_t>(), ctx->dDraws.as<int32_t>(), ctx->stream))
		return fail(ctx, PSM_E_LIMIT, "glibc_fill_device: rand replay launch failed (count above 248 * 2^24?)");
	CU(cudaGetLastError());
	if (nOut) CU(cudaMemcpyAsync(out, ctx->dDraws.as<int32_t>() + skip, (size_t)nOut * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}

int psm_batch_get_training(psm_ctx* ctx, uint32_t p, uint32_t* sizes, double* out) {
	if (!ctx || !ctx->assembled || p < ctx->p0 || p >= ctx->p1) return fail(ctx, PSM_E_ARG, "batch_get_training: partition not assembled");
	const PartDesc& d = ctx->hParts[p - ctx->p0];
	if (sizes) { sizes[0] = d.nPosOut; sizes[1] = d.nNegOut; sizes[2] = d.Ntr; }
	if (out) {
		const uint32_t ld = ctx->m + 1;
		std::vector<double> col((size_t)d.Ntr * ld);
		CU(cudaMemcpyAsync(col.data(), ctx->dD.as<double>() + d.dOffset, col.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
		for (uint32_t c = 0; c < ld; c++)
			for (uint32_t r = 0; r < d.Ntr; r++) out[(size_t)r * ld + c] = col[(size_t)c * d.Ntr + r];
	}
	return PSM_OK;
}

int psm_batch_train(psm_ctx* ctx, uint32_t T, uint32_t mtry, const uint32_t* forest_seeds, const uint64_t* inj_bootstrap,
                    const uint32_t* inj_cand, uint32_t inj_cand_nodes) {
	if (!ctx || !ctx->assembled) return fail(ctx, PSM_E_ARG, "batch_train: batch not assembled");
	if (T == 0 || mtry == 0 || mtry > ctx->m) return fail(ctx, PSM_E_ARG, "batch_train: need 1 <= mtry <= m and nTrees >= 1");
	if (!forest_seeds && !(inj_bootstrap && inj_cand)) return fail(ctx, PSM_E_ARG, "batch_train: forest_seeds required unless all draws are injected");
	if (tree_smem_bytes(ctx->m, mtry) > 200 * 1024) return fail(ctx, PSM_E_LIMIT, "batch_train: m too large for the candidate generator");
	CU(cudaSetDevice(ctx->device));
	const uint32_t m = ctx->m, nB = ctx->nB, stride = (ctx->maxNtr + 3) & ~3u;   // multiple of 4: keeps every list 16-byte aligned
	const uint32_t nTrees = nB * T;
	const uint32_t NC = 2 * stride + 2;   // even
	const uint64_t itemCap64 = (uint64_t)nTrees * (stride / 11 + 2);
	if (itemCap64 >= (1ull << 31)) return fail(ctx, PSM_E_LIMIT, "batch_train: batch too large (reduce the batch budget)");
	const uint32_t itemCap = (uint32_t)itemCap64;
	ctx->T = T; ctx->mtry = mtry;

	ENSURE(ctx->dS, ctx->totalS * 8);
	const bool sortInSmem = ctx->maxNtr <= 16384 && !getenv("PSM_FORCE_RADIX_SORT");
	if (!sortInSmem) {
		ENSURE(ctx->dKeysA, ctx->totalS * 8); ENSURE(ctx->dKeysB, ctx->totalS * 8);
		ENSURE(ctx->dValsA, ctx->totalS * 4); ENSURE(ctx->dValsB, ctx->totalS * 4);
	}
	int entry32 = (ctx->maxNtr <= Ent<uint32_t>::kMaxRowsE && !getenv("PSM_FORCE_ENTRY64")) ? 1 : 0;
	size_t listBytes = (size_t)nTrees * m * stride * (entry32 ? 4 : 8);
	ENSURE(ctx->dLists0, listBytes); ENSURE(ctx->dLists1, listBytes);
	ENSURE(ctx->dInbag, (size_t)nTrees * stride * 4);
	ENSURE(ctx->dMt, (size_t)nTrees * 312 * 8);
	ENSURE(ctx->dTs, (size_t)nTrees * sizeof(TreeState));
	ENSURE(ctx->dNodes, (size_t)nTrees * NC * sizeof(Node16));
	ENSURE(ctx->dNodes8, (size_t)nTrees * NC * 8);
	const uint32_t flagWords = (stride + 31) / 32;
	ENSURE(ctx->dFlagBits, (size_t)nTrees * flagWords * 4);
	ENSURE(ctx->dItems0, (size_t)itemCap * sizeof(Item)); ENSURE(ctx->dItems1, (size_t)itemCap * sizeof(Item));
	ENSURE(ctx->dCls, (size_t)2 * NCLS * itemCap * 4);
	ENSURE(ctx->dCand, (size_t)itemCap * mtry * 4);
	ENSURE(ctx->dRes, (size_t)itemCap * mtry * sizeof(SearchRes));
	ENSURE(ctx->dDec, (size_t)itemCap * sizeof(SplitDec));
	ENSURE(ctx->dPtask, (size_t)itemCap * sizeof(PartTask));
	ENSURE(ctx->dSplitFeat, (size_t)itemCap * 4);
	ENSURE(ctx->dCounters, CNT_COUNT * 4);
	ENSURE(ctx->dStats, STAT_COUNT * 8);
	ENSURE(ctx->dSeeds, (size_t)nB * 4);

	TrainArgs& A = ctx->A;
	A.nTrees = nTrees; A.T = T; A.m = m; A.mtry = mtry; A.stride = stride; A.NC = NC; A.itemCap = itemCap;
	A.injCandNodes = inj_cand ? inj_cand_nodes : 0; A.listBuf = 0;
	A.parts = ctx->dParts.as<PartDesc>(); A.Dall = ctx->dD.as<double>(); A.Sall = ctx->dS.as<entry_t>();
	A.lists[0] = ctx->dLists0.p; A.lists[1] = ctx->dLists1.p; A.entry32 = entry32;
	A.inbag = ctx->dInbag.as<uint32_t>(); A.mt = ctx->dMt.as<unsigned long long>(); A.ts = ctx->dTs.as<TreeState>();
	A.nodes = ctx->dNodes.as<Node16>();
	A.flagBits = ctx->dFlagBits.as<uint32_t>(); A.flagWords = flagWords;
	A.items[0] = ctx->dItems0.as<Item>(); A.items[1] = ctx->dItems1.as<Item>();
	for (int b2 = 0; b2 < 2; b2++)
		for (int k = 0; k < NCLS; k++) A.clsItems[b2][k] = ctx->dCls.as<uint32_t>() + ((size_t)b2 * NCLS + k) * itemCap;
	A.useClasses = getenv("PSM_NO_SIZE_CLASSES") ? 0 : 1;
	A.cand = ctx->dCand.as<uint32_t>(); A.res = ctx->dRes.as<SearchRes>(); A.dec = ctx->dDec.as<SplitDec>(); A.ptask = ctx->dPtask.as<PartTask>(); A.splitFeat = ctx->dSplitFeat.as<uint32_t>();
	A.counters = ctx->dCounters.as<uint32_t>(); A.stats = ctx->dStats.as<unsigned long long>();
	A.injCand = nullptr;

	cudaStream_t s = ctx->stream;
	CU(cudaMemsetAsync(A.inbag, 0, (size_t)nTrees * stride * 4, s));
	CU(cudaMemsetAsync(A.ts, 0, (size_t)nTrees * sizeof(TreeState), s));
	CU(cudaMemsetAsync(A.counters, 0, CNT_COUNT * 4, s));
	CU(cudaMemsetAsync(A.stats, 0, STAT_COUNT * 8, s));
	if (forest_seeds) CU(cudaMemcpyAsync(ctx->dSeeds.p, forest_seeds, (size_t)nB * 4, cudaMemcpyHostToDevice, s));
	if (inj_cand) {
		size_t bytes = (size_t)nTrees * inj_cand_nodes * mtry * 4;
		ENSURE(ctx->dInjCand, bytes);
		CU(cudaMemcpyAsync(ctx->dInjCand.p, inj_cand, bytes, cudaMemcpyHostToDevice, s));
		A.injCand = ctx->dInjCand.as<uint32_t>();
	}
	{
		ProfScope ps(ctx, K_SORT, 1);
		if (sortInSmem) {
			ENSURE(ctx->dSortFlags, (size_t)nB * m * 4);
			if (launch_feature_sort_smem(A.Dall, A.parts, nB, m, ctx->maxNtr, ctx->dS.as<entry_t>(), ctx->dSortFlags.as<uint32_t>(), nullptr, s))
				return fail(ctx, PSM_E_CUDA, "batch_train: cannot configure the shared-memory sort");
		} else {
			launch_feature_sort(A.Dall, A.parts, nB, m, ctx->dKeysA.as<unsigned long long>(), ctx->dKeysB.as<unsigned long long>(),
			                    ctx->dValsA.as<uint32_t>(), ctx->dValsB.as<uint32_t>(), ctx->dS.as<entry_t>(), s);
		}
	}
	if (!(inj_bootstrap && inj_cand)) {
		if (!forest_seeds) return fail(ctx, PSM_E_ARG, "batch_train: forest_seeds required");
		ProfScope ps(ctx, K_BOOT, 1);
		launch_bootstrap(A, ctx->dSeeds.as<uint32_t>(), s);   // also leaves each tree's mt19937_64 state after the bootstrap draws
	}
	if (inj_bootstrap) {
		std::vector<uint64_t> off(nB);
		uint64_t o = 0;
		for (uint32_t i = 0; i < nB; i++) { off[i] = o; o += (uint64_t)T * ctx->hParts[i].Ntr; }
		ENSURE(ctx->dInjBoot, o * 8); ENSURE(ctx->dInjBootOff, (size_t)nB * 8);
		CU(cudaMemcpyAsync(ctx->dInjBoot.p, inj_bootstrap, o * 8, cudaMemcpyHostToDevice, s));
		CU(cudaMemcpyAsync(ctx->dInjBootOff.p, off.data(), (size_t)nB * 8, cudaMemcpyHostToDevice, s));
		CU(cudaMemsetAsync(A.inbag, 0, (size_t)nTrees * stride * 4, s));
		ProfScope ps(ctx, K_BOOT, 1);
		launch_inbag_from_ids(A, ctx->dInjBoot.as<unsigned long long>(), ctx->dInjBootOff.as<uint64_t>(), s);
		CU(cudaStreamSynchronize(s));   // `off` is a local
	}
	const bool listsMulti = entry32 && T <= 16 && !getenv("PSM_LISTS_PER_TREE");
	if (listsMulti) ENSURE(ctx->dInbagT, (size_t)nB * stride * 16);
	{
		ProfScope ps(ctx, K_LISTS, listsMulti ? 2 : 1);
		launch_build_lists(A, listsMulti ? ctx->dInbagT.as<unsigned char>() : nullptr, s);
	}
	if (entry32) {
		// the compact 4-byte entries hold multiplicities up to 15; fall back to the 8-byte layout if a tree exceeds that
		CU(cudaMemcpyAsync(ctx->hCounters, A.counters, CNT_COUNT * 4, cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));
		if (ctx->hCounters[CNT_ERROR] & ERR_COUNT_LIMIT) {
			entry32 = 0;
			listBytes = (size_t)nTrees * m * stride * 8;
			ENSURE(ctx->dLists0, listBytes); ENSURE(ctx->dLists1, listBytes);
			A.lists[0] = ctx->dLists0.p; A.lists[1] = ctx->dLists1.p; A.entry32 = 0;
			CU(cudaMemsetAsync(A.counters, 0, CNT_COUNT * 4, s));
			ProfScope ps(ctx, K_LISTS, 1);
			launch_build_lists(A, nullptr, s);
		}
	}
	{
		ProfScope ps(ctx, K_FINALIZE, 1);
		if (launch_init_trees(A, s)) return fail(ctx, PSM_E_CUDA, "batch_train: cannot configure shared memory");
	}
	CU(cudaGetLastError());

	int buf = 0;
	uint64_t levels = 0;
	unsigned long long tracePrevBytes = 0;   // PSM_TRACE_LEVELS: split bytes up to the previous level (per call: fold lanes train concurrently)
	// The next level's item count is read back right after level_finalize and BEFORE the list partition is queued, and the
	// host waits on an event instead of the stream: the round trip (copy, wake-up, five launches) overlaps the partition
	// kernel of the level that is still running instead of leaving the GPU idle between levels.
	if (!ctx->evLevel) CU(cudaEventCreateWithFlags(&ctx->evLevel, cudaEventDisableTiming));
	CU(cudaMemcpyAsync(ctx->hCounters, A.counters, CNT_COUNT * 4, cudaMemcpyDeviceToHost, s));
	CU(cudaEventRecord(ctx->evLevel, s));
	while (true) {
		CU(cudaEventSynchronize(ctx->evLevel));
		const uint32_t nItems = ctx->hCounters[CNT_NEXT_ITEMS];
		const uint32_t e = ctx->hCounters[CNT_ERROR];
		if (e & ERR_COUNT_LIMIT) return fail(ctx, PSM_E_LIMIT, "batch_train: a bootstrap multiplicity exceeds 32767");
		if (e & ERR_POOL_OVERFLOW) return fail(ctx, PSM_E_OVERFLOW, "batch_train: node or work-item pool overflow");
		if (e & ERR_INJ_CAND) return fail(ctx, PSM_E_ARG, "batch_train: injected candidate table has fewer nodes than a tree grew");
		if (nItems == 0) break;
		if (nItems > itemCap) return fail(ctx, PSM_E_OVERFLOW, "batch_train: work-item pool overflow");
		uint32_t clsCount[NCLS], clsSum = 0;
		for (int k = 0; k < NCLS; k++) { clsCount[k] = ctx->hCounters[CNT_CLS0 + k]; clsSum += clsCount[k]; }
		if (clsSum != nItems) return fail(ctx, PSM_E_OVERFLOW, "batch_train: size-class lists do not add up to the level's work items");
		static_assert(CNT_NEXT_ITEMS == 0 && CNT_CLS0 == 1, "the per-level counters are cleared with one memset");
		// (the merged split-search launch clears them itself: one stream operation less per level)
		if (getenv("PSM_SS_SEPARATE")) CU(cudaMemsetAsync(A.counters, 0, (1 + NCLS) * 4, s));
		const bool trace = getenv("PSM_TRACE_LEVELS") != nullptr;
		cudaEvent_t tev[5];
		if (trace) for (auto& e : tev) cudaEventCreate(&e);
		if (trace) cudaEventRecord(tev[0], s);
		{ ProfScope ps(ctx, K_SEARCH, 0); ps.add_launches(launch_split_search(A, clsCount, buf, s)); }
		if (trace) cudaEventRecord(tev[1], s);
		{ ProfScope ps(ctx, K_SELECT, 1); launch_split_select(A, nItems, buf, s); }
		if (trace) cudaEventRecord(tev[2], s);
		{
			ProfScope ps(ctx, K_FINALIZE, 1);
			if (launch_tree_level(A, buf, s)) return fail(ctx, PSM_E_CUDA, "batch_train: cannot configure the flag bitmap");
		}
		if (trace) cudaEventRecord(tev[3], s);
		CU(cudaMemcpyAsync(ctx->hCounters, A.counters, CNT_COUNT * 4, cudaMemcpyDeviceToHost, s));   // nItems of the next level
		CU(cudaEventRecord(ctx->evLevel, s));
		{ ProfScope ps(ctx, K_PARTITION, 1); launch_partition_lists(A, nItems, buf, s); }
		if (trace) {
			cudaEventRecord(tev[4], s);
			cudaStreamSynchronize(s);
			float t[4];
			for (int i = 0; i < 4; i++) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
			unsigned long long hs2[STAT_COUNT];
			cudaMemcpy(hs2, A.stats, STAT_COUNT * 8, cudaMemcpyDeviceToHost);
			if (levels == 0) tracePrevBytes = 0;
			unsigned long long& prevBytes = tracePrevBytes;
			const unsigned long long samples = (hs2[STAT_SPLIT_BYTES] - prevBytes) / ((unsigned long long)mtry * 8ull + 8ull);
			prevBytes = hs2[STAT_SPLIT_BYTES];
			fprintf(stderr, "[psm level %3llu] items %7u samples %9llu  search %8.1f us  select %7.1f us  finalize %7.1f us  partition %8.1f us\n",
			        (unsigned long long)levels, nItems, samples, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3);
			for (auto& e : tev) cudaEventDestroy(e);
		}
		CU(cudaGetLastError());
		buf ^= 1; A.listBuf ^= 1; levels++;
	}
	unsigned long long hs[STAT_COUNT];
	CU(cudaMemcpyAsync(hs, A.stats, STAT_COUNT * 8, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	ctx->statSplitBytes += hs[STAT_SPLIT_BYTES];
	ctx->statNodes += hs[STAT_NODES] + nTrees;
	ctx->statLevels += levels;
	{   // largest tree of the batch: sizes the shared-memory tree buffers of the predict kernel
		std::vector<TreeState> hts(nTrees);
		CU(cudaMemcpyAsync(hts.data(), A.ts, (size_t)nTrees * sizeof(TreeState), cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));
		ctx->maxNodes = 0;
		for (const TreeState& t : hts) ctx->maxNodes = std::max(ctx->maxNodes, t.nNodes);
	}
	ctx->trained = true; ctx->inbagValid = true;
	return PSM_OK;
}

static int fetch_tree(psm_ctx* ctx, uint32_t p, uint32_t t, TreeState* ts, std::vector<Node16>* nodes) {
	if (!ctx || !ctx->trained || p < ctx->p0 || p >= ctx->p1 || t >= ctx->T) return fail(ctx, PSM_E_ARG, "tree export: no such trained tree");
	const size_t tree = (size_t)(p - ctx->p0) * ctx->T + t;
	CU(cudaMemcpyAsync(ts, ctx->A.ts + tree, sizeof(TreeState), cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	if (nodes) {
		nodes->resize(ts->nNodes);
		CU(cudaMemcpyAsync(nodes->data(), ctx->A.nodes + tree * ctx->A.NC, (size_t)ts->nNodes * sizeof(Node16), cudaMemcpyDeviceToHost, ctx->stream));
		CU(cudaStreamSynchronize(ctx->stream));
	}
	return PSM_OK;
}

int psm_batch_tree_nodes(psm_ctx* ctx, uint32_t p, uint32_t t, uint64_t* nNodes) {
	if (!nNodes) return fail(ctx, PSM_E_ARG, "tree_nodes: null output");
	TreeState ts;
	int rc = fetch_tree(ctx, p, t, &ts, nullptr);
	if (rc) return rc;
	*nNodes = ts.nNodes;
	return PSM_OK;
}

int psm_batch_export_tree(psm_ctx* ctx, uint32_t p, uint32_t t, uint64_t* left, uint64_t* right, uint64_t* var, double* value, double* frac) {
	TreeState ts;
	std::vector<Node16> nodes;
	int rc = fetch_tree(ctx, p, t, &ts, &nodes);
	if (rc) return rc;
	for (uint32_t i = 0; i < ts.nNodes; i++) {
		const Node16& nd = nodes[i];
		if (node_is_leaf(nd)) {
			left[i] = right[i] = 0; var[i] = 0; value[i] = 0.0;
			unsigned long long bits = nd.b & 0x7FFFFFFFFFFFFFFFull;
			double f1; memcpy(&f1, &bits, 8);
			frac[2 * i] = nd.a; frac[2 * i + 1] = f1;
		} else {
			left[i] = nd.b >> 32; right[i] = left[i] + 1; var[i] = nd.b & 0xFFFFFFFFull; value[i] = nd.a;
			frac[2 * i] = frac[2 * i + 1] = std::nan("");
		}
	}
	return PSM_OK;
}

int psm_batch_import_forests(psm_ctx* ctx, uint32_t p0, uint32_t p1, uint32_t T, const uint64_t* nNodes, const uint64_t* left,
                             const uint64_t* right, const uint64_t* var, const double* value, const double* frac) {
	if (!ctx || !ctx->foldOpen) return fail(ctx, PSM_E_ARG, "import_forests: no open fold");
	if (p1 <= p0 || T == 0 || !nNodes || !left || !right || !var || !value || !frac) return fail(ctx, PSM_E_ARG, "import_forests: bad arguments");
	CU(cudaSetDevice(ctx->device));
	const uint32_t nB = p1 - p0;
	const uint64_t nTrees64 = (uint64_t)nB * T;
	if (nTrees64 >= (1ull << 31)) return fail(ctx, PSM_E_LIMIT, "import_forests: too many trees in one batch");
	const uint32_t nTrees = (uint32_t)nTrees64;
	uint64_t maxN = 0;
	for (uint32_t t = 0; t < nTrees; t++) {
		if (nNodes[t] == 0 || nNodes[t] >= (1ull << 31)) return fail(ctx, PSM_E_ARG, "import_forests: a tree has no nodes or too many");
		maxN = std::max(maxN, nNodes[t]);
	}
	const uint32_t NC = (uint32_t)((maxN + 1) & ~1ull);   // even, like the training layout
	std::vector<Node16> hn((size_t)nTrees * NC);
	std::vector<TreeState> hts(nTrees);
	memset(hn.data(), 0, hn.size() * sizeof(Node16));
	memset(hts.data(), 0, hts.size() * sizeof(TreeState));
	size_t o = 0;
	for (uint32_t t = 0; t < nTrees; t++) {
		const uint64_t N = nNodes[t];
		hts[t].nNodes = (uint32_t)N;
		Node16* out = hn.data() + (size_t)t * NC;
		for (uint64_t i = 0; i < N; i++, o++) {
			if (left[o] == 0 && right[o] == 0) {                               // terminal: both children 0 (Tree.cpp:152-156)
				const double f0 = frac[2 * o], f1 = frac[2 * o + 1];
				if (!(f0 >= 0.0) || !(f1 >= 0.0)) return fail(ctx, PSM_E_ARG, "import_forests: a terminal node has no class fractions");
				unsigned long long bits; memcpy(&bits, &f1, 8);
				out[i].a = f0; out[i].b = bits | 0x8000000000000000ull;
			} else {
				if (right[o] != left[o] + 1 || right[o] >= N || left[o] <= i) return fail(ctx, PSM_E_ARG, "import_forests: child ids are not ranger's (consecutive, after the parent)");
				if (var[o] >= ctx->m) return fail(ctx, PSM_E_ARG, "import_forests: split variable out of range");
				out[i].a = value[o]; out[i].b = (unsigned long long)var[o] | ((unsigned long long)left[o] << 32);
			}
		}
	}
	ENSURE(ctx->dNodes, hn.size() * sizeof(Node16));
	ENSURE(ctx->dNodes8, hn.size() * 8);
	ENSURE(ctx->dTs, (size_t)nTrees * sizeof(TreeState));
	CU(cudaMemcpyAsync(ctx->dNodes.p, hn.data(), hn.size() * sizeof(Node16), cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaMemcpyAsync(ctx->dTs.p, hts.data(), (size_t)nTrees * sizeof(TreeState), cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));   // hn / hts are locals
	TrainArgs& A = ctx->A;
	A.nodes = ctx->dNodes.as<Node16>(); A.ts = ctx->dTs.as<TreeState>(); A.NC = NC; A.nTrees = nTrees; A.T = T;
	ctx->nB = nB; ctx->p0 = p0; ctx->p1 = p1; ctx->T = T; ctx->maxNodes = (uint32_t)maxN;
	ctx->batchOpen = false; ctx->assembled = false; ctx->trained = true;
	ctx->inbagValid = false;   // imported forests carry no bootstrap counts (psm_batch_get_inbag refuses)
	return PSM_OK;
}

int psm_batch_get_inbag(psm_ctx* ctx, uint32_t p, uint32_t t, uint32_t* counts) {
	if (!ctx || !ctx->trained || !ctx->inbagValid || p < ctx->p0 || p >= ctx->p1 || t >= ctx->T || !counts || (size_t)(p - ctx->p0) >= ctx->hParts.size())
		return fail(ctx, PSM_E_ARG, "get_inbag: no such tree trained in this context (imported forests have no bootstrap counts)");
	const size_t tree = (size_t)(p - ctx->p0) * ctx->T + t;
	CU(cudaMemcpyAsync(counts, ctx->A.inbag + tree * ctx->A.stride, (size_t)ctx->hParts[p - ctx->p0].Ntr * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}

int psm_batch_predict_accumulate(psm_ctx* ctx, uint32_t nPartsTotal) {
	if (!ctx || !ctx->trained || nPartsTotal == 0) return fail(ctx, PSM_E_ARG, "predict_accumulate: batch not trained");
	if (ctx->Nte == 0) return PSM_OK;
	CU(cudaSetDevice(ctx->device));
	ENSURE(ctx->dPredScratch, (size_t)ctx->nB * 2 * ctx->Nte * 8);
	{
		ProfScope ps(ctx, K_PREDICT, 3);
		launch_pack_forest(ctx->A.nodes, ctx->A.ts, ctx->A.NC, ctx->A.nTrees, ctx->dNodes8.p, ctx->stream);
		const double divider = 1.0 / (double)nPartsTotal;
		const int how = launch_predict(ctx->x, ctx->m, ctx->dTest.as<uint32_t>(), ctx->Nte, ctx->A.nodes, ctx->dNodes8.p, ctx->A.NC, ctx->A.ts, ctx->maxNodes, ctx->nB, ctx->T,
		                               ctx->dPredScratch.as<double>(), divider, ctx->dClass12.as<double>(), ctx->stream);
		if (how < 0) return fail(ctx, PSM_E_CUDA, "predict_accumulate: cannot configure shared memory");
		if (how == 0) launch_accumulate(ctx->dPredScratch.as<double>(), ctx->Nte, ctx->nB, divider, ctx->dClass12.as<double>(), ctx->stream);
	}
	CU(cudaGetLastError());
	return PSM_OK;
}

// ------------------------------------------------------------------------------------------ scores
int psm_fold_scores_get(psm_ctx* ctx, double* class1, double* class2) {
	if (!ctx || !ctx->foldOpen || !class1) return fail(ctx, PSM_E_ARG, "fold_scores_get: no open fold");
	if (ctx->Nte == 0) return PSM_OK;
	CU(cudaMemcpyAsync(class1, ctx->dClass12.p, (size_t)ctx->Nte * 8, cudaMemcpyDeviceToHost, ctx->stream));
	if (class2) CU(cudaMemcpyAsync(class2, ctx->dClass12.as<double>() + ctx->Nte, (size_t)ctx->Nte * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}

int psm_fold_scores_devptr(psm_ctx* ctx, double** class12_dev, uint32_t* Nte) {
	if (!ctx || !ctx->foldOpen || !class12_dev || !Nte) return fail(ctx, PSM_E_ARG, "fold_scores_devptr: no open fold");
	*class12_dev = ctx->dClass12.as<double>();
	*Nte = ctx->Nte;
	return PSM_OK;
}

int psm_fold_scatter_host(psm_ctx* ctx, double* class1Prob, double* class2Prob) {
	if (!ctx || !ctx->foldOpen || !class1Prob) return fail(ctx, PSM_E_ARG, "fold_scatter_host: no open fold");
	std::vector<double> c((size_t)ctx->Nte * 2);
	if (ctx->Nte == 0) return PSM_OK;
	{ int rc = ensure_host_test_ids(ctx); if (rc) return rc; }
	CU(cudaMemcpyAsync(c.data(), ctx->dClass12.p, c.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	for (uint32_t i = 0; i < ctx->Nte; i++) {
		class1Prob[ctx->hTestIdx[i]] += c[i];
		if (class2Prob) class2Prob[ctx->hTestIdx[i]] += c[(size_t)ctx->Nte + i];
	}
	return PSM_OK;
}

// ------------------------------------------------------------------------------------------ curves
// core: labels (u32) and scores (f64) already on the device; hScores (host copy) is only needed for tie_mode 0
// The permutation of the reference's  std::sort(idx.begin(), idx.end(), [&](size_t i, size_t j) { return preds[i] > preds[j]; })
// (src/curves.cpp:31-33), reproduced in parallel.  libstdc++'s std::sort is __introsort_loop (median-of-three quicksort that
// recurses into the RIGHT part and loops on the left, heapsort once the depth limit 2*lg(N) is spent) followed by one insertion
// sort over the blocks of <= 16 elements it leaves.  Its control flow depends only on the outcomes of the comparisons, so
//   * sorting (score, index) pairs with the same comparator moves the elements exactly like sorting the indices does, without
//     the dependent load preds[idx] in every comparison;
//   * the recursive call and the loop work on disjoint ranges: run as separate tasks they give the same array;
//   * the final insertion sort never moves an element across a block boundary (every element left of a block compares
//     not-after it) and is stable, so sorting each block where the loop leaves it gives the same order.
// The partition, insertion-sort and heap steps are libstdc++'s own (std::__unguarded_partition_pivot, std::__insertion_sort,
// std::partial_sort): the same code the reference's call instantiates.  tests/test_host_logic.py pins it against std::sort.
namespace {
struct ScoreIdx { double s; uint64_t i; };
struct ScoreDesc { bool operator()(const ScoreIdx& a, const ScoreIdx& b) const { return a.s > b.s; } };
struct SortPool {
	std::atomic<int> freeThreads;
	std::mutex m;
	std::vector<std::thread> threads;
	explicit SortPool(int n) : freeThreads(n) {}
};
void intro_loop_parallel(ScoreIdx* first, ScoreIdx* last, long depth, SortPool* pool) {
	auto cmp = __gnu_cxx::__ops::__iter_comp_iter(ScoreDesc());
	while (last - first > 16) {                               // _S_threshold
		if (depth == 0) { std::partial_sort(first, last, last, ScoreDesc()); return; }
		--depth;
		ScoreIdx* cut = std::__unguarded_partition_pivot(first, last, cmp);
		bool spawned = false;
		if (pool && last - cut > 8192 && pool->freeThreads.fetch_sub(1) > 0) {
			std::lock_guard<std::mutex> lk(pool->m);
			pool->threads.emplace_back([=]() { intro_loop_parallel(cut, last, depth, pool); pool->freeThreads.fetch_add(1); });
			spawned = true;
		} else if (pool && last - cut > 8192) pool->freeThreads.fetch_add(1);
		if (!spawned) intro_loop_parallel(cut, last, depth, pool);
		last = cut;
	}
	std::__insertion_sort(first, last, cmp);
}
void reference_sort_permutation(const double* scores, uint64_t N, uint32_t* perm, int threads) {
	std::vector<ScoreIdx> a(N);
	for (uint64_t i = 0; i < N; i++) { a[i].s = scores[i]; a[i].i = i; }
	if (N > 1) {
		if (threads <= 0) threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
		SortPool pool(4 * threads);   // tasks in flight (the top partitions are sequential: the speed-up is ~1.3x at 1M random scores, more with ties)
		intro_loop_parallel(a.data(), a.data() + N, std::__lg((long)N) * 2, threads > 1 ? &pool : nullptr);
		for (size_t t = 0;; t++) {                            // tasks may add tasks: join until the list stops growing
			std::thread th;
			{ std::lock_guard<std::mutex> lk(pool.m); if (t >= pool.threads.size()) break; th = std::move(pool.threads[t]); }
			th.join();
		}
	}
	for (uint64_t i = 0; i < N; i++) perm[i] = (uint32_t)a[i].i;
}
}  // namespace

extern "C" int psm_host_sort_permutation(const double* scores, uint64_t N, uint32_t* perm, int threads) {
	if (!scores || !perm || N >= (1ull << 32)) return PSM_E_ARG;
	reference_sort_permutation(scores, N, perm, threads);
	return PSM_OK;
}

// tie_mode 2: how far the two label-ordered tie extremes may lie apart for their midpoint to be reported (the reference's value
// lies between them, so the midpoint is within half of this of it -- against the 1e-6 bar of the north star)
static const double kTieSpreadTol = 1e-6;

static int curves_device_pass(psm_ctx* ctx, const uint32_t* dLabels, const double* dScores, uint64_t N, uint32_t totP, uint32_t totN, int order /*0 plain, 1 positives first, 2 negatives first*/,
                              bool countMixed, double* out2, unsigned long long* mixed) {
	cudaStream_t s = ctx->stream;
	const uint64_t nBlocks = (N + 4095) / 4096;
	ENSURE(ctx->cKeysA, N * 8); ENSURE(ctx->cKeysB, N * 8); ENSURE(ctx->cValsA, N * 4); ENSURE(ctx->cValsB, N * 4);
	ENSURE(ctx->cHist, (nBlocks + 1) * 256 * 4);
	{
		ProfScope ps(ctx, K_CURVES, order ? 3 + 3 * 9 : 1 + 3 * 8);
		int rc = order == 0 ? launch_sort_scores_desc(dScores, N, ctx->cKeysA.as<unsigned long long>(), ctx->cKeysB.as<unsigned long long>(),
		                                              ctx->cValsA.as<uint32_t>(), ctx->cValsB.as<uint32_t>(), ctx->cHist.as<uint32_t>(), s)
		                    : launch_sort_scores_desc_labelties(dScores, dLabels, N, order == 1, ctx->cKeysA.as<unsigned long long>(), ctx->cKeysB.as<unsigned long long>(),
		                                                        ctx->cValsA.as<uint32_t>(), ctx->cValsB.as<uint32_t>(), ctx->cHist.as<uint32_t>(), s);
		if (rc) return fail(ctx, PSM_E_ARG, "curves: sort launch failed");
	}
	{
		ProfScope ps(ctx, K_CURVES, 5 + (countMixed ? 1 : 0));
		launch_gather_labels(dLabels, ctx->cValsA.as<uint32_t>(), N, ctx->cSorted.as<unsigned char>(), s);
		if (countMixed) launch_tie_mixed_count(ctx->cKeysA.as<unsigned long long>(), ctx->cSorted.as<unsigned char>(), N, ctx->cOut.as<unsigned long long>() + 2, s);
		if (launch_curves(ctx->cSorted.as<unsigned char>(), N, totP, totN, ctx->cTp.as<uint32_t>(), ctx->cPartials.as<double>(), nBlocks * 2, ctx->cOut.as<double>(), s))
			return fail(ctx, PSM_E_ARG, "curves: launch failed");
	}
	CU(cudaGetLastError());
	unsigned long long res[3] = {0, 0, 0};
	CU(cudaMemcpyAsync(res, ctx->cOut.p, countMixed ? 24 : 16, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	memcpy(out2, res, 16);
	if (mixed) *mixed = res[2];
	return PSM_OK;
}

// core: labels (u32) and scores (f64) already on the device; hScores (host copy) is only needed for the host std::sort
static int curves_core(psm_ctx* ctx, const uint32_t* dLabels, const double* dScores, const double* hScores, uint64_t N, uint32_t totP, uint32_t totN,
                       int tie_mode, double* out) {
	if (tie_mode < 0 || tie_mode > 2) return fail(ctx, PSM_E_ARG, "curves: tie_mode must be 0 (host std::sort), 1 (stable device sort) or 2 (auto)");
	cudaStream_t s = ctx->stream;
	ENSURE(ctx->cSorted, N + 64);
	const uint64_t nBlocks = (N + 4095) / 4096;
	ENSURE(ctx->cTp, (nBlocks + 1) * 4); ENSURE(ctx->cPartials, nBlocks * 16); ENSURE(ctx->cOut, 32);
	ctx->curvesMixedPairs = 0; ctx->curvesUsedHostSort = 0; ctx->curvesSpread[0] = ctx->curvesSpread[1] = 0.0;
	if (tie_mode == 1) return curves_device_pass(ctx, dLabels, dScores, N, totP, totN, 0, false, out, nullptr);
	if (tie_mode == 2) {
		// (a) ties ordered positives first: the upper end of what any tie order can give.  No run of tied scores with both
		// labels (curves.cu: tie_mixed_count_kernel) -> the tie order cannot change a single TP/FP prefix: done, exactly.
		double hi[2], lo[2];
		unsigned long long mixed = 0;
		int rc = curves_device_pass(ctx, dLabels, dScores, N, totP, totN, 1, true, hi, &mixed);
		if (rc) return rc;
		ctx->curvesMixedPairs = mixed;
		if (mixed == 0) { out[0] = hi[0]; out[1] = hi[1]; return PSM_OK; }
		// (b) negatives first: the lower end.  Both curves are monotone in the position of every positive inside its run of
		// ties, so whatever order the reference's std::sort leaves the ties in (curves.cpp:31-33), its AUROC / AUPRC lie
		// between the two; when they are closer than kTieSpreadTol their midpoint is the answer to half of that.
		// (the negatives-first label sequence is the positives-first one with every run of ties reversed: no second sort)
		if (getenv("PSM_CURVES_TWO_SORTS")) { if ((rc = curves_device_pass(ctx, dLabels, dScores, N, totP, totN, 2, false, lo, nullptr))) return rc; }
		else {
			ENSURE(ctx->cSorted2, N + 64);
			{
				ProfScope ps(ctx, K_CURVES, 6);
				launch_tie_reverse_labels(ctx->cKeysA.as<unsigned long long>(), ctx->cSorted.as<unsigned char>(), N, ctx->cSorted2.as<unsigned char>(), s);
				if (launch_curves(ctx->cSorted2.as<unsigned char>(), N, totP, totN, ctx->cTp.as<uint32_t>(), ctx->cPartials.as<double>(), nBlocks * 2, ctx->cOut.as<double>(), s))
					return fail(ctx, PSM_E_ARG, "curves: launch failed");
			}
			CU(cudaGetLastError());
			CU(cudaMemcpyAsync(lo, ctx->cOut.p, 16, cudaMemcpyDeviceToHost, s));
			CU(cudaStreamSynchronize(s));
		}
		ctx->curvesSpread[0] = hi[0] - lo[0]; ctx->curvesSpread[1] = hi[1] - lo[1];
		if (std::fabs(hi[0] - lo[0]) <= kTieSpreadTol && std::fabs(hi[1] - lo[1]) <= kTieSpreadTol) {
			out[0] = 0.5 * (hi[0] + lo[0]); out[1] = 0.5 * (hi[1] + lo[1]);
			return PSM_OK;
		}
		// (c) the tie order matters: reproduce it (below)
	}
	{
		std::vector<uint32_t> perm;
		std::vector<double> tmp;
		if (!hScores) {
			tmp.resize(N);
			CU(cudaMemcpyAsync(tmp.data(), dScores, N * 8, cudaMemcpyDeviceToHost, s));
			CU(cudaStreamSynchronize(s));
			hScores = tmp.data();
		}
		// the permutation of the identical call the reference makes (curves.cpp:31-33): same libstdc++ introsort => same tie order
		perm.resize(N);
		reference_sort_permutation(hScores, N, perm.data(), 0);
		ENSURE(ctx->cPerm, N * 4);
		CU(cudaMemcpyAsync(ctx->cPerm.p, perm.data(), N * 4, cudaMemcpyHostToDevice, s));
		ctx->curvesUsedHostSort = 1;
		{
			ProfScope ps(ctx, K_CURVES, 5);
			launch_gather_labels(dLabels, ctx->cPerm.as<uint32_t>(), N, ctx->cSorted.as<unsigned char>(), s);
			if (launch_curves(ctx->cSorted.as<unsigned char>(), N, totP, totN, ctx->cTp.as<uint32_t>(), ctx->cPartials.as<double>(), nBlocks * 2,
			                  ctx->cOut.as<double>(), s))
				return fail(ctx, PSM_E_ARG, "curves: launch failed");
		}
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(out, ctx->cOut.p, 16, cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));   // also keeps `perm` alive until the upload is done
	}
	return PSM_OK;
}

int psm_curves_info(psm_ctx* ctx, uint64_t* mixed_tie_pairs, int* used_host_sort, double* tie_spread2) {
	if (!ctx) return PSM_E_ARG;
	if (mixed_tie_pairs) *mixed_tie_pairs = ctx->curvesMixedPairs;
	if (used_host_sort) *used_host_sort = ctx->curvesUsedHostSort;
	if (tie_spread2) { tie_spread2[0] = ctx->curvesSpread[0]; tie_spread2[1] = ctx->curvesSpread[1]; }
	return PSM_OK;
}

int psm_curves(psm_ctx* ctx, const uint32_t* labels, const double* scores, uint64_t N, int tie_mode, double* out) {
	if (!ctx || !labels || !scores || !out || N == 0 || N >= (1ull << 32)) return fail(ctx, PSM_E_ARG, "curves: bad arguments");
	CU(cudaSetDevice(ctx->device));
	uint32_t totP = 0, totN = 0;
	for (uint64_t i = 0; i < N; i++) { totP += labels[i] == 1; totN += labels[i] == 0; }   // curves.cpp:16-17
	ENSURE(ctx->cLabels, N * 4);
	ENSURE(ctx->cScores, N * 8);
	CU(cudaMemcpyAsync(ctx->cLabels.p, labels, N * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (tie_mode != 0) CU(cudaMemcpyAsync(ctx->cScores.p, scores, N * 8, cudaMemcpyHostToDevice, ctx->stream));
	return curves_core(ctx, ctx->cLabels.as<uint32_t>(), ctx->cScores.as<double>(), scores, N, totP, totN, tie_mode, out);
}

int psm_labels_upload(psm_ctx* ctx, const uint32_t* labels, uint32_t n) {
	if (!ctx || !labels || n == 0 || n != ctx->n) return fail(ctx, PSM_E_ARG, "labels_upload: need the dataset first and n labels");
	CU(cudaSetDevice(ctx->device));
	ENSURE(ctx->dGlob, (size_t)n * 16);
	CU(cudaMemsetAsync(ctx->dGlob.p, 0, (size_t)n * 16, ctx->stream));
	// a host that runs one CV after the other hands in the same labels every time: the device copy and the class totals are
	// kept when the vector is unchanged (one memcmp instead of a host copy, an H2D copy and two O(n) counting passes per CV)
	if (ctx->haveLabels && ctx->hLabels.size() == n && same_u32(ctx->hLabels.data(), labels, n)) return PSM_OK;
	ctx->haveLabels = false;
	ctx->hLabels.assign(labels, labels + n);
	ENSURE(ctx->dLabels, (size_t)n * 4);
	CU(cudaMemcpyAsync(ctx->dLabels.p, ctx->hLabels.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
	uint32_t totP = 0, totN = 0;
	for (uint32_t i = 0; i < n; i++) { totP += labels[i] == 1; totN += labels[i] == 0; }
	ctx->labelsPos = totP; ctx->labelsNeg = totN;
	CU(cudaStreamSynchronize(ctx->stream));
	ctx->haveLabels = true;
	return PSM_OK;
}

int psm_scores_reset(psm_ctx* ctx) {
	if (!ctx || !ctx->haveLabels) return fail(ctx, PSM_E_ARG, "scores_reset: upload the labels first");
	CU(cudaMemsetAsync(ctx->dGlob.p, 0, (size_t)ctx->n * 16, ctx->stream));
	return PSM_OK;
}

int psm_fold_scatter_device(psm_ctx* ctx) {
	if (!ctx || !ctx->foldOpen || !ctx->haveLabels) return fail(ctx, PSM_E_ARG, "fold_scatter_device: need an open fold and uploaded labels");
	ProfScope ps(ctx, K_GATHER, 1);
	launch_scatter_scores(ctx->dClass12.as<double>(), ctx->dTest.as<uint32_t>(), ctx->Nte, ctx->n, ctx->dGlob.as<double>(), ctx->stream);
	CU(cudaGetLastError());
	return PSM_OK;
}

int psm_fold_curves(psm_ctx* ctx, int tie_mode, double* out) {
	if (!ctx || !ctx->foldOpen || !ctx->haveLabels || !out || ctx->Nte == 0) return fail(ctx, PSM_E_ARG, "fold_curves: need an open fold and uploaded labels");
	CU(cudaSetDevice(ctx->device));
	ENSURE(ctx->dFoldLabels, (size_t)ctx->Nte * 4);
	uint32_t totP = 0, totN = 0;
	if (ctx->foldFromLists) { totP = ctx->nTP; totN = ctx->nTN; }             // the fold lists are split by label (0/1): no host pass needed
	else for (uint32_t i = 0; i < ctx->Nte; i++) { uint32_t l = ctx->hLabels[ctx->hTestIdx[i]]; totP += l == 1; totN += l == 0; }
	{
		ProfScope ps(ctx, K_GATHER, 1);
		launch_gather_u32(ctx->dLabels.as<uint32_t>(), ctx->dTest.as<uint32_t>(), ctx->Nte, ctx->dFoldLabels.as<uint32_t>(), ctx->stream);
	}
	return curves_core(ctx, ctx->dFoldLabels.as<uint32_t>(), ctx->dClass12.as<double>(), nullptr, ctx->Nte, totP, totN, tie_mode, out);
}

static int scores_fetch(psm_ctx* ctx, double* class1Prob, double* class2Prob, bool overwrite);
int psm_scores_get(psm_ctx* ctx, double* class1Prob, double* class2Prob) { return scores_fetch(ctx, class1Prob, class2Prob, false); }
int psm_scores_read(psm_ctx* ctx, double* class1Prob, double* class2Prob) { return scores_fetch(ctx, class1Prob, class2Prob, true); }
static int scores_fetch(psm_ctx* ctx, double* class1Prob, double* class2Prob, bool overwrite) {
	if (!ctx || !ctx->haveLabels || !class1Prob) return fail(ctx, PSM_E_ARG, "scores_get: no device scores");
	const size_t need = (size_t)ctx->n * 16;
	if (ctx->hStageCap < need) {                       // pinned staging buffer, reused across calls
		if (ctx->hStage) cudaFreeHost(ctx->hStage);
		ctx->hStage = nullptr; ctx->hStageCap = 0;
		CU(cudaMallocHost(&ctx->hStage, need));
		ctx->hStageCap = need;
	}
	// The copy out of the pinned staging buffer is a pure memory stream (28M doubles at BASELINE configs[2]) that costs as much
	// as the D2H copy itself, so the two are pipelined: the device arrays come over in chunks, each followed by an event, and a
	// few host threads move (or add) chunk k while chunk k+1 is still on the bus.
	const double* h = (const double*)ctx->hStage;
	const uint32_t n = ctx->n;
	const size_t total = (size_t)n * (class2Prob ? 2 : 1);                    // doubles wanted: class1 [n], then class2 [n]
	const uint32_t nChunks = total >= (1u << 20) ? 16u : 1u;
	// The copies run on a stream of their own, behind everything queued on the context's stream so far: a host that calls this
	// from a second thread (hyperSMURF::smurfIt does) overlaps the fetch with the global curves, which only read the arrays.
	// (private events: the context's event pool belongs to the thread that drives the context's stream)
	if (!ctx->copyStream) CU(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
	if (!ctx->evCopy) CU(cudaEventCreateWithFlags(&ctx->evCopy, cudaEventDisableTiming));
	CU(cudaEventRecord(ctx->evCopy, ctx->stream));
	CU(cudaStreamWaitEvent(ctx->copyStream, ctx->evCopy, 0));
	std::vector<cudaEvent_t> ev(nChunks);
	for (uint32_t c = 0; c < nChunks; c++) {
		const size_t a0 = total * c / nChunks, a1 = total * (c + 1) / nChunks;
		CU(cudaMemcpyAsync((double*)ctx->hStage + a0, ctx->dGlob.as<double>() + a0, (a1 - a0) * 8, cudaMemcpyDeviceToHost, ctx->copyStream));
		CU(cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming));
		CU(cudaEventRecord(ev[c], ctx->copyStream));
	}
	auto outPtr = [&](size_t i) { return i < n ? class1Prob + i : class2Prob + (i - n); };
	auto moveChunk = [&](uint32_t c) {
		cudaEventSynchronize(ev[c]);
		size_t a0 = total * c / nChunks;
		const size_t a1 = total * (c + 1) / nChunks;
		while (a0 < a1) {                                                      // a chunk may straddle the class1 / class2 boundary
			const size_t e = (a0 < n && a1 > n) ? n : a1;
			double* dst = outPtr(a0);
			if (overwrite) memcpy(dst, h + a0, (e - a0) * 8);
			else for (size_t i = a0; i < e; i++) dst[i - a0] += h[i];
			a0 = e;
		}
	};
	const uint32_t nThreads = std::min<uint32_t>(nChunks, 4);
	if (nThreads == 1) moveChunk(0);
	else {
		const int device = ctx->device;
		std::vector<std::thread> th;
		for (uint32_t t = 0; t < nThreads; t++) th.emplace_back([&, t]() { cudaSetDevice(device); for (uint32_t c = t; c < nChunks; c += nThreads) moveChunk(c); });
		for (auto& x : th) x.join();
	}
	for (auto e : ev) cudaEventDestroy(e);
	return PSM_OK;
}

int psm_scores_merge(psm_ctx* ctx, psm_ctx* src) {
	if (!ctx || !src || !ctx->haveLabels || !src->haveLabels || ctx->n != src->n || ctx->device != src->device)
		return fail(ctx, PSM_E_ARG, "scores_merge: both contexts need device scores of the same dataset on the same device");
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(src->stream));
	launch_add_f64(ctx->dGlob.as<double>(), src->dGlob.as<double>(), (size_t)ctx->n * 2, ctx->stream);
	CU(cudaGetLastError());
	return PSM_OK;
}

int psm_scores_curves(psm_ctx* ctx, int tie_mode, double* out) {
	if (!ctx || !ctx->haveLabels || !out) return fail(ctx, PSM_E_ARG, "scores_curves: no device scores");
	CU(cudaSetDevice(ctx->device));
	return curves_core(ctx, ctx->dLabels.as<uint32_t>(), ctx->dGlob.as<double>(), nullptr, ctx->n, ctx->labelsPos, ctx->labelsNeg, tie_mode, out);
}

int psm_scores_devptr(psm_ctx* ctx, double** dev, uint32_t* n) {
	if (!ctx || !ctx->haveLabels || !dev || !n) return fail(ctx, PSM_E_ARG, "scores_devptr: no device scores");
	*dev = ctx->dGlob.as<double>();
	*n = ctx->n;
	return PSM_OK;
}

int psm_scores_fold_curves(psm_ctx* ctx, const uint32_t* testPosIdx, uint32_t nTP, const uint32_t* testNegIdx, uint32_t nTN, int tie_mode, double* out) {
	const uint64_t N = (uint64_t)nTP + nTN;
	if (!ctx || !ctx->haveLabels || !out || N == 0 || (nTP && !testPosIdx) || (nTN && !testNegIdx)) return fail(ctx, PSM_E_ARG, "scores_fold_curves: bad arguments");
	CU(cudaSetDevice(ctx->device));
	ENSURE(ctx->cIdx, N * 4); ENSURE(ctx->cFoldScores, N * 8); ENSURE(ctx->dFoldLabels, N * 4);
	uint32_t totP = 0, totN = 0;
	for (uint32_t i = 0; i < nTP; i++) { if (testPosIdx[i] >= ctx->n) return fail(ctx, PSM_E_ARG, "scores_fold_curves: index out of range"); const uint32_t l = ctx->hLabels[testPosIdx[i]]; totP += l == 1; totN += l == 0; }
	for (uint32_t i = 0; i < nTN; i++) { if (testNegIdx[i] >= ctx->n) return fail(ctx, PSM_E_ARG, "scores_fold_curves: index out of range"); const uint32_t l = ctx->hLabels[testNegIdx[i]]; totP += l == 1; totN += l == 0; }
	if (nTP) CU(cudaMemcpyAsync(ctx->cIdx.p, testPosIdx, (size_t)nTP * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (nTN) CU(cudaMemcpyAsync(ctx->cIdx.as<uint32_t>() + nTP, testNegIdx, (size_t)nTN * 4, cudaMemcpyHostToDevice, ctx->stream));
	{
		ProfScope ps(ctx, K_GATHER, 2);
		launch_gather_u32(ctx->dLabels.as<uint32_t>(), ctx->cIdx.as<uint32_t>(), N, ctx->dFoldLabels.as<uint32_t>(), ctx->stream);
		launch_gather_f64(ctx->dGlob.as<double>(), ctx->cIdx.as<uint32_t>(), N, ctx->cFoldScores.as<double>(), ctx->stream);   // class1Prob[idx]
	}
	return curves_core(ctx, ctx->dFoldLabels.as<uint32_t>(), ctx->cFoldScores.as<double>(), nullptr, N, totP, totN, tie_mode, out);   // syncs: the index arrays are only borrowed
}

int psm_scores_fold_curves_device(psm_ctx* ctx, uint32_t fold, int tie_mode, double* out) {
	if (!ctx || !ctx->haveLabels || !out || !ctx->fPos || fold + 1 >= ctx->fPosOff.size()) return fail(ctx, PSM_E_ARG, "scores_fold_curves_device: need device scores and fold lists");
	CU(cudaSetDevice(ctx->device));
	const uint32_t nTP = ctx->fPosOff[fold + 1] - ctx->fPosOff[fold], nTN = ctx->fNegOff[fold + 1] - ctx->fNegOff[fold];
	const uint64_t N = (uint64_t)nTP + nTN;
	if (N == 0) return fail(ctx, PSM_E_ARG, "scores_fold_curves_device: empty fold");
	ENSURE(ctx->cIdx, N * 4); ENSURE(ctx->cFoldScores, N * 8); ENSURE(ctx->dFoldLabels, N * 4);
	if (nTP) CU(cudaMemcpyAsync(ctx->cIdx.p, ctx->fPos + ctx->fPosOff[fold], (size_t)nTP * 4, cudaMemcpyDeviceToDevice, ctx->stream));
	if (nTN) CU(cudaMemcpyAsync(ctx->cIdx.as<uint32_t>() + nTP, ctx->fNeg + ctx->fNegOff[fold], (size_t)nTN * 4, cudaMemcpyDeviceToDevice, ctx->stream));
	{
		ProfScope ps(ctx, K_GATHER, 2);
		launch_gather_u32(ctx->dLabels.as<uint32_t>(), ctx->cIdx.as<uint32_t>(), N, ctx->dFoldLabels.as<uint32_t>(), ctx->stream);
		launch_gather_f64(ctx->dGlob.as<double>(), ctx->cIdx.as<uint32_t>(), N, ctx->cFoldScores.as<double>(), ctx->stream);
	}
	return curves_core(ctx, ctx->dFoldLabels.as<uint32_t>(), ctx->cFoldScores.as<double>(), nullptr, N, nTP, nTN, tie_mode, out);
}

int psm_scratch_f64(psm_ctx* ctx, uint32_t count, double** dev) {
	if (!ctx || !dev || count == 0) return fail(ctx, PSM_E_ARG, "scratch_f64: bad arguments");
	CU(cudaSetDevice(ctx->device));
	ENSURE(ctx->cScratch, (size_t)count * 8);
	CU(cudaMemsetAsync(ctx->cScratch.p, 0, (size_t)count * 8, ctx->stream));
	*dev = ctx->cScratch.as<double>();
	return PSM_OK;
}
int psm_scratch_write(psm_ctx* ctx, const double* host, uint32_t count) {
	if (!ctx || !host || (size_t)count * 8 > ctx->cScratch.cap) return fail(ctx, PSM_E_ARG, "scratch_write: bad arguments");
	CU(cudaMemcpyAsync(ctx->cScratch.p, host, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}
int psm_scratch_read(psm_ctx* ctx, double* host, uint32_t count) {
	if (!ctx || !host || (size_t)count * 8 > ctx->cScratch.cap) return fail(ctx, PSM_E_ARG, "scratch_read: bad arguments");
	CU(cudaMemcpyAsync(host, ctx->cScratch.p, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}

// ------------------------------------------------------------------------------------------ cross-rank sum (NCCL)
// The reference's multi-node build gathers every worker's score vectors on the master and sums them there
// (src/hyperSMURF_MPI.cpp:594-646, communicator src/MPIHelper.cpp:29-37).  Here the ranks are GPUs of one NVSwitch domain and
// the sum is one ncclAllReduce on the context's stream.  NCCL is bound at run time: libnccl.so.2 is dlopen'ed at the first
// communicator call (a host process that already loaded a copy -- torch's -- gets that copy back), so the library has no
// link-time dependency and single-GPU hosts never touch it.  Only the few entry points used are declared (ABI of nccl.h 2.x).
namespace {
typedef struct { char internal[PSM_COMM_ID_BYTES]; } NcclId;
struct NcclApi {
	void* lib = nullptr;
	int (*GetVersion)(int*) = nullptr;
	int (*GetUniqueId)(NcclId*) = nullptr;
	int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
	int (*CommDestroy)(void*) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	std::string err;
};
NcclApi* nccl_api() {
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		const char* names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char* nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
		if (!api.lib) { api.err = std::string("NCCL not found: ") + dlerror(); return; }
		auto sym = [&](const char* n) { void* f = dlsym(api.lib, n); if (!f && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + n; return f; };
		api.GetVersion = (int (*)(int*))sym("ncclGetVersion");
		api.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
		api.CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
		api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
		api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
		api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
		api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
	});
	return &api;
}
int nccl_fail(psm_ctx* ctx, NcclApi* a, int rc, const char* where) {
	return fail(ctx, PSM_E_CUDA, std::string(where) + ": " + (a->GetErrorString ? a->GetErrorString(rc) : "NCCL error"));
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h
}  // namespace

int psm_comm_unique_id(void* id128) {
	if (!id128) return PSM_E_ARG;
	NcclApi* a = nccl_api();
	if (!a->err.empty()) return PSM_E_CUDA;
	NcclId id;
	if (a->GetUniqueId(&id) != 0) return PSM_E_CUDA;
	memcpy(id128, &id, sizeof(id));
	return PSM_OK;
}

int psm_comm_init(psm_ctx* ctx, const void* id128, uint32_t rank, uint32_t world) {
	if (!ctx || !id128 || world == 0 || rank >= world) return fail(ctx, PSM_E_ARG, "comm_init: bad arguments");
	NcclApi* a = nccl_api();
	if (!a->err.empty()) return fail(ctx, PSM_E_CUDA, "comm_init: " + a->err);
	if (ctx->comm) { int rc = psm_comm_destroy(ctx); if (rc) return rc; }
	CU(cudaSetDevice(ctx->device));
	NcclId id;
	memcpy(&id, id128, sizeof(id));
	void* comm = nullptr;
	const int rc = a->CommInitRank(&comm, (int)world, id, (int)rank);
	if (rc != 0) return nccl_fail(ctx, a, rc, "comm_init: ncclCommInitRank");
	ctx->comm = comm; ctx->commRank = rank; ctx->commWorld = world;
	return PSM_OK;
}

int psm_comm_info(psm_ctx* ctx, uint32_t* rank, uint32_t* world, int* nccl_version) {
	if (!ctx) return PSM_E_ARG;
	if (rank) *rank = ctx->commRank;
	if (world) *world = ctx->comm ? ctx->commWorld : 0;
	if (nccl_version) { *nccl_version = 0; NcclApi* a = nccl_api(); if (a->err.empty() && a->GetVersion) a->GetVersion(nccl_version); }
	return PSM_OK;
}

int psm_comm_allreduce_sum_f64(psm_ctx* ctx, double* dev, uint64_t count) {
	if (!ctx || !ctx->comm) return fail(ctx, PSM_E_ARG, "comm_allreduce: no communicator (psm_comm_init first)");
	if (count == 0) return PSM_OK;
	if (!dev) return fail(ctx, PSM_E_ARG, "comm_allreduce: null buffer");
	NcclApi* a = nccl_api();
	CU(cudaSetDevice(ctx->device));
	const int rc = a->AllReduce(dev, dev, (size_t)count, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
	if (rc != 0) return nccl_fail(ctx, a, rc, "comm_allreduce: ncclAllReduce");
	CU(cudaStreamSynchronize(ctx->stream));
	return PSM_OK;
}

// The dataset is replicated on every rank, but it does not have to cross every rank's PCIe link: each rank uploads its
// 1/world slice of the rows and one in-place ncclAllGather over NVLink completes every copy (8 ranks pushing 216 MB each through
// the host at once took 9-10 ms per CV at BASELINE configs[1], 150 ms for 3 GB each at configs[2]).  Every rank passes the same n
// and m; a rank only reads ITS rows of x_host, so a host that loads just its slice from disk may pass a pointer offset such
// that row i is at x_host + i * (m + 1).
int psm_dataset_upload_sharded(psm_ctx* ctx, const double* x_host, uint32_t n, uint32_t m) {
	if (!ctx || !x_host || n == 0 || m == 0) return fail(ctx, PSM_E_ARG, "dataset_upload_sharded: bad arguments");
	if (!ctx->comm) return fail(ctx, PSM_E_ARG, "dataset_upload_sharded: no communicator (psm_comm_init first)");
	NcclApi* a = nccl_api();
	CU(cudaSetDevice(ctx->device));
	const size_t ld = (size_t)m + 1;
	const uint32_t W = ctx->commWorld, r = ctx->commRank;
	const size_t chunk = ((size_t)n + W - 1) / W;                          // rows per rank (the last slice may be short)
	ENSURE(ctx->xOwned, chunk * W * ld * sizeof(double));
	double* xd = ctx->xOwned.as<double>();
	const size_t lo = std::min<size_t>((size_t)r * chunk, n), hi = std::min<size_t>(lo + chunk, n);
	if (hi > lo) CU(cudaMemcpyAsync(xd + lo * ld, x_host + lo * ld, (hi - lo) * ld * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	const int rc = a->AllGather(xd + (size_t)r * chunk * ld, xd, chunk * ld, kNcclFloat64, ctx->comm, ctx->stream);
	if (rc != 0) return nccl_fail(ctx, a, rc, "dataset_upload_sharded: ncclAllGather");
	CU(cudaStreamSynchronize(ctx->stream));   // x_host is only borrowed for the call
	ctx->x = xd;
	ctx->n = n; ctx->m = m;
	ctx->foldOpen = ctx->batchOpen = false;
	ctx->haveLabels = false;
	ctx->hFoldIds.clear();
	return PSM_OK;
}

int psm_comm_destroy(psm_ctx* ctx) {
	if (!ctx) return PSM_E_ARG;
	if (!ctx->comm) return PSM_OK;
	NcclApi* a = nccl_api();
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	const int rc = a->CommDestroy(ctx->comm);
	ctx->comm = nullptr; ctx->commWorld = 1; ctx->commRank = 0;
	return rc == 0 ? PSM_OK : nccl_fail(ctx, a, rc, "comm_destroy: ncclCommDestroy");
}

// ------------------------------------------------------------------------------------------ convenience
int psm_batch_capacity(psm_ctx* ctx, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t nTrees, uint32_t mtry, uint32_t p0, uint32_t* capacity) {
	if (!ctx || !ctx->foldOpen || !capacity || nParts == 0 || p0 >= nParts || nTrees == 0) return fail(ctx, PSM_E_ARG, "batch_capacity: bad arguments");
	// per-partition device bytes of one batch (upper bound with Ntr of partition p0; later partitions are never larger)
	PartDesc d0;
	int rc = partition_sizes(ctx->P, ctx->Nneg, nParts, fp, ratio, p0, &d0);
	if (rc) return fail(ctx, rc, "batch_capacity: invalid partition geometry");
	const uint64_t Ntr = d0.Ntr, m = ctx->m, T = nTrees;
	const uint64_t perPart = Ntr * (8 * (m + 1) + 8 * m + 24 * m + 16 * T * m + 4 * T + T + 32 * T + 8) + (uint64_t)ctx->Nte * 16 +
	                         T * (Ntr / 11 + 2) * (2 * sizeof(Item) + mtry * 4 + mtry * sizeof(SearchRes) + sizeof(SplitDec)) + T * 312 * 8 +
	                         (uint64_t)ctx->P * fp * (m + 1) * 4;
	uint64_t nb = std::max<uint64_t>(1, ctx->budget / std::max<uint64_t>(perPart, 1));
	// keep the work-item pool index below 2^31
	const uint64_t itemsPerPart = T * (Ntr / 11 + 2);
	nb = std::min<uint64_t>(nb, std::max<uint64_t>(1, ((1ull << 31) - 1) / std::max<uint64_t>(itemsPerPart, 1)));
	*capacity = (uint32_t)std::min<uint64_t>(nb, nParts - p0);
	return PSM_OK;
}

int psm_fold_run_partitions(psm_ctx* ctx, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t nTrees, uint32_t mtry, uint32_t seed,
                            uint32_t fold, uint32_t p0, uint32_t p1, const int32_t* draws_host) {
	if (!ctx || !ctx->foldOpen) return fail(ctx, PSM_E_ARG, "fold_run_partitions: no open fold");
	if (nParts == 0 || p0 >= p1 || p1 > nParts || nTrees == 0) return fail(ctx, PSM_E_ARG, "fold_run_partitions: bad arguments");
	int rc;
	uint32_t nb = 1;
	if ((rc = psm_batch_capacity(ctx, nParts, fp, ratio, nTrees, mtry, p0, &nb))) return rc;
	const uint64_t perDraw = (uint64_t)ctx->P * fp * (ctx->m + 1);
	for (uint32_t b0 = p0; b0 < p1;) {
		uint32_t b1 = (uint32_t)std::min<uint64_t>(p1, (uint64_t)b0 + nb);
		if ((rc = psm_batch_begin(ctx, nParts, fp, ratio, b0, b1))) return rc;
		if ((rc = psm_batch_assemble(ctx, draws_host ? draws_host + (uint64_t)(b0 - p0) * perDraw : nullptr, (uint64_t)(b1 - b0) * perDraw))) return rc;
		std::vector<uint32_t> seeds(b1 - b0);
		for (uint32_t p = b0; p < b1; p++) seeds[p - b0] = seed + p + fold * nParts;     // hyperSMURF.cpp:273
		if ((rc = psm_batch_train(ctx, nTrees, mtry, seeds.data(), nullptr, nullptr, 0))) return rc;
		if ((rc = psm_batch_predict_accumulate(ctx, nParts))) return rc;
		b0 = b1;
	}
	return PSM_OK;
}

// ------------------------------------------------------------------------------------------ instrumentation
int psm_profile_enable(psm_ctx* ctx, int on) {
	if (!ctx) return PSM_E_ARG;
	prof_resolve(ctx);
	ctx->profOn = on != 0;
	return PSM_OK;
}
int psm_profile_classes(psm_ctx* ctx, uint32_t class_mask) {
	if (!ctx) return PSM_E_ARG;
	prof_resolve(ctx);
	ctx->profMask = class_mask;
	return PSM_OK;
}
int psm_profile_reset(psm_ctx* ctx) {
	if (!ctx) return PSM_E_ARG;
	prof_resolve(ctx);
	for (int i = 0; i < PSM_NUM_KERNEL_CLASSES; i++) { ctx->profMs[i] = 0; ctx->profLaunches[i] = 0; }
	ctx->statSplitBytes = ctx->statNodes = ctx->statLevels = 0;
	return PSM_OK;
}
int psm_profile_get(psm_ctx* ctx, double* ms, uint64_t* launches) {
	if (!ctx) return PSM_E_ARG;
	prof_resolve(ctx);
	for (int i = 0; i < PSM_NUM_KERNEL_CLASSES; i++) { if (ms) ms[i] = ctx->profMs[i]; if (launches) launches[i] = ctx->profLaunches[i]; }
	return PSM_OK;
}
int psm_stats_get(psm_ctx* ctx, uint64_t* split_bytes, uint64_t* nodes_total, uint64_t* levels_total) {
	if (!ctx) return PSM_E_ARG;
	if (split_bytes) *split_bytes = ctx->statSplitBytes;
	if (nodes_total) *nodes_total = ctx->statNodes;
	if (levels_total) *levels_total = ctx->statLevels;
	return PSM_OK;
}

}  // extern "C"
