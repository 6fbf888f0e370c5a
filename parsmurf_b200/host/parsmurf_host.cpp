// parsmurf_b200 host layer -- see parsmurf_host.h.  Every method cites the reference code it mirrors
// (paths relative to the reference repository).  No example data is touched on the CPU here: all of it goes
// through the C ABI (include/parsmurf_b200.h).
#include "parsmurf_host.h"

#include <algorithm>
#include <numeric>
#include <random>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include <atomic>
#include <chrono>
#include <ctime>
#include <exception>

namespace parsmurf {

double g_phase[PH_COUNT] = {0};
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
PhaseScope::PhaseScope(int p) : ph(p), t0(now_s()) {}
// PSM_TIMELINE=1: wall-clock marks of one run (ms since its start), printed by the C entry point -- shows where a rank of
// a multi-GPU run waits
static std::vector<std::pair<std::string, double>> g_marks;
static std::mutex g_markMtx;
static const bool g_trace = getenv("PSM_TIMELINE") != nullptr;
static void trace_mark(const std::string& what) { if (!g_trace) return; std::lock_guard<std::mutex> lk(g_markMtx); g_marks.emplace_back(what, now_s()); }
static std::mutex g_phaseMtx;   // fold lanes account their phases concurrently (the sums then exceed the wall time)
PhaseScope::~PhaseScope() { const double dt = now_s() - t0; std::lock_guard<std::mutex> lk(g_phaseMtx); g_phase[ph] += dt; }

// ------------------------------------------------------------------------------------------ GlibcRand
// glibc stdlib/random_r.c: TYPE_3 (x^31 + x^3 + 1), seeded through the 16807 Park-Miller LCG, first 310 discarded
void GlibcRand::srand(uint32_t seed) {
	if (seed == 0) seed = 1;
	int32_t word = (int32_t)seed;
	tbl[0] = word;
	for (int i = 1; i < 31; i++) {
		long hi = word / 127773, lo = word % 127773;
		word = (int32_t)(16807 * lo - 2836 * hi);
		if (word < 0) word += 2147483647;
		tbl[i] = word;
	}
	f = 3; r = 0;
	for (int i = 0; i < 310; i++) rand();
}
int32_t GlibcRand::rand() {
	uint32_t v = (uint32_t)tbl[f] + (uint32_t)tbl[r];
	tbl[f] = (int32_t)v;
	if (++f >= 31) { f = 0; ++r; }
	else if (++r >= 31) r = 0;
	return (int32_t)(v >> 1);
}
void GlibcRand::fill(int32_t* out, size_t count) {
	for (size_t i = 0; i < count; i++) out[i] = rand();
}
void GlibcRand::window(uint32_t* out) const {
	for (int i = 0; i < 31; i++) out[i] = (uint32_t)tbl[(f + i) % 31];
}
void GlibcRand::discard(uint64_t n) {
	if (n < 64) { for (uint64_t i = 0; i < n; i++) rand(); return; }
	uint32_t w[31];
	window(w);
	psm_glibc_jump(w, n);
	for (int i = 0; i < 31; i++) tbl[i] = (int32_t)w[i];
	f = 0; r = 28;
}
void GlibcRand::shuffle(uint32_t* a, size_t n) {
	// rand() is in [0, 2^31): the modulus is done in 32 bits while i + 1 fits (identical value, much cheaper division)
	size_t i = 1;
	for (; i < n && i < 0x7FFFFFFFull; i++) {
		const uint32_t j = (uint32_t)rand() % (uint32_t)(i + 1);
		if (i != j) std::swap(a[i], a[j]);
	}
	for (; i < n; i++) {
		size_t j = (size_t)(rand() % (int64_t)(i + 1));
		if (i != j) std::swap(a[i], a[j]);
	}
}

// ------------------------------------------------------------------------------------------ index-buffer pool
// The fold / train-test index arrays are a few MB each and rebuilt for every CV.  Fresh allocations of that size come
// straight from mmap, and their first-touch page faults cost more than filling them (measured on the B200 host: 3.2 MB
// filled in 1.6-2 ms fresh vs 0.2 ms recycled), so the library keeps released index vectors and hands them out again.
namespace {
struct IndexPool {
	std::mutex m;
	std::vector<std::vector<uint32_t>> spare;
	static constexpr size_t kMaxSpare = 64;
	std::vector<uint32_t> take(size_t minCap) {
		std::vector<uint32_t> v;
		{
			std::lock_guard<std::mutex> lk(m);
			size_t best = spare.size();
			for (size_t i = 0; i < spare.size(); i++)
				if (spare[i].capacity() >= minCap && (best == spare.size() || spare[i].capacity() < spare[best].capacity())) best = i;
			if (best != spare.size()) { v.swap(spare[best]); spare.erase(spare.begin() + best); }
		}
		v.clear();
		v.reserve(minCap);
		return v;
	}
	void give(std::vector<uint32_t>& v) {
		if (v.capacity() < 16384) return;
		std::lock_guard<std::mutex> lk(m);
		if (spare.size() < kMaxSpare) { spare.emplace_back(); spare.back().swap(v); }
	}
};
IndexPool g_pool;
}  // namespace

// ------------------------------------------------------------------------------------------ Folds (src/folds.cpp:26-117)
Folds::Folds(uint32_t numberOfFolds, uint32_t classSize_, const uint32_t* labels_, std::vector<uint32_t>& ff, bool randomFold_, GlibcRand* rng_)
    : classSize(classSize_), nFolds(numberOfFolds), labels(labels_), fromFile(ff), nPos(numberOfFolds), nNeg(numberOfFolds),
      foldsIdx(numberOfFolds + 1), randomFold(randomFold_), rng(rng_), ids(ff.data()), nIds(ff.size()) {}
Folds::Folds(uint32_t numberOfFolds, uint32_t classSize_, const uint32_t* labels_, const uint32_t* ids_, size_t nIds_, bool randomFold_, GlibcRand* rng_)
    : classSize(classSize_), nFolds(numberOfFolds), labels(labels_), fromFile(noIds), nPos(numberOfFolds), nNeg(numberOfFolds),
      foldsIdx(numberOfFolds + 1), randomFold(randomFold_), rng(rng_), ids(ids_), nIds(ids_ ? nIds_ : 0) {}
Folds::~Folds() { g_pool.give(folds); g_pool.give(posList); g_pool.give(negList); }

void Folds::computeFolds() {
	std::fill(nPos.begin(), nPos.end(), 0);
	std::fill(nNeg.begin(), nNeg.end(), 0);
	auto offsets = [&]() {
		posOff.assign(nFolds + 1, 0); negOff.assign(nFolds + 1, 0);
		for (uint32_t f = 0; f < nFolds; f++) { posOff[f + 1] = posOff[f] + nPos[f]; negOff[f + 1] = negOff[f] + nNeg[f]; }
		totPos = posOff[nFolds]; totNeg = negOff[nFolds];
		posList = g_pool.take(totPos); posList.resize(totPos);
		negList = g_pool.take(totNeg); negList.resize(totNeg);
	};
	if (randomFold || keepFoldsArray) { folds = g_pool.take(classSize); folds.resize(classSize); }
	if (randomFold) {
		std::vector<uint32_t> pos = g_pool.take(classSize), neg = g_pool.take(classSize);
		struct Back { std::vector<uint32_t>& v; ~Back() { g_pool.give(v); } } backPos{pos}, backNeg{neg};
		for (uint32_t i = 0; i < classSize; i++) (labels[i] > 0 ? pos : neg).push_back(i);
		totPos = (uint32_t)pos.size(); totNeg = (uint32_t)neg.size();
		rng->shuffle(pos.data(), pos.size());
		rng->shuffle(neg.data(), neg.size());
		uint32_t ffSize = classSize / nFolds, ffRemn = classSize % nFolds;
		foldsIdx[0] = 0;
		for (uint32_t i = 0; i < nFolds; i++) {
			foldsIdx[i + 1] = foldsIdx[i] + ffSize;
			if (ffRemn != 0) { foldsIdx[i + 1]++; ffRemn--; }
		}
		// round-robin deal: all positives first, then the negatives (src/folds.cpp:52-63)
		for (uint32_t i = 0; i < totPos; i++) { folds[foldsIdx[i % nFolds] + i / nFolds] = pos[i]; nPos[i % nFolds]++; }
		for (uint32_t i = totPos; i < totPos + totNeg; i++) { folds[foldsIdx[i % nFolds] + i / nFolds] = neg[i - totPos]; nNeg[i % nFolds]++; }
		offsets();
		for (uint32_t f = 0; f < nFolds; f++) {                              // fold f = its positives, then its negatives, in deal order
			std::copy(folds.begin() + foldsIdx[f], folds.begin() + foldsIdx[f] + nPos[f], posList.begin() + posOff[f]);
			std::copy(folds.begin() + foldsIdx[f] + nPos[f], folds.begin() + foldsIdx[f + 1], negList.begin() + negOff[f]);
		}
	} else {
		if (&fromFile != &noIds) { ids = fromFile.data(); nIds = fromFile.size(); }   // the caller may have filled the vector after construction
		if (nIds < classSize) throw Error(PSM_E_ARG, "Folds: fold assignment shorter than the label vector");
		const bool fillFolds = keepFoldsArray;
		// two passes (count, then place) over contiguous slices of the examples, a few host threads on large inputs: slice t
		// writes after everything slices < t write, so every fold keeps ascending example order (src/folds.cpp:93-110)
		const uint32_t hw = std::max(1u, std::thread::hardware_concurrency());
		const uint32_t nT = classSize >= (1u << 18) ? std::min<uint32_t>(std::min<uint32_t>(std::max<uint32_t>(4u, classSize >> 20), 16u), std::max(4u, hw / 2)) : 1u;
		std::vector<std::vector<uint32_t>> hP(nT, std::vector<uint32_t>(nFolds, 0)), hN(nT, std::vector<uint32_t>(nFolds, 0));
		std::atomic<bool> bad(false);
		auto slice = [&](uint32_t t, uint32_t* a, uint32_t* b) { *a = (uint32_t)((uint64_t)classSize * t / nT); *b = (uint32_t)((uint64_t)classSize * (t + 1) / nT); };
		auto run = [&](const std::function<void(uint32_t)>& fn) {
			if (nT == 1) { fn(0); return; }
			std::vector<std::thread> th;
			for (uint32_t t = 0; t < nT; t++) th.emplace_back(fn, t);
			for (auto& x : th) x.join();
		};
		run([&](uint32_t t) {
			uint32_t a, b; slice(t, &a, &b);
			for (uint32_t i = a; i < b; i++) {
				const uint32_t v = ids[i];
				if (v >= nFolds) { bad = true; return; }
				if (labels[i] > 0) hP[t][v]++; else hN[t][v]++;
			}
		});
		if (bad) throw Error(PSM_E_ARG, "Folds: fold id out of range");
		foldsIdx[0] = 0;
		for (uint32_t f = 0; f < nFolds; f++) {
			for (uint32_t t = 0; t < nT; t++) { nPos[f] += hP[t][f]; nNeg[f] += hN[t][f]; }
			foldsIdx[f + 1] = foldsIdx[f] + nPos[f] + nNeg[f];
		}
		offsets();
		run([&](uint32_t t) {
			std::vector<uint32_t> cf(nFolds), cp(nFolds), cn(nFolds);
			for (uint32_t f = 0; f < nFolds; f++) {
				uint32_t bp = 0, bn = 0;
				for (uint32_t q = 0; q < t; q++) { bp += hP[q][f]; bn += hN[q][f]; }
				cp[f] = posOff[f] + bp; cn[f] = negOff[f] + bn; cf[f] = foldsIdx[f] + bp + bn;
			}
			uint32_t a, b; slice(t, &a, &b);
			for (uint32_t i = a; i < b; i++) {
				const uint32_t bin = ids[i];
				if (fillFolds) folds[cf[bin]++] = i;
				if (labels[i] > 0) posList[cp[bin]++] = i; else negList[cn[bin]++] = i;
			}
		});
	}
	maxPos = *std::max_element(nPos.begin(), nPos.end());
	maxNeg = *std::max_element(nNeg.begin(), nNeg.end());
	hostListsValid = true; devBuilt = nullptr;
}

void Folds::computeFoldsOnDevice(Device* dev) {
	if (randomFold) throw Error(PSM_E_ARG, "Folds::computeFoldsOnDevice: random folds are generated on the host");
	if (&fromFile != &noIds) { ids = fromFile.data(); nIds = fromFile.size(); }
	if (nIds < classSize) throw Error(PSM_E_ARG, "Folds: fold assignment shorter than the label vector");
	dev->check(psm_folds_build_device(dev->ctx, ids, nFolds, nPos.data(), nNeg.data()));
	posOff.assign(nFolds + 1, 0); negOff.assign(nFolds + 1, 0);
	foldsIdx[0] = 0;
	for (uint32_t f = 0; f < nFolds; f++) {
		posOff[f + 1] = posOff[f] + nPos[f]; negOff[f + 1] = negOff[f] + nNeg[f];
		foldsIdx[f + 1] = foldsIdx[f] + nPos[f] + nNeg[f];
	}
	totPos = posOff[nFolds]; totNeg = negOff[nFolds];
	maxPos = *std::max_element(nPos.begin(), nPos.end());
	maxNeg = *std::max_element(nNeg.begin(), nNeg.end());
	devBuilt = dev; hostListsValid = false;
}

void Folds::hostLists() const {
	std::lock_guard<std::mutex> lk(hostListsMtx);
	if (hostListsValid) return;
	if (!devBuilt) throw Error(PSM_E_ARG, "Folds: computeFolds has not run");
	Folds* self = const_cast<Folds*>(this);
	self->posList = g_pool.take(totPos); self->posList.resize(totPos);
	self->negList = g_pool.take(totNeg); self->negList.resize(totNeg);
	devBuilt->check(psm_folds_download(devBuilt->ctx, self->posList.data(), self->negList.data()));
	hostListsValid = true;
}

// ------------------------------------------------------------------------------------------ TestTrainDivider
static void ttd_blk(std::vector<uint32_t>& dst, const std::vector<uint32_t>& src, const std::vector<uint32_t>& off, uint32_t f) {
	dst.insert(dst.end(), src.begin() + off[f], src.begin() + off[f + 1]);
}
void TestTrainDivider::buildTest(uint32_t i, uint32_t cur) {
	const Folds* F = folds;
	F->hostLists();
	testPosIdx[i] = g_pool.take(F->nPos[cur]); testNegIdx[i] = g_pool.take(F->nNeg[cur]);
	ttd_blk(testPosIdx[i], F->posList, F->posOff, cur); ttd_blk(testNegIdx[i], F->negList, F->negOff, cur);
}
void TestTrainDivider::buildTrain(uint32_t i, uint32_t cur, int64_t jump) {
	const Folds* F = folds;
	F->hostLists();
	const uint32_t jumpPos = jump >= 0 ? F->nPos[jump] : 0, jumpNeg = jump >= 0 ? F->nNeg[jump] : 0;
	trngPosIdx[i] = g_pool.take(F->totPos - F->nPos[cur] - jumpPos); trngNegIdx[i] = g_pool.take(F->totNeg - F->nNeg[cur] - jumpNeg);
	for (uint32_t kf = 0; kf < F->nFolds; kf++) {                            // ascending fold order, as src/testtraindivider.cpp:57-80
		if (kf == cur || (int64_t)kf == jump) continue;
		ttd_blk(trngPosIdx[i], F->posList, F->posOff, kf); ttd_blk(trngNegIdx[i], F->negList, F->negOff, kf);
	}
}

// src/testtraindivider.cpp:5-106
TestTrainDivider::TestTrainDivider(const Folds* folds_, uint8_t wmode_, GlibcRand* rng) : folds(folds_), wmode(wmode_), nFolds(folds_->nFolds), lazy(true), start(*rng) {
	testPosIdx.resize(nFolds); testNegIdx.resize(nFolds); trngPosIdx.resize(nFolds); trngNegIdx.resize(nFolds);
	testState.assign(nFolds, 0); trainState.assign(nFolds, 0);
	// Fold i's std::random_shuffle (:90) consumes trngNeg_i - 1 rand() calls, so its place in the one rand() stream is known
	// up front; with the generator's jump-ahead the folds are built independently of each other (concurrently, or not at
	// all) and still shuffle exactly like the sequential reference.
	shuffleOff.assign(nFolds + 1, 0);
	for (uint32_t i = 0; i < nFolds; i++) {
		// exec.mode "train": the reference overrides trngNegNum[0] with testNegNum[0] BEFORE the shuffle (:33-36,90), so fold 0's
		// std::random_shuffle runs over that many (still unfilled) slots and burns their rand() draws although the arrays are
		// overwritten with the unshuffled fold afterwards (:101-104)
		uint64_t nneg = (wmode == MODE_TRAIN && i == 0) ? folds->nNeg[0] : folds->totNeg - folds->nNeg[i];
		shuffleOff[i + 1] = shuffleOff[i] + (nneg > 0 ? nneg - 1 : 0);
	}
	for (uint32_t i = 0; i < nFolds; i++) {                                  // counts are known without building anything
		testPosNum.push_back(folds->nPos[i]); testNegNum.push_back(folds->nNeg[i]);
		trngPosNum.push_back(folds->totPos - folds->nPos[i]); trngNegNum.push_back(folds->totNeg - folds->nNeg[i]);
	}
	if (wmode == MODE_TRAIN) { trngPosNum[0] = testPosNum[0]; trngNegNum[0] = testNegNum[0]; }   // :33-36
	rng->discard(shuffleOff[nFolds]);
}

TestTrainDivider::~TestTrainDivider() {
	if (worker.joinable()) worker.join();
	for (auto* vv : {&testPosIdx, &testNegIdx, &trngPosIdx, &trngNegIdx})
		for (auto& v : *vv) g_pool.give(v);
}

void TestTrainDivider::shuffleWindow(uint32_t fold, uint32_t* w31) const {
	GlibcRand local = start;
	local.discard(shuffleOff.empty() ? 0 : shuffleOff[fold]);
	local.window(w31);
}

void TestTrainDivider::prefetch(std::vector<uint32_t> order) {
	if (!lazy || worker.joinable()) return;
	worker = std::thread([this, order]() { for (uint32_t f : order) if (f < nFolds) ensure(f); });
}

void TestTrainDivider::ensureTest(uint32_t fold) const {
	if (!lazy) return;                                                      // inner-CV constructor builds synchronously
	std::unique_lock<std::mutex> lk(mtx);
	if (testState[fold] == 0) {
		testState[fold] = 1;
		lk.unlock();
		const_cast<TestTrainDivider*>(this)->buildTest(fold, fold);
		lk.lock();
		testState[fold] = 2;
		cv.notify_all();
	} else {
		cv.wait(lk, [&] { return testState[fold] == 2; });
	}
}

void TestTrainDivider::ensure(uint32_t fold) const {
	if (!lazy) return;
	ensureTest(fold);
	std::unique_lock<std::mutex> lk(mtx);
	if (trainState[fold] == 0) {
		trainState[fold] = 1;
		lk.unlock();
		TestTrainDivider* self = const_cast<TestTrainDivider*>(this);
		if (wmode == MODE_TRAIN && fold == 0) {                              // :101-104 training on everything: the "training" arrays are the fold itself
			self->trngPosIdx[0] = testPosIdx[0]; self->trngNegIdx[0] = testNegIdx[0];
		} else {
			self->buildTrain(fold, fold, -1);
			GlibcRand local = start;
			local.discard(shuffleOff[fold]);
			local.shuffle(self->trngNegIdx[fold].data(), self->trngNegIdx[fold].size());   // :90
		}
		lk.lock();
		trainState[fold] = 2;
		cv.notify_all();
	} else {
		cv.wait(lk, [&] { return trainState[fold] == 2; });
	}
}

// src/testtraindivider.cpp:126-236 -- inner CV: fold `foldToJump` is left out entirely, no shuffle
TestTrainDivider::TestTrainDivider(const Folds* folds_, uint8_t wmode_, uint32_t foldToJump) : folds(folds_), wmode(wmode_), nFolds(folds_->nFolds - 1) {
	testPosIdx.resize(nFolds); testNegIdx.resize(nFolds); trngPosIdx.resize(nFolds); trngNegIdx.resize(nFolds);
	uint32_t cur = 0;
	for (uint32_t i = 0; i < nFolds; i++) {
		if (cur == foldToJump) cur++;
		buildTest(i, cur);
		buildTrain(i, cur, foldToJump);
		cur++;
	}
	for (uint32_t i = 0; i < nFolds; i++) {
		testPosNum.push_back(testPosIdx[i].size()); testNegNum.push_back(testNegIdx[i].size());
		trngPosNum.push_back(trngPosIdx[i].size()); trngNegNum.push_back(trngNegIdx[i].size());
	}
}

// ------------------------------------------------------------------------------------------ Partition (src/partition.cpp:7-37)
Partition::Partition(const Folds* folds_, const TestTrainDivider* ttd_, uint32_t nPart_, uint32_t fp_, uint32_t ratio_, uint32_t m_, uint32_t n_)
    : folds(folds_), ttd(ttd_), nPart(nPart_), fp(fp_), ratio(ratio_), m(m_), n(n_) {
	maxSize = ratio == 0 ? ((fp + 1) * folds->maxPos + folds->maxNeg) * (folds->nFolds - 1) : (fp + 1) * (ratio + 1) * folds->maxPos * (folds->nFolds - 1);
	if (folds->nFolds == 1) maxSize = ratio == 0 ? ((fp + 1) * folds->maxPos + folds->maxNeg) : (fp + 1) * (ratio + 1) * folds->maxPos;
}
void Partition::setFold(uint32_t f) {
	currentFold = f;
	ttd->ensure(f);
	trngPosNum = ttd->trngPosNum[f]; trngNegNum = ttd->trngNegNum[f]; testPosNum = ttd->testPosNum[f]; testNegNum = ttd->testNegNum[f];
	trngPosIdx = ttd->trngPosIdx[f].data(); trngNegIdx = ttd->trngNegIdx[f].data();
	testPosIdx = ttd->testPosIdx[f].data(); testNegIdx = ttd->testNegIdx[f].data();
}

void Partition::setFoldCounts(uint32_t f) {
	currentFold = f;
	trngPosNum = ttd->trngPosNum[f]; trngNegNum = ttd->trngNegNum[f]; testPosNum = ttd->testPosNum[f]; testNegNum = ttd->testNegNum[f];
	trngPosIdx = trngNegIdx = testPosIdx = testNegIdx = nullptr;
}

// ------------------------------------------------------------------------------------------ Device
Device::Device(int device, void* cudaStream) : deviceId(device) {
	int rc = psm_ctx_create(device, &ctx);
	if (rc) throw Error(rc, "parsmurf_b200: no usable CUDA device (there is no CPU fallback)");
	if (cudaStream) check(psm_set_stream(ctx, cudaStream));
}
Device::~Device() { if (ctx) psm_ctx_destroy(ctx); }
void Device::check(int rc) const { if (rc) throw Error(rc, psm_last_error(ctx)); }
void Device::commUniqueId(void* id128) { if (psm_comm_unique_id(id128)) throw Error(PSM_E_CUDA, "NCCL is not available (libnccl.so.2 could not be loaded)"); }
void Device::commInit(const void* id128, uint32_t rank, uint32_t world) { check(psm_comm_init(ctx, id128, rank, world)); }
bool Device::hasComm() const { uint32_t w = 0; return psm_comm_info(ctx, nullptr, &w, nullptr) == PSM_OK && w > 0; }
void Device::allReduce(double* dev_ptr, uint64_t count) { check(psm_comm_allreduce_sum_f64(ctx, dev_ptr, count)); }
void Device::upload(const double* x, uint32_t n_, uint32_t m_) {
	// with a communicator the ranks share the upload: a slice each over PCIe, the rest over NVLink (psm_dataset_upload_sharded)
	if (shardedUpload && hasComm()) check(psm_dataset_upload_sharded(ctx, x, n_, m_));
	else check(psm_dataset_upload(ctx, x, n_, m_));
	n = n_; m = m_; foldTag = nullptr;
}
void Device::attachDevice(const double* x, uint32_t n_, uint32_t m_) { check(psm_dataset_attach_device(ctx, x, n_, m_)); n = n_; m = m_; foldTag = nullptr; }
void Device::setLaneCount(uint32_t lanes) {
	while (laneCount() < lanes) {
		std::unique_ptr<Device> d(new Device(deviceId, nullptr));
		d->check(psm_ctx_own_stream(d->ctx));
		extra.push_back(std::move(d));
	}
}
void Device::ensureFold(const Partition* part) {
	if (foldTag == (const void*)part->ttd && foldId == part->currentFold) return;
	check(psm_fold_begin(ctx, part->trngPosIdx, part->trngPosNum, part->trngNegIdx, part->trngNegNum, part->testPosIdx, part->testPosNum,
	                     part->testNegIdx, part->testNegNum));
	foldTag = part->ttd; foldId = part->currentFold; knnK = 0;
}
void Device::uploadFolds(const Folds* F) {
	if (F->builtOn() == this) { foldTag = nullptr; return; }                 // the lists were built here (Folds::computeFoldsOnDevice)
	F->hostLists();
	check(psm_folds_upload(ctx, F->nFolds, F->posList.data(), F->posOff.data(), F->negList.data(), F->negOff.data()));
	foldTag = nullptr;
}
void Device::attachFolds(const Device* src) { check(psm_folds_attach(ctx, src->ctx)); foldTag = nullptr; }
void Device::beginFoldOnDevice(const Partition* part) {
	uint32_t w[31];
	part->ttd->shuffleWindow(part->currentFold, w);
	check(psm_fold_begin_device(ctx, part->currentFold, w));
	foldTag = part->ttd; foldId = part->currentFold; knnK = 0;
}
void Device::ensureKnn(uint32_t k) {
	if (knnK == k) return;
	check(psm_fold_knn(ctx, k));
	knnK = k;
}

// ------------------------------------------------------------------------------------------ SamplerKNN
SamplerKNN::SamplerKNN(const Partition* part_, uint8_t wmode_, Device* dev_, uint32_t m_, uint32_t n_, uint32_t k_, GlibcRand* rng_)
    : part(part_), dev(dev_), m(m_), n(n_), wmode(wmode_), k(k_), rng(rng_) {
	testSize = part->testPosNum + part->testNegNum;                         // src/sampler.cpp:11
}
void SamplerKNN::setPartitionRange(uint32_t a, uint32_t b) {
	if (a >= b || b > part->nPart) throw Error(PSM_E_ARG, "SamplerKNN::setPartitionRange: bad range");
	p0 = a; p1 = b; currentPartition = a; currentFold = part->currentFold;
	posForOverSmpl = part->trngPosNum;                                       // src/sampler.cpp:67
	uint32_t negInEach = (uint32_t)std::ceil(part->trngNegNum / (double)part->nPart);
	negForUndrSmpl = (a != part->nPart - 1) ? negInEach : part->trngNegNum - negInEach * (part->nPart - 1);
}
void SamplerKNN::sample() {
	dev->ensureFold(part);
	std::vector<int32_t> draws;
	const uint64_t ndraws = (uint64_t)(p1 - p0) * posForOverSmpl * part->fp * (m + 1);   // src/samplerKNN.cpp:95,99 in partition order
	if (part->fp > 0) {
		dev->ensureKnn(k);                                                   // src/samplerKNN.cpp:60-93, once per fold
		if (!deviceDraws) {
			PhaseScope ps(PH_DRAWS);
			draws.resize(ndraws);
			rng->fill(draws.data(), draws.size());
		}
	}
	PhaseScope ps(PH_ASSEMBLE);
	dev->check(psm_batch_begin(dev->ctx, part->nPart, part->fp, part->ratio, p0, p1));
	if (deviceDraws && ndraws) {
		uint32_t w[31];
		rng->window(w);
		dev->check(psm_batch_assemble_glibc(dev->ctx, w));                   // the same stream, replayed on the GPU
		rng->discard(ndraws);
	} else {
		dev->check(psm_batch_assemble(dev->ctx, draws.empty() ? nullptr : draws.data(), draws.size()));
	}
	uint32_t sz[3];
	dev->check(psm_batch_get_training(dev->ctx, p0, sz, nullptr));
	trngPos = sz[0]; trngNeg = sz[1]; lineLen = sz[2]; trngSize = trngPos + trngNeg;   // src/samplerKNN.cpp:8-13
}
std::vector<double> SamplerKNN::trngData(uint32_t p) {
	uint32_t sz[3];
	dev->check(psm_batch_get_training(dev->ctx, p, sz, nullptr));
	std::vector<double> out((size_t)sz[2] * (m + 1));
	dev->check(psm_batch_get_training(dev->ctx, p, sz, out.data()));
	return out;
}

// ------------------------------------------------------------------------------------------ rfRanger
rfRanger::rfRanger(Device* dev_, uint32_t /*m*/, bool prediction_mode_, const SamplerKNN* samp_, uint32_t numTrees_, uint32_t mtry_, uint32_t rfThrd, uint32_t seedBase)
    : dev(dev_), samp(samp_), mtry(mtry_), rfSeed(seedBase), num_threads(rfThrd), numTrees(numTrees_), prediction_mode(prediction_mode_) {}
void rfRanger::train(bool) {
	if (!samp) throw Error(PSM_E_ARG, "rfRanger::train: this forest was loaded from a file");
	std::vector<uint32_t> seeds(samp->p1 - samp->p0);
	for (uint32_t p = samp->p0; p < samp->p1; p++) seeds[p - samp->p0] = rfSeed + p;
	dev->check(psm_batch_train(dev->ctx, numTrees, mtry, seeds.data(), nullptr, nullptr, 0));
}
void rfRanger::predict(bool) { dev->check(psm_batch_predict_accumulate(dev->ctx, samp ? samp->part->nPart : nPartTotal_)); }
rfRanger::rfRanger(Device* dev_, const std::string& forestDirname, uint32_t p0, uint32_t p1, uint32_t nPartTotal, uint32_t m, uint32_t numTrees_,
                   uint32_t mtry_, uint32_t rfThrd, uint32_t seed)
    : dev(dev_), samp(nullptr), mtry(mtry_), rfSeed(seed), num_threads(rfThrd), numTrees(numTrees_), prediction_mode(true), m_(m), nPartTotal_(nPartTotal) {
	if (p1 <= p0) throw Error(PSM_E_ARG, "rfRanger: empty partition range");
	std::vector<uint64_t> nNodes, left, right, var;
	std::vector<double> value, frac;
	uint64_t T = 0;
	for (uint32_t p = p0; p < p1; p++) {
		const std::string fn = forestDirname + "/" + std::to_string(p) + ".out.forest";   // src/hyperSMURF.cpp:421
		ForestFile ff = ForestFile::read(fn);
		// the file knows its own num_trees (Forest::loadFromFile); one batch must be uniform
		if (p == p0) T = ff.trees.size();
		else if (ff.trees.size() != T) throw Error(PSM_E_ARG, "rfRanger: " + fn + " holds a different number of trees than the first forest of the batch");
		if (ff.numVariables != (uint64_t)m + 1 || ff.dependentVarID != m) throw Error(PSM_E_ARG, "rfRanger: " + fn + " was grown on data with a different number of features");
		if (ff.classValues.size() != 2 || ff.classValues[0] != 1.0 || ff.classValues[1] != 2.0)
			throw Error(PSM_E_ARG, "rfRanger: " + fn + " does not have class values {1, 2} (positive first)");
		for (const TreeExport& t : ff.trees) {
			nNodes.push_back(t.left.size());
			left.insert(left.end(), t.left.begin(), t.left.end()); right.insert(right.end(), t.right.begin(), t.right.end());
			var.insert(var.end(), t.splitVarIDs.begin(), t.splitVarIDs.end());
			value.insert(value.end(), t.splitValues.begin(), t.splitValues.end());
			frac.insert(frac.end(), t.terminalClassCounts.begin(), t.terminalClassCounts.end());
		}
	}
	if (T == 0) throw Error(PSM_E_ARG, "rfRanger: forest file without trees");
	numTrees = (uint32_t)T;
	dev->check(psm_batch_import_forests(dev->ctx, p0, p1, numTrees, nNodes.data(), left.data(), right.data(), var.data(), value.data(), frac.data()));
}
void rfRanger::saveForest(uint32_t nPart, const std::string& forestDirname) const {
	if (!samp) throw Error(PSM_E_ARG, "rfRanger::saveForest: this forest was loaded from a file");
	ForestFile ff;
	ff.dependentVarID = samp->m; ff.numVariables = (uint64_t)samp->m + 1;    // label column "Labels" is the last one (src/hyperSMURF.cpp:283-291)
	ff.classValues = {1.0, 2.0};                                            // first-appearance order: the training matrix starts with the positives
	for (uint32_t t = 0; t < numTrees; t++) ff.trees.push_back(getTree(nPart, t));
	ff.write(forestDirname + "/" + std::to_string(nPart) + ".out.forest");   // src/rfRanger.cpp:199
}

// ------------------------------------------------------------------------------------------ ForestFile
namespace {
void put64(std::ofstream& f, uint64_t v) { f.write((const char*)&v, 8); }
template <typename T> void putVec(std::ofstream& f, const std::vector<T>& v) { put64(f, v.size()); f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T))); }
uint64_t get64(std::ifstream& f, const std::string& path) {
	uint64_t v = 0;
	f.read((char*)&v, 8);
	if (!f) throw Error(PSM_E_ARG, "forest file truncated: " + path);
	return v;
}
template <typename T> void getVec(std::ifstream& f, std::vector<T>& v, const std::string& path) {
	const uint64_t len = get64(f, path);
	if (len > (1ull << 32)) throw Error(PSM_E_ARG, "forest file corrupt (vector length): " + path);
	v.resize(len);
	f.read((char*)v.data(), (std::streamsize)(len * sizeof(T)));
	if (!f) throw Error(PSM_E_ARG, "forest file truncated: " + path);
}
}  // namespace
void ForestFile::write(const std::string& path) const {
	std::ofstream f(path.c_str(), std::ios::binary);
	if (!f.good()) throw Error(PSM_E_ARG, "Could not write to output file: " + path + ".");   // Forest.cpp:396-398
	put64(f, dependentVarID);
	put64(f, trees.size());
	put64(f, numVariables);                                                  // is_ordered_variable: one byte per column, all ordered
	for (uint64_t i = 0; i < numVariables; i++) { const char one = 1; f.write(&one, 1); }
	put64(f, numVariables);                                                  // ForestProbability::saveToFileInternal
	const uint32_t treetype = 9;                                             // TREE_PROBABILITY (src/ranger/globals.h)
	f.write((const char*)&treetype, 4);
	putVec(f, classValues);
	for (const rfRanger::TreeExport& t : trees) {
		put64(f, 2);                                                         // child_nodeIDs: [left[], right[]]
		putVec(f, t.left); putVec(f, t.right);
		putVec(f, t.splitVarIDs);
		putVec(f, t.splitValues);
		std::vector<uint64_t> terminal;                                     // TreeProbability::appendToFileInternal
		for (uint64_t i = 0; i < t.left.size(); i++) if (t.left[i] == 0 && t.right[i] == 0) terminal.push_back(i);
		putVec(f, terminal);
		put64(f, terminal.size());
		for (uint64_t i : terminal) { put64(f, 2); f.write((const char*)&t.terminalClassCounts[2 * i], 16); }
	}
	if (!f.good()) throw Error(PSM_E_ARG, "write error on " + path);
}
ForestFile ForestFile::read(const std::string& path) {
	std::ifstream f(path.c_str(), std::ios::binary);
	if (!f.good()) throw Error(PSM_E_ARG, "Could not read from input file: " + path + ".");   // Forest.cpp loadFromFile
	ForestFile ff;
	ff.dependentVarID = get64(f, path);
	const uint64_t nTrees = get64(f, path);
	const uint64_t nOrdered = get64(f, path);
	if (nTrees > (1ull << 24) || nOrdered > (1ull << 24)) throw Error(PSM_E_ARG, "forest file corrupt (header): " + path);
	f.ignore((std::streamsize)nOrdered);
	ff.numVariables = get64(f, path);
	uint32_t treetype = 0;
	f.read((char*)&treetype, 4);
	if (!f || treetype != 9) throw Error(PSM_E_ARG, "Wrong treetype. Loaded file is not a probability estimation forest.");   // ForestProbability.cpp:284-286
	getVec(f, ff.classValues, path);
	ff.trees.resize(nTrees);
	for (rfRanger::TreeExport& t : ff.trees) {
		if (get64(f, path) != 2) throw Error(PSM_E_ARG, "forest file corrupt (child_nodeIDs): " + path);
		getVec(f, t.left, path); getVec(f, t.right, path);
		getVec(f, t.splitVarIDs, path);
		getVec(f, t.splitValues, path);
		const size_t N = t.left.size();
		if (t.right.size() != N || t.splitVarIDs.size() != N || t.splitValues.size() != N) throw Error(PSM_E_ARG, "forest file corrupt (tree vectors): " + path);
		std::vector<uint64_t> terminal;
		getVec(f, terminal, path);
		if (get64(f, path) != terminal.size()) throw Error(PSM_E_ARG, "forest file corrupt (terminal nodes): " + path);
		t.terminalClassCounts.assign(2 * N, std::nan(""));
		for (uint64_t id : terminal) {
			std::vector<double> c;
			getVec(f, c, path);
			if (id >= N || c.empty() || c.size() > 2) throw Error(PSM_E_ARG, "forest file corrupt (class counts): " + path);
			t.terminalClassCounts[2 * id] = c[0];
			t.terminalClassCounts[2 * id + 1] = c.size() > 1 ? c[1] : 0.0;
		}
	}
	return ff;
}

rfRanger::TreeExport rfRanger::getTree(uint32_t p, uint32_t t) const {
	uint64_t nn = 0;
	dev->check(psm_batch_tree_nodes(dev->ctx, p, t, &nn));
	TreeExport e;
	e.left.resize(nn); e.right.resize(nn); e.splitVarIDs.resize(nn); e.splitValues.resize(nn); e.terminalClassCounts.resize(2 * nn);
	dev->check(psm_batch_export_tree(dev->ctx, p, t, e.left.data(), e.right.data(), e.splitVarIDs.data(), e.splitValues.data(), e.terminalClassCounts.data()));
	return e;
}

// ------------------------------------------------------------------------------------------ simulated data
void generateRandomSet(uint32_t nn, uint32_t mm, std::vector<double>& xx, std::vector<uint32_t>& yy, double prob, uint32_t seed) {
	std::default_random_engine gen(seed);
	std::normal_distribution<> neg(0, 1), pos(5, 1);
	std::bernoulli_distribution lab(prob);
	yy.resize(nn); xx.resize((size_t)nn * (mm + 1));
	for (auto& v : yy) v = lab(gen);                                         // all the labels first, then the rows in order
	for (uint32_t i = 0; i < nn; i++) {
		double* row = xx.data() + (size_t)i * (mm + 1);
		for (uint32_t j = 0; j < mm; j++) row[j] = yy[i] > 0 ? pos(gen) : neg(gen);
		row[mm] = yy[i] > 0 ? 1.0 : 2.0;
	}
}

// ------------------------------------------------------------------------------------------ Curves
Curves::Curves(Device* dev_, const std::vector<uint32_t>& labels, const double* predictions, int tieMode_)
    : dev(dev_), labls(labels), preds(predictions), tieMode(tieMode_) {}
void Curves::eval() {
	if (done) return;
	dev->check(psm_curves(dev->ctx, labls.data(), preds, labls.size(), tieMode, res));
	done = true;
}
double Curves::evalAUROC_ok() { eval(); return res[0]; }
double Curves::evalAUPRC() { eval(); return res[1]; }

// ------------------------------------------------------------------------------------------ hyperSMURF
hyperSMURF::hyperSMURF(const CommonParams* cp, std::vector<GridParams>& gp, Folds* fm, Device* dev_, const std::vector<uint32_t>& y_,
                       double* c1, double* c2, GlibcRand* rng_)
    : commonParams(cp), gridParams(gp), foldManager(fm), dev(dev_), y(y_), class1Prob(c1), class2Prob(c2), rng(rng_) {}
hyperSMURF::~hyperSMURF() {}

void hyperSMURF::setSharding(uint32_t rank_, uint32_t world_, std::function<void(double*, uint64_t)> reduce_) {
	rank = rank_; world = world_ ? world_ : 1; reduce = std::move(reduce_);
}

void hyperSMURF::noteCurves(Device* d) {
	uint64_t mixed = 0; int host = 0; double sp[2] = {0, 0};
	if (psm_curves_info(d->ctx, &mixed, &host, sp) == PSM_OK) {
		curvesMixedPairs += mixed; curvesHostSorts += (uint64_t)host;
		std::lock_guard<std::mutex> lk(spreadMtx);
		curvesMaxSpread = std::max(curvesMaxSpread, std::max(std::fabs(sp[0]), std::fabs(sp[1])));
	}
}

void hyperSMURF::initStandard() {                                           // src/hyperSMURF.cpp:22-51
	nn = commonParams->nn; mm = commonParams->mm; nFolds = commonParams->nFolds;
	setGridParams(0);
	seed = commonParams->seed; verboseLevel = commonParams->verboseLevel;
	wmode = commonParams->wmode; woptimiz = commonParams->woptimiz; inInternalCV = false;
	{ PhaseScope ps(PH_FOLDS_TTD); ttd.reset(new TestTrainDivider(foldManager, wmode, rng)); }
	part.reset(new Partition(foldManager, ttd.get(), nPart, fp, ratio, mm, nn));
	// start building the folds this rank will train on, in the order it will need them (overlaps the upload of x)
	uint32_t f0 = 0, f1 = nFolds;
	if (commonParams->minFold != (uint32_t)-1) f0 = commonParams->minFold;
	if (commonParams->maxFold != (uint32_t)-1) f1 = std::min(commonParams->maxFold, nFolds);
	std::vector<uint32_t> order;
	for (uint32_t f = f0; f < f1; f++) {
		uint32_t a = 0, b = 0;
		partRange(willShardUnits(), f1 - f0, nPart, world, rank, f - f0, &a, &b);
		if (b > a || !willShardUnits()) order.push_back(f);
	}
	if (!foldsOnDevice()) ttd->prefetch(order);
}
bool hyperSMURF::foldsOnDevice() const {
	return deviceFolds && commonParams->wmode == MODE_CV && !inInternalCV && ttd && ttd->builtOnDemand();
}
bool hyperSMURF::willShardUnits() const {
	return shardUnits && world > 1 && (bool)reduce && commonParams->wmode == MODE_CV && commonParams->woptimiz == OPT_NO && !inInternalCV && !commonParams->customCV;
}
void hyperSMURF::initInternalCV(uint32_t jump) {                            // src/hyperSMURF.cpp:53-78
	nn = commonParams->nn; mm = commonParams->mm; nFolds = commonParams->nFolds - 1;
	seed = commonParams->seed; verboseLevel = commonParams->verboseLevel;
	wmode = MODE_CV; woptimiz = OPT_NO; inInternalCV = true; foldToJump = jump;
	ttd.reset(new TestTrainDivider(foldManager, wmode, jump));
}
void hyperSMURF::setGridParams(uint32_t idx) {                              // src/hyperSMURF.cpp:80-88
	idxInGridParams = idx;
	nPart = gridParams[idx].nParts; numTrees = gridParams[idx].nTrees; fp = gridParams[idx].fp; ratio = gridParams[idx].ratio;
	k = gridParams[idx].k; mtry = gridParams[idx].mtry;
}
void hyperSMURF::createPart() { part.reset(new Partition(foldManager, ttd.get(), nPart, fp, ratio, mm, nn)); }

static uint32_t laneMinUnits();
void hyperSMURF::smurfIt() {                                                // src/hyperSMURF.cpp:151-496 (CV paths)
	if (wmode == MODE_PREDICT) { predictIt(); return; }                     // :401-446
	if (wmode != MODE_CV && wmode != MODE_TRAIN) throw Error(PSM_E_ARG, "hyperSMURF::smurfIt: unknown mode");
	if (wmode == MODE_TRAIN && commonParams->forestDirname.empty()) throw Error(PSM_E_ARG, "mode \"train\" needs data.forestDir");
	uint32_t startingFold = 0, endingFold = nFolds;
	if (!inInternalCV && commonParams->minFold != (uint32_t)-1) startingFold = commonParams->minFold;
	if (!inInternalCV && commonParams->maxFold != (uint32_t)-1) endingFold = std::min(commonParams->maxFold, nFolds);   // clamped like initStandard (an "endingFold" beyond nFolds would index past the fold lists)
	if (woptimiz != OPT_NO) { tempcl1.assign(nn, 0.0); tempcl2.assign(nn, 0.0); }
	if (!inInternalCV) { foldAuroc.assign(nFolds, 0.0); foldAuprc.assign(nFolds, 0.0); }
	// scores live on the device for the whole CV (class1Prob/class2Prob of src/parSMURF.cpp:90-93); they are added into the
	// caller's host arrays once, at the end
	dev->check(psm_labels_upload(dev->ctx, y.data(), nn));
	dev->check(psm_scores_reset(dev->ctx));
	// fold lanes only where every fold runs with the same parameters and nothing is flushed to the host between folds
	if (foldsOnDevice()) { PhaseScope ps(PH_FOLD_BEGIN); dev->uploadFolds(foldManager); }
	foldBase = startingFold; foldsRun = endingFold - startingFold;
	unitMode = willShardUnits();
	const uint32_t lanesHere = (wmode == MODE_CV && woptimiz == OPT_NO && !(commonParams->customCV && !inInternalCV) && endingFold > startingFold + 1)
	                               ? std::min<uint32_t>(std::max<uint32_t>(foldLanes, 1), endingFold - startingFold) : 1;

	for (uint32_t currentFold = startingFold; currentFold < endingFold; currentFold++) {
		if (woptimiz == OPT_GRID) {                                         // :176-228 grid exploration on the inner folds
			hyperSMURF inner(commonParams, gridParams, foldManager, dev, y, tempcl1.data(), tempcl2.data(), rng);
			// Multi-rank: the grid points of one outer fold are independent inner CVs, so each rank runs WHOLE inner CVs (its
			// share of the grid points, balanced by cost) without any communication, and one all-reduce of the [G][2] metric
			// table -- every entry written by exactly one rank, zeros elsewhere, so the sum is exact -- gives every rank the
			// table a single rank would have computed.  The one rand() stream the inner CVs draw their SMOTE samples from is
			// entered at each grid point's own offset (an inner CV consumes sum_f nPart * P_f * fp * (m+1) draws, known up front).
			const bool shardGrid = shardGridPoints && world > 1 && (bool)reduce;
			if (shardGrid) inner.setSharding(0, 1, nullptr); else inner.setSharding(rank, world, reduce);
			inner.curvesTieMode = curvesTieMode;
			inner.foldLanes = foldLanes;
			inner.initInternalCV(currentFold);
			const uint32_t G = (uint32_t)gridParams.size();
			std::vector<uint64_t> drawOff(G + 1, 0);
			{
				uint64_t posSum = 0;
				for (uint32_t f = 0; f + 1 < nFolds; f++) posSum += inner.ttd->trngPosNum[f];
				for (uint32_t i = 0; i < G; i++) drawOff[i + 1] = drawOff[i] + (uint64_t)gridParams[i].nParts * posSum * gridParams[i].fp * (mm + 1);
			}
			const std::vector<uint32_t> owner = shardGrid ? gridOwners(gridParams, world) : std::vector<uint32_t>(G, 0);
			const GlibcRand base = *rng;
			GlibcRand local = base;
			if (shardGrid) inner.rng = &local;
			std::string statFileName = "fold" + std::to_string(currentFold) + ".dat";
			auto statLine = [&](uint32_t i) {
				std::ofstream tf(statFileName.c_str(), std::ios::app);   // :191-199
				tf << std::to_string(i) << " " << std::to_string(gridParams[i].nParts) << " " << std::to_string(gridParams[i].fp) << " "
				   << std::to_string(gridParams[i].ratio) << " " << std::to_string(gridParams[i].k) << " " << std::to_string(gridParams[i].nTrees) << " "
				   << std::to_string(gridParams[i].mtry) << " " << std::to_string(gridParams[i].auroc) << " " << std::to_string(gridParams[i].auprc) << " "
				   << std::endl;
			};
			for (uint32_t i = 0; i < G; i++) {
				if (shardGrid) {
					if (owner[i] != rank) { gridParams[i].auroc = 0; gridParams[i].auprc = 0; continue; }
					local = base;
					local.discard(drawOff[i]);
				}
				std::fill(tempcl1.begin(), tempcl1.end(), 0.0);
				std::fill(tempcl2.begin(), tempcl2.end(), 0.0);
				inner.setGridParams(i);
				inner.createPart();
				inner.smurfIt();
				if (rank == 0 && !shardGrid) statLine(i);
			}
			if (shardGrid) {
				PhaseScope ps(PH_REDUCE);
				std::vector<double> gm(2 * (size_t)G);
				for (uint32_t i = 0; i < G; i++) { gm[2 * i] = gridParams[i].auroc; gm[2 * i + 1] = gridParams[i].auprc; }
				double* sp = nullptr;
				dev->check(psm_scratch_f64(dev->ctx, 2 * G, &sp));
				dev->check(psm_scratch_write(dev->ctx, gm.data(), 2 * G));
				reduce(sp, 2ull * G);
				dev->check(psm_scratch_read(dev->ctx, gm.data(), 2 * G));
				for (uint32_t i = 0; i < G; i++) { gridParams[i].auroc = gm[2 * i]; gridParams[i].auprc = gm[2 * i + 1]; if (rank == 0) statLine(i); }
				*rng = base;
				rng->discard(drawOff[G]);
			}
			// the inner CVs used the device score arrays: start the outer fold from a clean device state (the outer scores of
			// earlier folds were already added into the host arrays below)
			dev->check(psm_labels_upload(dev->ctx, y.data(), nn));
			dev->check(psm_scores_reset(dev->ctx));
			auto best = std::max_element(gridParams.begin(), gridParams.end(), [](const GridParams& a, const GridParams& b) { return a.auprc < b.auprc; });
			nPart = best->nParts; fp = best->fp; ratio = best->ratio; k = best->k; numTrees = best->nTrees; mtry = best->mtry;   // :207-224
			part->nPart = nPart; part->fp = fp; part->ratio = ratio;
		} else if (woptimiz == OPT_AUTOGP) {
			throw Error(PSM_E_ARG, "optimizer \"autogp\" (spearmint via std::system) is out of scope of the GPU path");
		}
		if (!inInternalCV && commonParams->customCV) {                      // :230-248 params.dat override
			const GridParams& g = gridParams[currentFold];
			nPart = g.nParts; fp = g.fp; ratio = g.ratio; k = g.k; numTrees = g.nTrees; mtry = g.mtry;
			part->nPart = nPart; part->fp = fp; part->ratio = ratio;
		}
		if (lanesHere > 1 || unitMode) {
			uint32_t lanesNow = lanesHere;
			if (unitMode) {                                                 // a lane is only worth its set-up with ~16 or more partitions of its own
				uint64_t mine = 0;
				for (uint32_t f = startingFold; f < endingFold; f++) { uint32_t a = 0, b = 0; partRange(true, foldsRun, nPart, world, rank, f - foldBase, &a, &b); mine += b - a; }
				lanesNow = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(mine / laneMinUnits(), 1), std::max<uint32_t>(foldLanes, 1));
			}
			runFoldsInLanes(startingFold, endingFold, lanesNow);
			trace_mark("lanes joined, scores merged");
			break;
		}
		uint32_t first = 0, last = 0;
		partRange(false, foldsRun, nPart, world, rank, currentFold - foldBase, &first, &last);
		runFold(currentFold, first, last, dev, part.get(), rng, nullptr);
	}

	if (wmode == MODE_TRAIN) return;
	// the final scores come to the host (rank 0 only with scoresOnRootOnly) on a thread and a stream of their own, next to the
	// curves, which read the device arrays (plain CV; the inner CVs and the grid's outer folds compute their curves from the
	// fetched host arrays, so there the fetch comes first)
	const bool skipFetch = scoresOnRootOnly && world > 1 && rank != 0 && woptimiz == OPT_NO && !inInternalCV;
	auto fetch = [&]() {
		PhaseScope ps(PH_SCATTER);
		if (skipFetch) return;
		if (scoresAreOutputs && woptimiz == OPT_NO && !inInternalCV) dev->check(psm_scores_read(dev->ctx, class1Prob, class2Prob));
		else dev->check(psm_scores_get(dev->ctx, class1Prob, class2Prob));
	};
	std::thread fetcher;
	std::exception_ptr fetchErr;
	const bool overlapFetch = !inInternalCV && woptimiz == OPT_NO && !skipFetch;
	auto startFetch = [&]() {
		if (overlapFetch && !fetcher.joinable())
			fetcher = std::thread([&]() { try { dev->check(psm_ctx_bind_thread(dev->ctx)); fetch(); } catch (...) { fetchErr = std::current_exception(); } });
	};
	bool globalCurvesDone = false;
	if (unitMode) {
		{
			PhaseScope ps(PH_REDUCE);                                          // the one cross-rank sum of a unit-sharded CV
			double* ptr = nullptr; uint32_t cntN = 0;
			dev->check(psm_scores_devptr(dev->ctx, &ptr, &cntN));
			dev->check(psm_sync(dev->ctx));
			trace_mark("own work done (stream drained)");
			reduce(ptr, 2ull * cntN);
			trace_mark("scores all-reduced");
		}
		startFetch();                                                       // the summed arrays are final from here on
		{
			// Curves from the summed arrays, spread over the ranks: the per-fold AUROC / AUPRC (:365-396) of fold f on rank
			// f % world, the global pair (:484-488) on the next rank (one that has no fold when world > nFolds, and not the rank
			// that is fetching the scores); one all-reduce of the [nFolds + 1][2] table -- every entry written by one rank,
			// zeros elsewhere -- hands all of them to everybody.
			PhaseScope ps(PH_CURVES);
			std::vector<double> fm(2 * (size_t)nFolds + 2, 0.0);
			if (evalFoldCurves)
				for (uint32_t f = startingFold; f < endingFold; f++) {
					if ((f - startingFold) % world != rank) continue;
					double r2[2];
					if (foldsOnDevice()) dev->check(psm_scores_fold_curves_device(dev->ctx, f, curvesTieMode, r2));
					else {
						ttd->ensureTest(f);
						dev->check(psm_scores_fold_curves(dev->ctx, ttd->testPosIdx[f].data(), ttd->testPosNum[f], ttd->testNegIdx[f].data(), ttd->testNegNum[f], curvesTieMode, r2));
					}
					noteCurves(dev);
					fm[2 * f] = r2[0]; fm[2 * f + 1] = r2[1];
				}
			const uint32_t globalRank = evalFoldCurves ? foldsRun % world : world - 1;
			if (rank == globalRank) {
				double r2[2];
				dev->check(psm_scores_curves(dev->ctx, curvesTieMode, r2));       // Curves(y, class1Prob), scores on the device
				noteCurves(dev);
				fm[2 * (size_t)nFolds] = r2[0]; fm[2 * (size_t)nFolds + 1] = r2[1];
			}
			double* sp = nullptr;
			dev->check(psm_scratch_f64(dev->ctx, 2 * nFolds + 2, &sp));
			dev->check(psm_scratch_write(dev->ctx, fm.data(), 2 * nFolds + 2));
			reduce(sp, 2ull * nFolds + 2);
			dev->check(psm_scratch_read(dev->ctx, fm.data(), 2 * nFolds + 2));
			if (evalFoldCurves)
				for (uint32_t f = startingFold; f < endingFold; f++) {
					foldAuroc[f] = fm[2 * f]; foldAuprc[f] = fm[2 * f + 1];
					if (verboseLevel > VERBSILENT && rank == 0)
						std::cout << "External CV on fold " << f << " -> AUROC: " << foldAuroc[f] << " - AUPRC: " << foldAuprc[f] << std::endl;
				}
			auroc = fm[2 * (size_t)nFolds]; auprc = fm[2 * (size_t)nFolds + 1];
			globalCurvesDone = true;
		}
	}
	trace_mark("fold curves done");
	if (overlapFetch) startFetch();
	else { fetch(); trace_mark("scores on the host"); }
	struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{fetcher};
	if (inInternalCV) {                                                     // :456-482
		std::vector<uint32_t> tl; std::vector<double> tp;
		for (uint32_t f = 0; f < nFolds; f++) {
			ttd->ensure(f);
			for (uint32_t v : ttd->testPosIdx[f]) { tl.push_back(y[v]); tp.push_back(class1Prob[v]); }
			for (uint32_t v : ttd->testNegIdx[f]) { tl.push_back(y[v]); tp.push_back(class1Prob[v]); }
		}
		Curves c(dev, tl, tp.data(), curvesTieMode);
		gridParams[idxInGridParams].auroc = c.evalAUROC_ok();
		gridParams[idxInGridParams].auprc = c.evalAUPRC();
	} else {                                                                // :484-488
		PhaseScope ps(PH_CURVES);
		if (!globalCurvesDone) {
			double r2[2];
			if (woptimiz == OPT_NO) dev->check(psm_scores_curves(dev->ctx, curvesTieMode, r2));        // Curves(y, class1Prob), scores still on the device
			else dev->check(psm_curves(dev->ctx, y.data(), class1Prob, nn, curvesTieMode, r2));
			noteCurves(dev);
			auroc = r2[0];
			auprc = r2[1];
		}
		if (verboseLevel > VERBSILENT && rank == 0) std::cout << "AUROC: " << auroc << " - AUPRC: " << auprc << std::endl;
	}
	if (fetcher.joinable()) { fetcher.join(); trace_mark("scores on the host"); }
	if (fetchErr) std::rethrow_exception(fetchErr);
}

// one fold of the CV loop (src/hyperSMURF.cpp:255-396) on Device lane `d`: `prt` is this lane's Partition view, `r` the
// rand() replica positioned at the fold's first SMOTE draw
struct hyperSMURF::FoldGate {
	std::mutex m;
	std::condition_variable cv;
	std::vector<int> state;            // per fold: 0 running, 1 scores ready for the cross-rank reduce, 2 reduced
	std::vector<double*> ptr;
	std::vector<uint64_t> count;
	bool failed = false;
};

// which rank runs the inner CV of each grid point: longest-processing-time first over cost ~ partitions x training rows x trees
std::vector<uint32_t> hyperSMURF::gridOwners(const std::vector<GridParams>& grid, uint32_t world) {
	const uint32_t G = (uint32_t)grid.size();
	std::vector<uint32_t> owner(G, 0), order(G);
	std::iota(order.begin(), order.end(), 0u);
	auto cost = [&](uint32_t i) { return (uint64_t)grid[i].nParts * (grid[i].fp + 1) * (grid[i].ratio ? grid[i].ratio + 1 : 4) * grid[i].nTrees; };
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cost(a) > cost(b); });
	std::vector<uint64_t> load(std::max(world, 1u), 0);
	for (uint32_t i : order) {
		const uint32_t r = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
		owner[i] = r; load[r] += cost(i);
	}
	return owner;
}

// how many pieces each fold share of a unit-sharded rank is cut into for `lanes` fold lanes (see runFoldsInLanes)
std::vector<uint32_t> hyperSMURF::laneCuts(const std::vector<uint32_t>& share, uint32_t lanes) {
	uint64_t mine = 0;
	for (uint32_t v : share) mine += v;
	std::vector<uint32_t> cut(share.size(), 0);
	if (!mine || !lanes) return cut;
	const uint64_t perLane = (mine + lanes - 1) / lanes;
	const uint32_t target = lanes * (perLane >= 64 ? 2u : 1u);
	uint32_t total = 0;
	for (size_t f = 0; f < share.size(); f++) {
		if (!share[f]) continue;
		cut[f] = (uint32_t)std::max<uint64_t>(1, ((uint64_t)share[f] * target + mine / 2) / mine);
		cut[f] = std::min(cut[f], std::max(1u, share[f] / std::max(1u, laneMinUnits() / 2)));   // no crumbs
		total += cut[f];
	}
	while (total > target) {                                               // rounding gave too many: merge where the pieces are smallest
		size_t best = share.size();
		for (size_t f = 0; f < share.size(); f++)
			if (cut[f] > 1 && (best == share.size() || share[f] / cut[f] < share[best] / cut[best])) best = f;
		if (best == share.size()) break;
		cut[best]--; total--;
	}
	return cut;
}

void hyperSMURF::partRange(bool units, uint32_t nFoldsRun, uint32_t nPart, uint32_t world, uint32_t rank, uint32_t foldIdx, uint32_t* first, uint32_t* last) {
	auto chunk = [&](uint64_t total, uint64_t* a, uint64_t* b) {           // total/world (+1 for the first total%world ranks), contiguous
		const uint64_t base = total / world, rem = total % world;
		*a = base * rank + std::min<uint64_t>(rank, rem);
		*b = *a + base + (rank < rem ? 1 : 0);
	};
	uint64_t a, b;
	if (!units) { chunk(nPart, &a, &b); *first = (uint32_t)a; *last = (uint32_t)b; return; }
	chunk((uint64_t)nFoldsRun * nPart, &a, &b);
	const uint64_t f0 = (uint64_t)foldIdx * nPart, f1 = f0 + nPart;
	const uint64_t lo = std::max(a, f0), hi = std::min(b, f1);
	if (lo >= hi) { *first = *last = 0; return; }
	*first = (uint32_t)(lo - f0); *last = (uint32_t)(hi - f0);
}

// partitions [first, last) of fold `currentFold` (this rank's share of it, or one piece of the share)
void hyperSMURF::runFold(uint32_t currentFold, uint32_t first, uint32_t last, Device* d, Partition* prt, GlibcRand* r, FoldGate* gate) {
	const uint32_t cnt = last - first;
	d->invalidateFold();
	if (foldsOnDevice()) {                                                 // :255 with the index arrays built (and shuffled) on the device
		PhaseScope ps(PH_FOLD_BEGIN);
		prt->setFoldCounts(currentFold);
		d->beginFoldOnDevice(prt);
	} else {
		{ PhaseScope ps(PH_FOLDS_TTD); prt->setFold(currentFold); }  // :255 (builds + shuffles the fold on the host if nobody has yet)
		{ PhaseScope ps(PH_FOLD_BEGIN); d->ensureFold(prt); }
	}
	trace_mark("  fold " + std::to_string(currentFold) + " begun");
	const uint64_t perPartDraws = (uint64_t)prt->trngPosNum * fp * (mm + 1);
	auto skipDraws = [&](uint64_t count) { r->discard(count); };      // keep the one SMOTE stream in partition order on every rank
	{ PhaseScope ps(PH_DRAWS); skipDraws(perPartDraws * first); }
	uint32_t cap = 1;
	if (cnt) d->check(psm_batch_capacity(d->ctx, nPart, fp, ratio, numTrees, mtry, first, &cap));
	for (uint32_t b0 = first; b0 < last; b0 += cap) {                   // the `#pragma omp for` over partitions (:257-261), as batches
		const uint32_t b1 = std::min(last, b0 + cap);
		SamplerKNN samp(prt, wmode, d, mm, nn, k, r);                   // :275
		samp.setPartitionRange(b0, b1);                                 // :276
		samp.sample();                                                  // :277
		trace_mark("  batch [" + std::to_string(b0) + "," + std::to_string(b1) + ") sampled");
		rfRanger rf(d, mm, false, &samp, numTrees, mtry, commonParams->rfThr, seed + currentFold * nPart);   // :273,302
		{ PhaseScope ps(PH_TRAIN); rf.train(commonParams->rfVerbose); }   // :303
		trace_mark("  trained");
		if (wmode == MODE_TRAIN) {                                      // :349-360 training only: one forest file per partition
			for (uint32_t p = b0; p < b1; p++) rf.saveForest(p, commonParams->forestDirname);
			continue;
		}
		samp.copyTestSet();                                             // :316
		rfRanger rfTest(d, mm, true, &samp, numTrees, mtry, commonParams->rfThr, 2 * (seed + currentFold * nPart));   // :330
		{ PhaseScope ps(PH_PREDICT); rfTest.predict(commonParams->rfVerbose); d->check(psm_sync(d->ctx)); }   // :331
		trace_mark("  predicted");
		samp.accumulateAndDivideResInProbVect(nPart);                   // :344-346 (done on the device by predict)
	}
	{ PhaseScope ps(PH_DRAWS); skipDraws(perPartDraws * (nPart - last)); }
	if (wmode == MODE_TRAIN) return;                                     // nothing was predicted
	if (unitMode) {                                                      // partial fold scores -> this rank's class1Prob/class2Prob; summed once at the end
		PhaseScope ps(PH_SCATTER);
		d->check(psm_fold_scatter_device(d->ctx));
		return;
	}
	if (world > 1 && reduce) {
		PhaseScope ps(PH_REDUCE);                                          // src/hyperSMURF_MPI.cpp:594-618
		double* ptr = nullptr; uint32_t nte = 0;
		d->check(psm_fold_scores_devptr(d->ctx, &ptr, &nte));
		d->check(psm_sync(d->ctx));
		if (!gate) reduce(ptr, 2ull * nte);
		else {
			// fold lanes: every rank must issue its collectives in the same order, so the calling thread of smurfIt does
			// the reduces in fold order (runFoldsInLanes) and this lane waits for its turn
			std::unique_lock<std::mutex> lk(gate->m);
			gate->ptr[currentFold] = ptr; gate->count[currentFold] = 2ull * nte; gate->state[currentFold] = 1;
			gate->cv.notify_all();
			gate->cv.wait(lk, [&] { return gate->state[currentFold] == 2 || gate->failed; });
			if (gate->failed) throw Error(PSM_E_CUDA, "hyperSMURF: another fold lane failed");
		}
	}
	{ PhaseScope ps(PH_SCATTER); d->check(psm_fold_scatter_device(d->ctx)); }   // :344-346 scatter by example id, on the device

	if (woptimiz != OPT_NO) {                                          // grid mode: flush this fold's scores to the host now (see above)
		PhaseScope ps(PH_SCATTER);
		d->check(psm_scores_get(d->ctx, class1Prob, class2Prob));
		d->check(psm_scores_reset(d->ctx));
	}
	if (evalFoldCurves && !inInternalCV) {                              // :365-396 per-fold AUROC / AUPRC (labels y[idx], scores class1Prob[idx])
		PhaseScope ps(PH_CURVES);
		double r2[2];
		d->check(psm_fold_curves(d->ctx, curvesTieMode, r2));       // Curves(tempLabels, tempPreds): AUROC first, then AUPRC (:382-384)
		noteCurves(d);
		foldAuroc[currentFold] = r2[0];
		foldAuprc[currentFold] = r2[1];
		if (verboseLevel > VERBSILENT && rank == 0)
			std::cout << "External CV on fold " << currentFold << " -> AUROC: " << foldAuroc[currentFold] << " - AUPRC: " << foldAuprc[currentFold] << std::endl;
	}
}

static uint32_t laneMinUnits() { static const uint32_t v = getenv("PSM_LANE_MIN_UNITS") ? (uint32_t)std::max(1, atoi(getenv("PSM_LANE_MIN_UNITS"))) : 12u; return v; }
void hyperSMURF::runFoldsInLanes(uint32_t startingFold, uint32_t endingFold, uint32_t lanes) {
	dev->setLaneCount(lanes);
	// where every fold starts in the one SMOTE rand() stream: a fold consumes nPart * P_fold * fp * (m+1) draws whatever
	// share of its partitions this rank assembles (runFold skips the others), and P_fold is known before anything is built
	std::vector<uint64_t> off(endingFold + 1, 0);
	for (uint32_t f = startingFold; f < endingFold; f++) off[f + 1] = off[f] + (uint64_t)nPart * ttd->trngPosNum[f] * fp * (mm + 1);
	// the pieces of work the lanes take, in order: this rank's share of every fold; unit sharding cuts a share further
	struct Piece { uint32_t fold, first, last; };
	std::vector<Piece> pieces;
	{
		// pieces: about two per lane, so that the lanes finish together whatever the fold boundaries do to the shares; one per
		// lane when the share is small (a piece costs a few ms whatever its size: fold set-up and the latency of ~30 tree levels).
		// The target number of pieces is dealt to the fold shares in proportion to their sizes (at least one each).
		std::vector<uint32_t> share(endingFold, 0);
		for (uint32_t f = startingFold; f < endingFold; f++) {
			uint32_t a = 0, b = 0;
			partRange(unitMode, foldsRun, nPart, world, rank, f - foldBase, &a, &b);
			share[f] = unitMode ? b - a : 0;
		}
		const std::vector<uint32_t> cut = laneCuts(share, lanes);
		for (uint32_t f = startingFold; f < endingFold; f++) {
			uint32_t a = 0, b = 0;
			partRange(unitMode, foldsRun, nPart, world, rank, f - foldBase, &a, &b);
			if (!unitMode) { pieces.push_back({f, a, b}); continue; }          // fold rule: one piece per fold, even an empty one (its reduce still happens)
			const uint32_t np = cut[f];                                         // equal pieces of the fold share
			for (uint32_t i = 0; i < np; i++) pieces.push_back({f, a + (uint32_t)((uint64_t)(b - a) * i / np), a + (uint32_t)((uint64_t)(b - a) * (i + 1) / np)});
		}
		// largest first: the lanes take the pieces in this order
		if (unitMode) std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) { return x.last - x.first > y.last - y.first; });
	}
	const GlibcRand base = *rng;
	const double* xdev = nullptr;
	dev->check(psm_dataset_devptr(dev->ctx, &xdev));
	std::atomic<uint32_t> next(0);
	FoldGate gate;
	gate.state.assign(endingFold, 0); gate.ptr.assign(endingFold, nullptr); gate.count.assign(endingFold, 0);
	const bool gated = world > 1 && (bool)reduce && !unitMode;
	std::exception_ptr err;
	auto fail = [&]() {
		std::lock_guard<std::mutex> lk(gate.m);
		if (!err) err = std::current_exception();
		gate.failed = true;
		gate.cv.notify_all();
	};
	auto worker = [&](uint32_t li) {
		try {
			Device* d = dev->lane(li);
			d->check(psm_ctx_bind_thread(d->ctx));
			if (li) {                                                          // the extra lanes borrow lane 0's copy of x and keep their own score arrays
				d->attachDevice(xdev, nn, mm);
				d->check(psm_labels_upload(d->ctx, y.data(), nn));
				d->check(psm_scores_reset(d->ctx));
				if (foldsOnDevice()) d->attachFolds(dev);
			}
			Partition prt(foldManager, ttd.get(), nPart, fp, ratio, mm, nn);
			for (uint32_t round = 0;; round++) {
				// unit sharding: lane li always takes pieces li, li + lanes, ... of the size-ordered list, so a lane's batch sizes -- and
				// with them its device buffers -- are the same in every CV of a long-lived host (a lane that meets a larger piece than
				// before has to grow its buffers, and cudaFree / cudaMalloc stall the whole device for tens of ms: seen as one
				// straggling rank per step on 8 GPUs).  Whole folds (equal sizes) are handed out dynamically.
				const uint32_t pi = unitMode ? li + round * lanes : next.fetch_add(1);
				if (pi >= pieces.size()) break;
				{ std::lock_guard<std::mutex> lk(gate.m); if (gate.failed) break; }
				const Piece& pc = pieces[pi];
				GlibcRand r = base;
				r.discard(off[pc.fold]);
				const std::string tag = "lane " + std::to_string(li) + " fold " + std::to_string(pc.fold) + " [" + std::to_string(pc.first) + "," + std::to_string(pc.last) + ")";
				trace_mark(tag + " begin");
				runFold(pc.fold, pc.first, pc.last, d, &prt, &r, gated ? &gate : nullptr);
				trace_mark(tag + " end");
			}
			d->check(psm_sync(d->ctx));
		} catch (...) { fail(); }
	};
	std::vector<std::thread> threads;
	for (uint32_t li = 0; li < lanes; li++) threads.emplace_back(worker, li);
	if (gated) {
		for (uint32_t f = startingFold; f < endingFold; f++) {                  // cross-rank reduces in fold order, on the caller's thread
			std::unique_lock<std::mutex> lk(gate.m);
			gate.cv.wait(lk, [&] { return gate.state[f] == 1 || gate.failed; });
			if (gate.failed) break;
			double* p = gate.ptr[f]; const uint64_t c = gate.count[f];
			lk.unlock();
			try { reduce(p, c); } catch (...) { fail(); break; }
			lk.lock();
			gate.state[f] = 2;
			gate.cv.notify_all();
		}
	}
	for (auto& t : threads) t.join();
	if (err) std::rethrow_exception(err);
	rng->discard(off[endingFold]);
	for (uint32_t li = 1; li < lanes; li++) dev->check(psm_scores_merge(dev->ctx, dev->lane(li)->ctx));
}

// predict mode (src/hyperSMURF.cpp:401-446): every example is a test row of fold 0; partition p is predicted by the forest
// saved as <forestDir>/<p>.out.forest and the ensemble average is accumulated exactly as in the CV path.
void hyperSMURF::predictIt() {
	if (commonParams->forestDirname.empty()) throw Error(PSM_E_ARG, "mode \"predict\" needs data.forestDir");
	dev->check(psm_labels_upload(dev->ctx, y.data(), nn));
	dev->check(psm_scores_reset(dev->ctx));
	{ PhaseScope ps(PH_FOLDS_TTD); part->setFold(0); }
	dev->invalidateFold();
	{ PhaseScope ps(PH_FOLD_BEGIN); dev->ensureFold(part.get()); }
	uint32_t cnt = nPart / world + ((nPart % world) > rank ? 1 : 0);
	uint32_t first = 0;
	for (uint32_t r2 = 0; r2 < rank; r2++) first += nPart / world + ((nPart % world) > r2 ? 1 : 0);
	const uint32_t last = first + cnt;
	// one batch = as many forests as keep the per-(partition, row) score planes within ~4 GiB
	const uint64_t nte = (uint64_t)part->testPosNum + part->testNegNum;
	const uint32_t cap = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(256, (4ull << 30) / std::max<uint64_t>(1, 16 * nte)));
	for (uint32_t b0 = first; b0 < last; b0 += cap) {
		const uint32_t b1 = std::min(last, b0 + cap);
		rfRanger rfTest(dev, commonParams->forestDirname, b0, b1, nPart, mm, numTrees, mtry, commonParams->rfThr, seed + b0);   // :421-425
		{ PhaseScope ps(PH_PREDICT); rfTest.predict(commonParams->rfVerbose); dev->check(psm_sync(dev->ctx)); }
	}
	if (world > 1 && reduce) {
		PhaseScope ps(PH_REDUCE);
		double* ptr = nullptr; uint32_t nte2 = 0;
		dev->check(psm_fold_scores_devptr(dev->ctx, &ptr, &nte2));
		dev->check(psm_sync(dev->ctx));
		reduce(ptr, 2ull * nte2);
	}
	{ PhaseScope ps(PH_SCATTER); dev->check(psm_fold_scatter_device(dev->ctx)); dev->check(psm_scores_get(dev->ctx, class1Prob, class2Prob)); }
	PhaseScope ps(PH_CURVES);                                                // :484-488
	double r2[2];
	dev->check(psm_scores_curves(dev->ctx, curvesTieMode, r2));
	noteCurves(dev);
	auroc = r2[0]; auprc = r2[1];
	if (verboseLevel > VERBSILENT && rank == 0) std::cout << "AUROC: " << auroc << " - AUPRC: " << auprc << std::endl;
}

// ------------------------------------------------------------------------------------------ JSON config
namespace {
struct JVal {
	enum T { NUL, BOOL, NUM, STR, ARR, OBJ } t = NUL;
	bool b = false; double num = 0; std::string s;
	std::vector<JVal> arr;
	std::vector<std::pair<std::string, JVal>> obj;
	const JVal* get(const std::string& k) const { for (auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
};
struct JParser {
	const std::string& s; size_t i = 0;
	explicit JParser(const std::string& s_) : s(s_) {}
	void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
	[[noreturn]] void bad(const char* w) { throw Error(PSM_E_ARG, std::string("config JSON: ") + w + " at offset " + std::to_string(i)); }
	JVal parse() {
		ws();
		if (i >= s.size()) bad("unexpected end");
		JVal v;
		char c = s[i];
		if (c == '{') {
			v.t = JVal::OBJ; i++;
			while (true) {
				ws();
				if (i < s.size() && s[i] == '}') { i++; break; }            // also accepts a trailing comma like jsoncons does (cfgEx/gridTune.json)
				JVal key = parse();
				if (key.t != JVal::STR) bad("object key must be a string");
				ws();
				if (i >= s.size() || s[i] != ':') bad("':' expected");
				i++;
				v.obj.emplace_back(key.s, parse());
				ws();
				if (i < s.size() && s[i] == ',') { i++; continue; }
				if (i < s.size() && s[i] == '}') { i++; break; }
				bad("',' or '}' expected");
			}
		} else if (c == '[') {
			v.t = JVal::ARR; i++;
			while (true) {
				ws();
				if (i < s.size() && s[i] == ']') { i++; break; }
				v.arr.push_back(parse());
				ws();
				if (i < s.size() && s[i] == ',') { i++; continue; }
				if (i < s.size() && s[i] == ']') { i++; break; }
				bad("',' or ']' expected");
			}
		} else if (c == '"') {
			v.t = JVal::STR; i++;
			while (i < s.size() && s[i] != '"') {
				if (s[i] == '\\' && i + 1 < s.size()) {
					char e = s[i + 1];
					v.s += e == 'n' ? '\n' : e == 't' ? '\t' : e;
					i += 2;
				} else v.s += s[i++];
			}
			if (i >= s.size()) bad("unterminated string");
			i++;
		} else if (!s.compare(i, 4, "true")) { v.t = JVal::BOOL; v.b = true; i += 4; }
		else if (!s.compare(i, 5, "false")) { v.t = JVal::BOOL; v.b = false; i += 5; }
		else if (!s.compare(i, 4, "null")) { i += 4; }
		else {
			size_t j = i;
			while (j < s.size() && (isdigit((unsigned char)s[j]) || s[j] == '-' || s[j] == '+' || s[j] == '.' || s[j] == 'e' || s[j] == 'E')) j++;
			if (j == i) bad("unexpected character");
			v.t = JVal::NUM; v.num = std::stod(s.substr(i, j - i)); i = j;
		}
		return v;
	}
};
template <typename T> T jnum(const JVal* o, const char* k, T def) { const JVal* v = o ? o->get(k) : nullptr; return (v && v->t == JVal::NUM) ? (T)v->num : def; }
bool jbool(const JVal* o, const char* k, bool def) { const JVal* v = o ? o->get(k) : nullptr; return (v && v->t == JVal::BOOL) ? v->b : def; }
std::string jstr(const JVal* o, const char* k, const std::string& def) { const JVal* v = o ? o->get(k) : nullptr; return (v && v->t == JVal::STR) ? v->s : def; }
std::vector<int64_t> jarr(const JVal* o, const char* k) {
	std::vector<int64_t> out;
	const JVal* v = o ? o->get(k) : nullptr;
	if (v && v->t == JVal::ARR) for (auto& e : v->arr) out.push_back((int64_t)e.num);
	else if (v && v->t == JVal::NUM) out.push_back((int64_t)v->num);
	if (out.empty()) out.push_back(-1);                                     // sentinel, ArgHandler_new.cpp:333-334
	return out;
}
}  // namespace

Config Config::fromString(const std::string& text) {
	JParser p(text);
	JVal root = p.parse();
	if (root.t != JVal::OBJ) throw Error(PSM_E_ARG, "config JSON: top level must be an object");
	const JVal *exec = root.get("exec"), *data = root.get("data"), *sim = root.get("simulate"), *flds = root.get("folds"), *params = root.get("params");
	Config c;
	c.dataFilename = jstr(data, "dataFile", ""); c.foldFilename = jstr(data, "foldFile", ""); c.labelFilename = jstr(data, "labelFile", "");
	c.common.outfilename = jstr(data, "outFile", "output.txt");             // default: ArgHandler_new.cpp:144-148
	c.common.forestDirname = jstr(data, "forestDir", "");
	if (jbool(exec, "saveTime", false)) c.common.timeFilename = jstr(exec, "timeFile", "");
	c.simulate = jbool(sim, "simulation", false);
	c.common.nn = c.simulate ? jnum<uint32_t>(sim, "n", 1000) : 0;
	c.common.mm = c.simulate ? jnum<uint32_t>(sim, "m", 10) : 0;
	c.prob = c.simulate ? jnum<double>(sim, "prob", 0.01) : 0.0;
	c.common.nFolds = jnum<uint32_t>(flds, "nFolds", 3);
	c.common.minFold = jnum<uint32_t>(flds, "startingFold", (uint32_t)-1);
	c.common.maxFold = jnum<uint32_t>(flds, "endingFold", (uint32_t)-1);
	c.common.seed = jnum<uint32_t>(exec, "seed", 0);
	if (c.common.seed == 0) { c.common.seed = (uint32_t)time(nullptr); c.seedFromClock = true; }   // ArgHandler_new.cpp:257-262 (the caller also srand()s it)
	c.common.verboseLevel = jnum<uint32_t>(exec, "verboseLevel", 0);
	c.common.nThr = jnum<uint32_t>(exec, "ensThrd", 1);
	c.common.rfThr = jnum<uint32_t>(exec, "rfThrd", 1);
	c.printCfg = jbool(exec, "printCfg", false);
	std::string mode = jstr(exec, "mode", "cv"), optim = jstr(exec, "optimizer", "no");
	if (mode == "cv") c.common.wmode = MODE_CV;
	else if (mode == "train") { c.common.wmode = MODE_TRAIN; c.common.nFolds = 1; }
	else if (mode == "predict") { c.common.wmode = MODE_PREDICT; c.common.nFolds = 1; }
	else throw Error(PSM_E_ARG, "Invalid prediction mode. Please specify either 'cv', 'train' or 'predict'");
	if (c.common.wmode != MODE_CV) c.foldFilename.clear();                  // "Ignoring fold filename in train or test mode" (ArgHandler_new.cpp:172-176)
	if (optim == "no") c.common.woptimiz = OPT_NO;
	else if (optim == "grid") c.common.woptimiz = OPT_GRID;
	else if (optim == "autogp") c.common.woptimiz = OPT_AUTOGP;
	else throw Error(PSM_E_ARG, "Invalid optimizer");
	// cartesian product in the reference's loop order (ArgHandler_new.cpp:343-361); defaults of checkConfig (:294-330)
	for (int64_t v1 : jarr(params, "nParts")) for (int64_t v2 : jarr(params, "fp")) for (int64_t v3 : jarr(params, "ratio"))
	for (int64_t v4 : jarr(params, "k")) for (int64_t v5 : jarr(params, "nTrees")) for (int64_t v6 : jarr(params, "mtry")) {
		GridParams g;
		g.nParts = v1 < 0 ? 3 : (uint32_t)v1; g.fp = v2 < 0 ? 1 : (uint32_t)v2; g.ratio = v3 < 0 ? 1 : (uint32_t)v3;
		g.k = v4 < 0 ? 5 : (uint32_t)v4; g.nTrees = v5 < 0 ? 50 : (uint32_t)v5; g.mtry = v6 < 0 ? 0 : (uint32_t)v6;
		g.auroc = g.auprc = 0;
		c.grid.push_back(g);
	}
	return c;
}
Config Config::fromFile(const std::string& path) {
	std::ifstream f(path.c_str());
	if (!f) throw Error(PSM_E_ARG, "cannot open config file " + path);
	std::stringstream ss; ss << f.rdbuf();
	Config c = fromString(ss.str());
	c.common.cfgFilename = path;
	return c;
}
void Config::processMtry(uint32_t mm) {                                      // ArgHandler_new.cpp:366-375
	for (auto& g : grid) { if (g.mtry > mm) g.mtry = 0; if (g.mtry == 0) g.mtry = (uint32_t)std::sqrt((double)mm); }
}

}  // namespace parsmurf

// ------------------------------------------------------------------------------------------ C entry points for bindings
extern "C" {

typedef int (*psmh_reduce_fn)(void* user, double* dev_ptr, uint64_t count);

// Fold lanes of the CV entry points (hyperSMURF::foldLanes): psmh_set_fold_lanes(n) fixes the count for this process,
// otherwise the environment variable PSM_FOLD_LANES, otherwise kDefaultFoldLanes.
static constexpr uint32_t kDefaultFoldLanes = 4;
static std::atomic<uint32_t> g_foldLanes{0};
void psmh_set_fold_lanes(uint32_t lanes) { g_foldLanes = lanes; }
// Sharding rule of multi-rank CV runs (hyperSMURF::shardUnits): psmh_set_shard_units(1 / 0) fixes it for this process (-1 =
// back to the default), otherwise PSM_SHARD=unit|fold, otherwise unit sharding.
static std::atomic<int> g_shardUnits{-1};
void psmh_set_shard_units(int v) { g_shardUnits = v; }
// hyperSMURF::scoresOnRootOnly of the CV entry points (multi-rank runs: only rank 0 fetches the final score arrays)
static std::atomic<int> g_scoresRootOnly{0};
void psmh_set_scores_root_only(int v) { g_scoresRootOnly = v; }
// NCCL inside the library (psm_comm_*): 128-byte id from one rank, then a collective init on every rank's device handle
int psmh_comm_unique_id(void* id128) { return psm_comm_unique_id(id128); }
int psmh_device_comm_init(void* devHandle, const void* id128, uint32_t rank, uint32_t world, char* errbuf, size_t errlen) {
	try { ((parsmurf::Device*)devHandle)->commInit(id128, rank, world); return 0; }
	catch (const std::exception& e) { if (errbuf && errlen) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; } return -1; }
}
int psmh_device_comm_info(void* devHandle, uint32_t* rank, uint32_t* world, int* ncclVersion) { return psm_comm_info(((parsmurf::Device*)devHandle)->ctx, rank, world, ncclVersion); }
static bool shard_units_default() {
	const int v = g_shardUnits.load();
	if (v >= 0) return v != 0;
	const char* e = getenv("PSM_SHARD");
	return !(e && std::string(e) == "fold");
}
void psmh_grid_owners(const uint32_t* grid, uint32_t G, uint32_t world, uint32_t* owners) {
	std::vector<parsmurf::GridParams> gp(G);
	for (uint32_t i = 0; i < G; i++) gp[i] = parsmurf::GridParams{grid[6 * i], grid[6 * i + 1], grid[6 * i + 2], grid[6 * i + 3], grid[6 * i + 4], grid[6 * i + 5], 0, 0};
	const std::vector<uint32_t> o = parsmurf::hyperSMURF::gridOwners(gp, world);
	for (uint32_t i = 0; i < G; i++) owners[i] = o[i];
}
void psmh_lane_cuts(const uint32_t* share, uint32_t nShares, uint32_t lanes, uint32_t* cuts) {
	const std::vector<uint32_t> c = parsmurf::hyperSMURF::laneCuts(std::vector<uint32_t>(share, share + nShares), lanes);
	for (uint32_t i = 0; i < nShares; i++) cuts[i] = c[i];
}
void psmh_part_range(int units, uint32_t nFoldsRun, uint32_t nPart, uint32_t world, uint32_t rank, uint32_t foldIdx, uint32_t* first, uint32_t* last) {
	parsmurf::hyperSMURF::partRange(units != 0, nFoldsRun, nPart, world, rank, foldIdx, first, last);
}
static uint32_t fold_lanes_default();
uint32_t psmh_get_fold_lanes(void) { return fold_lanes_default(); }
static uint32_t fold_lanes_default() {
	uint32_t v = g_foldLanes.load();
	if (v == 0) { const char* e = getenv("PSM_FOLD_LANES"); v = e ? (uint32_t)atoi(e) : kDefaultFoldLanes; }
	return std::min<uint32_t>(std::max<uint32_t>(v, 1), 16);
}

// Whole CV through the host classes, like parSMURF1's main (src/parSMURF.cpp:97-106) with data already in memory.
// foldAssign == NULL -> random stratified folds (src/folds.cpp:27-77).  metrics = {auroc, auprc}; foldMetrics [nFolds][2] or NULL.
// x_is_device != 0: `x` is a device pointer (dataset already resident in HBM).
static int run_mode(int mode, const char* forestDir, const double* x, int x_is_device, uint32_t n, uint32_t m, const uint32_t* labels,
                    const uint32_t* foldAssign, uint32_t nFolds,
                uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed, int device, void* stream,
                int tieMode, int evalFoldCurves, uint32_t rank, uint32_t world, psmh_reduce_fn reduce, void* reduceUser, uint64_t batchBudget,
                int profile, double* class1, double* class2, double* metrics, double* foldMetrics, double* kernelMs, uint64_t* kernelLaunches,
                uint64_t* stats3, double* phases, void* devHandle, char* errbuf, size_t errlen) {
	using namespace parsmurf;
	try {
		for (int i = 0; i < PH_COUNT; i++) g_phase[i] = 0;
		double tSetup = now_s();
		{ std::lock_guard<std::mutex> lk(g_markMtx); g_marks.clear(); }
		trace_mark("start");
		std::unique_ptr<Device> owned;
		if (!devHandle) owned.reset(new Device(device, stream));
		Device& dev = devHandle ? *(Device*)devHandle : *owned;
		if (devHandle && stream) dev.check(psm_set_stream(dev.ctx, stream));
		const uint32_t lanes = mode == MODE_CV ? std::min<uint32_t>(fold_lanes_default(), std::max<uint32_t>(nFolds, 1)) : 1;
		dev.setLaneCount(lanes);
		for (uint32_t li = 0; li < dev.laneCount(); li++) {
			Device* d = dev.lane(li);
			d->check(psm_profile_reset(d->ctx));
			// every lane holds its own partition batch: keep the sum of the budgets within one GPU's memory
			const uint64_t perLane = batchBudget ? batchBudget : std::min<uint64_t>(24ull << 30, (120ull << 30) / lanes);
			d->check(psm_set_batch_budget(d->ctx, perLane));
			// profile: 0 off, 1 every kernel class, 2 only split search (the kernel the roofline is reported on)
			d->check(psm_profile_classes(d->ctx, profile == 2 ? (1u << 5) : 0xFFFFFFFFu));
			d->check(psm_profile_enable(d->ctx, profile != 0));
		}
		GlibcRand rng(1);
		struct Pooled {                                     // the label copy lives in a recycled buffer too
			std::vector<uint32_t> v;
			Pooled(const uint32_t* src, size_t cnt) { if (src) { v = g_pool.take(cnt); v.assign(src, src + cnt); } }
			~Pooled() { g_pool.give(v); }
		} yBuf(labels, n);
		std::vector<uint32_t>& y = yBuf.v;
		Folds folds(nFolds, n, y.data(), foldAssign, n, foldAssign == nullptr, &rng);   // the fold ids are only borrowed for the call
		// folds from an assignment: the per-class lists are built in HBM (one stable device sort) once the dataset and the labels
		// are there; random folds are a rand()-driven shuffle and stay on the host
		const bool foldsOnDev = foldAssign != nullptr && mode == MODE_CV && getenv("PSM_HOST_FOLDS") == nullptr;
		if (foldsOnDev) {
			tSetup = now_s();
			if (x_is_device) dev.attachDevice(x, n, m); else dev.upload(x, n, m);
			g_phase[PH_SETUP] += now_s() - tSetup;
			trace_mark("dataset attached");
			PhaseScope ps(PH_FOLDS_TTD);
			dev.check(psm_labels_upload(dev.ctx, y.data(), n));
			folds.computeFoldsOnDevice(&dev);
		} else { PhaseScope ps(PH_FOLDS_TTD); folds.computeFolds(); }
		trace_mark("folds computed");
		CommonParams cp;
		cp.nn = n; cp.mm = m; cp.nFolds = nFolds; cp.seed = seed; cp.wmode = (uint8_t)mode; cp.woptimiz = OPT_NO;
		if (forestDir) cp.forestDirname = forestDir;
		std::vector<GridParams> gp(1);
		gp[0] = GridParams{nParts, fp, ratio, k, nTrees, mtry, 0, 0};
		hyperSMURF smurfer(&cp, gp, &folds, &dev, y, class1, class2, &rng);
		smurfer.curvesTieMode = tieMode; smurfer.evalFoldCurves = evalFoldCurves != 0;
		smurfer.foldLanes = lanes;
		smurfer.shardUnits = shard_units_default();
		smurfer.deviceFolds = getenv("PSM_HOST_FOLDS") == nullptr;
		smurfer.scoresAreOutputs = mode == MODE_CV;          // class1 / class2 of psmh_cv_run are outputs
		// the cross-rank sum: the caller's callback if it gave one, else the library's own NCCL communicator (Device::commInit)
		if (world > 1 && !reduce && !dev.hasComm()) throw Error(PSM_E_ARG, "multi-rank run without a reduce callback: initialise the device's communicator first (psmh_device_comm_init)");
		if (world > 1) smurfer.setSharding(rank, world, [&](double* p, uint64_t c) {
			if (!reduce) { dev.allReduce(p, c); return; }
			if (reduce(reduceUser, p, c)) throw Error(PSM_E_CUDA, "reduce callback failed");
		});
		smurfer.scoresOnRootOnly = g_scoresRootOnly.load() != 0;
		smurfer.initStandard();                             // starts the per-fold index builders in the background ...
		if (!foldsOnDev) {
			tSetup = now_s();
			if (x_is_device) dev.attachDevice(x, n, m); else dev.upload(x, n, m);   // ... which overlap the H2D copy of x
			g_phase[PH_SETUP] += now_s() - tSetup;
			trace_mark("dataset attached");
		}
		smurfer.smurfIt();
		trace_mark("smurfIt done");
		if (g_trace) {
			std::lock_guard<std::mutex> lk(g_markMtx);
			std::string line = "[psm timeline rank " + std::to_string(rank) + "]";
			char buf[96];
			for (auto& mk : g_marks) { snprintf(buf, sizeof buf, " %.2f", 1e3 * (mk.second - g_marks[0].second)); line += " | " + mk.first + buf; }
			// PSM_TIMELINE=1: stderr; PSM_TIMELINE=<path prefix>: appended to <prefix>.rank<r> (ranks sharing one stderr interleave)
			const char* tl = getenv("PSM_TIMELINE");
			FILE* out = (tl && strchr(tl, '/')) ? fopen((std::string(tl) + ".rank" + std::to_string(rank)).c_str(), "a") : nullptr;
			fprintf(out ? out : stderr, "%s\n", line.c_str());
			if (out) fclose(out);
		}
		if (metrics) { metrics[0] = smurfer.auroc; metrics[1] = smurfer.auprc; }
		if (foldMetrics && mode == MODE_CV) for (uint32_t f = 0; f < nFolds; f++) { foldMetrics[2 * f] = smurfer.foldAuroc[f]; foldMetrics[2 * f + 1] = smurfer.foldAuprc[f]; }
		if (kernelMs) for (int i = 0; i < PSM_NUM_KERNEL_CLASSES; i++) kernelMs[i] = 0;
		if (kernelLaunches) for (int i = 0; i < PSM_NUM_KERNEL_CLASSES; i++) kernelLaunches[i] = 0;
		if (stats3) { stats3[0] = stats3[1] = stats3[2] = 0; stats3[3] = smurfer.curvesMixedPairs; stats3[4] = smurfer.curvesHostSorts; stats3[5] = (uint64_t)std::llround(smurfer.curvesMaxSpread * 1e15); }
		for (uint32_t li = 0; li < dev.laneCount(); li++) {                  // kernel times / counts / byte statistics summed over the fold lanes
			Device* d = dev.lane(li);
			double ms[PSM_NUM_KERNEL_CLASSES]; uint64_t ln[PSM_NUM_KERNEL_CLASSES]; uint64_t st[3];
			d->check(psm_profile_get(d->ctx, ms, ln));
			d->check(psm_stats_get(d->ctx, &st[0], &st[1], &st[2]));
			for (int i = 0; i < PSM_NUM_KERNEL_CLASSES; i++) { if (kernelMs) kernelMs[i] += ms[i]; if (kernelLaunches) kernelLaunches[i] += ln[i]; }
			if (stats3) for (int i = 0; i < 3; i++) stats3[i] += st[i];
		}
		if (phases) for (int i = 0; i < PH_COUNT; i++) phases[i] = g_phase[i];
		return 0;
	} catch (const parsmurf::Error& e) {
		if (errbuf && errlen) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
		return e.code ? e.code : -1;
	} catch (const std::exception& e) {
		if (errbuf && errlen) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
		return -1;
	}
}

// Grid search through the host classes (cfgEx/gridTune.json: exec.optimizer = "grid").  grid = [G][6] {nParts, fp, ratio, k, nTrees,
// mtry}; gridMetrics [G][2] = {auroc, auprc} of the inner CVs of the last outer fold.  fold<k>.dat files are appended in the CWD.
int psmh_cv_run(const double* x, int x_is_device, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
                uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed, int device, void* stream,
                int tieMode, int evalFoldCurves, uint32_t rank, uint32_t world, psmh_reduce_fn reduce, void* reduceUser, uint64_t batchBudget,
                int profile, double* class1, double* class2, double* metrics, double* foldMetrics, double* kernelMs, uint64_t* kernelLaunches,
                uint64_t* stats3, double* phases, void* devHandle, char* errbuf, size_t errlen) {
	return run_mode(parsmurf::MODE_CV, nullptr, x, x_is_device, n, m, labels, foldAssign, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, device, stream,
	                tieMode, evalFoldCurves, rank, world, reduce, reduceUser, batchBudget, profile, class1, class2, metrics, foldMetrics, kernelMs,
	                kernelLaunches, stats3, phases, devHandle, errbuf, errlen);
}

// exec.mode "train" (mode 2: grow one forest per partition on all the data and save it to forestDir) and "predict"
// (mode 4: score every example with the saved forests); one fold, as ArgHandler_new.cpp forces for these modes.
int psmh_mode_run(int mode, const char* forestDir, const double* x, uint32_t n, uint32_t m, const uint32_t* labels, uint32_t nParts, uint32_t fp,
                  uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed, int device, int tieMode, uint32_t rank, uint32_t world,
                  psmh_reduce_fn reduce, void* reduceUser, double* class1, double* class2, double* metrics, char* errbuf, size_t errlen) {
	if (mode != parsmurf::MODE_TRAIN && mode != parsmurf::MODE_PREDICT) return PSM_E_ARG;
	// a single fold holding every example, generated like the reference's main does in these modes (the fold file is ignored,
	// ArgHandler_new.cpp:172-176, so Folds runs its random generation: src/folds.cpp:27-77)
	return run_mode(mode, forestDir, x, 0, n, m, labels, nullptr, 1, nParts, fp, ratio, k, nTrees, mtry, seed, device, nullptr, tieMode, 0, rank, world,
	                reduce, reduceUser, 0, 0, class1, class2, metrics, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, errbuf, errlen);
}

// psmh_cv_grid_run_ex: the same with the dataset optionally resident in HBM (x_is_device), a persistent device handle, the
// multi-rank sharding of psmh_cv_run (rank / world / reduce) and the launch statistics; psmh_cv_grid_run is the single-GPU form.
int psmh_cv_grid_run_ex(const double* x, int x_is_device, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
                        const uint32_t* grid, uint32_t G, uint32_t seed, int device, void* stream, int tieMode, uint32_t rank, uint32_t world,
                        psmh_reduce_fn reduce, void* reduceUser, double* class1, double* class2, double* metrics, double* gridMetrics,
                        uint64_t* kernelLaunches, uint64_t* stats5, void* devHandle, char* errbuf, size_t errlen) {
	using namespace parsmurf;
	try {
		std::unique_ptr<Device> owned;
		if (!devHandle) owned.reset(new Device(device, stream));
		Device& dev = devHandle ? *(Device*)devHandle : *owned;
		if (devHandle && stream) dev.check(psm_set_stream(dev.ctx, stream));
		const uint32_t lanes = std::min<uint32_t>(fold_lanes_default(), nFolds > 1 ? nFolds - 1 : 1);   // the inner CVs run their folds in lanes
		dev.setLaneCount(lanes);
		for (uint32_t li = 0; li < dev.laneCount(); li++) {
			Device* d = dev.lane(li);
			d->check(psm_profile_reset(d->ctx));
			d->check(psm_set_batch_budget(d->ctx, std::min<uint64_t>(24ull << 30, (120ull << 30) / lanes)));
			d->check(psm_profile_enable(d->ctx, 0));
		}
		GlibcRand rng(1);
		struct Pooled {                                     // the label copy lives in a recycled buffer too
			std::vector<uint32_t> v;
			Pooled(const uint32_t* src, size_t cnt) { if (src) { v = g_pool.take(cnt); v.assign(src, src + cnt); } }
			~Pooled() { g_pool.give(v); }
		} yBuf(labels, n);
		std::vector<uint32_t>& y = yBuf.v;
		Folds folds(nFolds, n, y.data(), foldAssign, n, foldAssign == nullptr, &rng);   // the fold ids are only borrowed for the call
		folds.computeFolds();
		CommonParams cp;
		cp.nn = n; cp.mm = m; cp.nFolds = nFolds; cp.seed = seed; cp.wmode = MODE_CV; cp.woptimiz = OPT_GRID;
		std::vector<GridParams> gp(G);
		for (uint32_t i = 0; i < G; i++) gp[i] = GridParams{grid[6 * i], grid[6 * i + 1], grid[6 * i + 2], grid[6 * i + 3], grid[6 * i + 4], grid[6 * i + 5], 0, 0};
		hyperSMURF smurfer(&cp, gp, &folds, &dev, y, class1, class2, &rng);
		smurfer.curvesTieMode = tieMode;
		smurfer.foldLanes = lanes;
		smurfer.shardUnits = shard_units_default();
		smurfer.deviceFolds = getenv("PSM_HOST_FOLDS") == nullptr;
		if (world > 1 && !reduce && !dev.hasComm()) throw Error(PSM_E_ARG, "multi-rank run without a reduce callback: initialise the device's communicator first (psmh_device_comm_init)");
		if (world > 1) smurfer.setSharding(rank, world, [&](double* p, uint64_t c) {
			if (!reduce) { dev.allReduce(p, c); return; }
			if (reduce(reduceUser, p, c)) throw Error(PSM_E_CUDA, "reduce callback failed");
		});
		smurfer.initStandard();
		if (x_is_device) dev.attachDevice(x, n, m); else dev.upload(x, n, m);
		smurfer.smurfIt();
		if (metrics) { metrics[0] = smurfer.auroc; metrics[1] = smurfer.auprc; }
		if (gridMetrics) for (uint32_t i = 0; i < G; i++) { gridMetrics[2 * i] = gp[i].auroc; gridMetrics[2 * i + 1] = gp[i].auprc; }
		if (kernelLaunches) for (int i = 0; i < PSM_NUM_KERNEL_CLASSES; i++) kernelLaunches[i] = 0;
		if (stats5) { stats5[0] = stats5[1] = stats5[2] = 0; stats5[3] = smurfer.curvesMixedPairs; stats5[4] = smurfer.curvesHostSorts; stats5[5] = (uint64_t)std::llround(smurfer.curvesMaxSpread * 1e15); }
		for (uint32_t li = 0; li < dev.laneCount(); li++) {
			Device* d = dev.lane(li);
			uint64_t ln[PSM_NUM_KERNEL_CLASSES]; uint64_t st[3];
			d->check(psm_profile_get(d->ctx, nullptr, ln));
			d->check(psm_stats_get(d->ctx, &st[0], &st[1], &st[2]));
			if (kernelLaunches) for (int i = 0; i < PSM_NUM_KERNEL_CLASSES; i++) kernelLaunches[i] += ln[i];
			if (stats5) for (int i = 0; i < 3; i++) stats5[i] += st[i];
		}
		return 0;
	} catch (const parsmurf::Error& e) {
		if (errbuf && errlen) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
		return e.code ? e.code : -1;
	} catch (const std::exception& e) {
		if (errbuf && errlen) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
		return -1;
	}
}
int psmh_cv_grid_run(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
                     const uint32_t* grid, uint32_t G, uint32_t seed, int device, int tieMode, double* class1, double* class2, double* metrics,
                     double* gridMetrics, char* errbuf, size_t errlen) {
	return psmh_cv_grid_run_ex(x, 0, n, m, labels, foldAssign, nFolds, grid, G, seed, device, nullptr, tieMode, 0, 1, nullptr, nullptr, class1, class2, metrics,
	                           gridMetrics, nullptr, nullptr, nullptr, errbuf, errlen);
}

// a persistent GPU context (device buffers are reused across psmh_cv_run calls)
// ForestFile::read followed by ForestFile::write (no GPU involved): lets the tests check the reader and the writer against a
// forest file written by the reference itself.
// generateRandomSet through the C boundary: x [n][m+1], y [n]
void psmh_generate_random_set(uint32_t n, uint32_t m, double prob, uint32_t seed, double* x, uint32_t* y) {
	std::vector<double> xx; std::vector<uint32_t> yy;
	parsmurf::generateRandomSet(n, m, xx, yy, prob, seed);
	memcpy(x, xx.data(), xx.size() * sizeof(double));
	memcpy(y, yy.data(), yy.size() * sizeof(uint32_t));
}
int psmh_forest_file_rewrite(const char* inPath, const char* outPath, uint64_t* nTreesOut, char* errbuf, size_t errlen) {
	try {
		parsmurf::ForestFile ff = parsmurf::ForestFile::read(inPath);
		if (nTreesOut) *nTreesOut = ff.trees.size();
		ff.write(outPath);
		return 0;
	} catch (const std::exception& e) {
		if (errbuf && errlen) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
		return -1;
	}
}

void* psmh_device_create(int device, void* stream) {
	try { return new parsmurf::Device(device, stream); } catch (...) { return nullptr; }
}
void psmh_device_destroy(void* h) { delete (parsmurf::Device*)h; }

// host-only helpers exposed for the CPU tests (no GPU needed)
void psmh_rand_fill(uint32_t seed, int32_t* out, uint64_t count) { parsmurf::GlibcRand g(seed); g.fill(out, count); }

// folds + ttd index arrays of one fold; sizes[4] always filled; index pointers may be NULL
int psmh_ttd_fold(const uint32_t* labels, const uint32_t* foldAssign, uint32_t n, uint32_t nFolds, uint32_t fold, int64_t foldToJump, uint32_t* sizes,
                  uint32_t* trngPos, uint32_t* trngNeg, uint32_t* testPos, uint32_t* testNeg) {
	using namespace parsmurf;
	try {
		GlibcRand rng(1);
		Folds folds(nFolds, n, labels, foldAssign, n, foldAssign == nullptr, &rng);   // borrowed ids, like psmh_cv_run
		folds.computeFolds();
		std::unique_ptr<TestTrainDivider> t(foldToJump < 0 ? new TestTrainDivider(&folds, MODE_CV, &rng) : new TestTrainDivider(&folds, MODE_CV, (uint32_t)foldToJump));
		if (fold >= t->nFolds) return PSM_E_ARG;
		t->ensure(fold);
		sizes[0] = t->trngPosNum[fold]; sizes[1] = t->trngNegNum[fold]; sizes[2] = t->testPosNum[fold]; sizes[3] = t->testNegNum[fold];
		if (trngPos) std::copy(t->trngPosIdx[fold].begin(), t->trngPosIdx[fold].end(), trngPos);
		if (trngNeg) std::copy(t->trngNegIdx[fold].begin(), t->trngNegIdx[fold].end(), trngNeg);
		if (testPos) std::copy(t->testPosIdx[fold].begin(), t->testPosIdx[fold].end(), testPos);
		if (testNeg) std::copy(t->testNegIdx[fold].begin(), t->testNegIdx[fold].end(), testNeg);
		return 0;
	} catch (...) { return PSM_E_ARG; }
}

// parse a cfgEx-style JSON; out[0..] = {nFolds, seed, wmode, woptimiz, gridSize, ensThrd, rfThrd, simulate, n, m}; grid [gridSize][6] if non-NULL
int psmh_config_parse(const char* json, uint32_t* out10, uint32_t* grid, uint32_t gridCap, char* errbuf, size_t errlen) {
	using namespace parsmurf;
	try {
		Config c = Config::fromString(json);
		out10[0] = c.common.nFolds; out10[1] = c.common.seed; out10[2] = c.common.wmode; out10[3] = c.common.woptimiz; out10[4] = (uint32_t)c.grid.size();
		out10[5] = c.common.nThr; out10[6] = c.common.rfThr; out10[7] = c.simulate; out10[8] = c.common.nn; out10[9] = c.common.mm;
		if (grid) for (size_t i = 0; i < c.grid.size() && i < gridCap; i++) {
			const GridParams& g = c.grid[i];
			uint32_t* r = grid + 6 * i;
			r[0] = g.nParts; r[1] = g.fp; r[2] = g.ratio; r[3] = g.k; r[4] = g.nTrees; r[5] = g.mtry;
		}
		return 0;
	} catch (const std::exception& e) {
		if (errbuf && errlen) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
		return PSM_E_ARG;
	}
}

}  // extern "C"
