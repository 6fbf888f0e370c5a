// TEST INFRASTRUCTURE -- not part of the product path.
//
// C-ABI harness around the UNMODIFIED reference classes (AnacletoLAB/parSMURF, compiled from
// /root/reference/src by oracle/Makefile into oracle/_ref/libparsmurf_ref.so).  It exists to
//   (1) pin oracle/psm_oracle.cpp (our CPU restatement) against the real reference,
//   (2) generate the golden fixtures under tests/golden/ (scripts/make_golden.py),
//   (3) serve as bench.py's `--impl reference` / cpu_baseline arm ("kind": "reference").
// Nothing here is reference source: it only *calls* the reference's public classes
//   Folds / TestTrainDivider / Partition / SamplerKNN / rfRanger / Curves / hyperSMURF
// in the order of src/hyperSMURF.cpp:257-346.
//
// rand() is interposed at link time (-Wl,--wrap=rand) so the SMOTE draws of
// src/samplerKNN.cpp:95,99 can be exported for injection into the CUDA path.

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <memory>
#include <sstream>
#include <iostream>
#include <omp.h>
#include <getopt.h>

#include "HyperSMURFUtils.h"
#include "folds.h"
#include "testtraindivider.h"
#include "partition.h"
#include "samplerKNN.h"
#include "curves.h"
#include "rfRanger.h"
#include "ForestProbability.h"
#include "DataDouble.h"
#include "hyperSMURF.h"
#include "fileImport.h"
#include "ANN.h"

// ---------------------------------------------------------------- rand() recorder
extern "C" int __real_rand(void);
static thread_local std::vector<int32_t>* g_rand_sink = nullptr;
extern "C" int __wrap_rand(void) {
	int r = __real_rand();
	if (g_rand_sink) g_rand_sink->push_back(r);
	return r;
}

namespace {

struct CoutSilencer {
	std::streambuf* old;
	std::ostringstream sink;
	CoutSilencer() { old = std::cout.rdbuf(sink.rdbuf()); }
	~CoutSilencer() { std::cout.rdbuf(old); }
};

struct RefCtx {
	uint32_t n, m, nFolds;
	std::vector<double> x;            // row-major [n][m+1], last col 1.0 pos / 2.0 neg
	std::vector<uint32_t> y;          // 0/1
	std::vector<uint32_t> ff;         // fold assignment
	std::unique_ptr<Folds> folds;
	std::unique_ptr<TestTrainDivider> ttd;
};

struct RefForest {
	std::unique_ptr<rfRanger> rf;
	uint32_t m, T, mtry, seed, rfThr;
};

std::vector<std::string> make_names(uint32_t m, const char* last) {
	std::vector<std::string> nomi = generateOrderedNames(m + 1);
	nomi[m] = last;
	return nomi;
}

}  // namespace

extern "C" {

// Build Folds (from a fold-assignment vector, src/folds.cpp:79-110) and TestTrainDivider
// (src/testtraindivider.cpp:5-106).  glibc's rand() state is reset with srand(1) first so the
// training-negative shuffle is the one a fresh `parSMURF1 --seed != 0` process performs.
void* ref_ctx_create(const double* x, uint32_t n, uint32_t m, const uint32_t* labels,
                     const uint32_t* foldAssign, uint32_t nFolds, int reset_rand) {
	RefCtx* c = new RefCtx;
	c->n = n; c->m = m; c->nFolds = nFolds;
	c->x.assign(x, x + (size_t)n * (m + 1));
	c->y.assign(labels, labels + n);
	c->ff.assign(foldAssign, foldAssign + n);
	if (reset_rand) srand(1);
	c->folds.reset(new Folds(nFolds, n, c->y.data(), c->ff, false));
	c->folds->computeFolds();
	c->ttd.reset(new TestTrainDivider(c->folds.get(), MODE_CV));
	return c;
}

void ref_ctx_destroy(void* h) { delete (RefCtx*)h; }

// sizes[4] = trngPosNum, trngNegNum, testPosNum, testNegNum
void ref_ctx_fold_sizes(void* h, uint32_t fold, uint32_t* sizes) {
	RefCtx* c = (RefCtx*)h;
	sizes[0] = c->ttd->trngPosNum[fold]; sizes[1] = c->ttd->trngNegNum[fold];
	sizes[2] = c->ttd->testPosNum[fold]; sizes[3] = c->ttd->testNegNum[fold];
}

void ref_ctx_fold_indices(void* h, uint32_t fold, uint32_t* trngPos, uint32_t* trngNeg,
                          uint32_t* testPos, uint32_t* testNeg) {
	RefCtx* c = (RefCtx*)h;
	std::memcpy(trngPos, c->ttd->trngPosIdx[fold], sizeof(uint32_t) * c->ttd->trngPosNum[fold]);
	std::memcpy(trngNeg, c->ttd->trngNegIdx[fold], sizeof(uint32_t) * c->ttd->trngNegNum[fold]);
	std::memcpy(testPos, c->ttd->testPosIdx[fold], sizeof(uint32_t) * c->ttd->testPosNum[fold]);
	std::memcpy(testNeg, c->ttd->testNegIdx[fold], sizeof(uint32_t) * c->ttd->testNegNum[fold]);
}

uint32_t ref_partition_maxsize(void* h, uint32_t nParts, uint32_t fp, uint32_t ratio) {
	RefCtx* c = (RefCtx*)h;
	Partition part(c->folds.get(), c->ttd.get(), nParts, fp, ratio, c->m, c->n);
	return part.maxSize;
}

// SamplerKNN::sample() for one (fold, partition).  Outputs the row-major training matrix, the
// sizes {trngPos, trngNeg, lineLen} and the rand() outputs it consumed (for injection).
// Returns the number of draws consumed, or -1 if a buffer is too small.
int64_t ref_sample_partition(void* h, uint32_t fold, uint32_t part_id, uint32_t nParts, uint32_t fp,
                             uint32_t ratio, uint32_t k, double* trngData, uint64_t trngCap,
                             uint32_t* sizes, int32_t* draws, uint64_t drawsCap) {
	RefCtx* c = (RefCtx*)h;
	Partition part(c->folds.get(), c->ttd.get(), nParts, fp, ratio, c->m, c->n);
	part.setFold(fold);
	SamplerKNN samp(&part, MODE_CV, c->x.data(), c->m, c->n, k);
	samp.setPartition(part_id);
	std::vector<int32_t> rec;
	g_rand_sink = &rec;
	samp.sample();
	g_rand_sink = nullptr;
	sizes[0] = samp.trngPos; sizes[1] = samp.trngNeg; sizes[2] = samp.lineLen;
	uint64_t need = (uint64_t)samp.lineLen * (c->m + 1);
	if (need > trngCap || rec.size() > drawsCap) return -1;
	std::memcpy(trngData, samp.trngData, need * sizeof(double));
	if (draws) std::memcpy(draws, rec.data(), rec.size() * sizeof(int32_t));
	return (int64_t)rec.size();
}

// Exact kNN as SamplerKNN::minOversample issues it (src/samplerKNN.cpp:60-93): ANNkd_tree over
// `P` points of dimension `dim`, annkSearch(q, k, idx, dist, eps=0); or ANNbruteForce.
int ref_knn(const double* pts, int P, int dim, int k, int32_t* nnIdx, double* nnDist, int brute) {
	ANNpointArray dataPts = annAllocPts(P, dim);
	for (int i = 0; i < P; i++) std::memcpy(dataPts[i], pts + (size_t)i * dim, sizeof(double) * dim);
	std::vector<ANNidx> idx(k);
	std::vector<ANNdist> dd(k);
	if (brute) {
		ANNbruteForce bf(dataPts, P, dim);
		for (int i = 0; i < P; i++) {
			bf.annkSearch(dataPts[i], k, idx.data(), dd.data(), 0);
			for (int j = 0; j < k; j++) { nnIdx[(size_t)i * k + j] = idx[j]; nnDist[(size_t)i * k + j] = dd[j]; }
		}
	} else {
		ANNkd_tree kd(dataPts, P, dim);
		for (int i = 0; i < P; i++) {
			kd.annkSearch(dataPts[i], k, idx.data(), dd.data(), 0);
			for (int j = 0; j < k; j++) { nnIdx[(size_t)i * k + j] = idx[j]; nnDist[(size_t)i * k + j] = dd[j]; }
		}
	}
	annDeallocPts(dataPts);
	annClose();
	return 0;
}

// the same kd-tree over all P points, but only the `nq` query points q[] are searched (large folds: a 1000-dimensional
// kd-tree search scans almost every point, so all P queries take minutes; the tree and the search code are the reference's)
int ref_knn_queries(const double* pts, int P, int dim, int k, const int32_t* q, int nq, int32_t* nnIdx, double* nnDist) {
	ANNpointArray dataPts = annAllocPts(P, dim);
	for (int i = 0; i < P; i++) std::memcpy(dataPts[i], pts + (size_t)i * dim, sizeof(double) * dim);
	std::vector<ANNidx> idx(k);
	std::vector<ANNdist> dd(k);
	{
		ANNkd_tree kd(dataPts, P, dim);
		for (int i = 0; i < nq; i++) {
			kd.annkSearch(dataPts[q[i]], k, idx.data(), dd.data(), 0);
			for (int j = 0; j < k; j++) { nnIdx[(size_t)i * k + j] = idx[j]; nnDist[(size_t)i * k + j] = dd[j]; }
		}
	}
	annDeallocPts(dataPts);
	annClose();
	return 0;
}

// rfRanger train ctor + train() on a column-major matrix [m+1][Ntr] (src/hyperSMURF.cpp:289-303).
void* ref_forest_train(const double* colmajor, uint32_t Ntr, uint32_t m, uint32_t T, uint32_t mtry,
                       uint32_t seed, uint32_t rfThr) {
	CoutSilencer quiet;
	std::vector<double> copy(colmajor, colmajor + (size_t)Ntr * (m + 1));
	std::unique_ptr<Data> input(new DataDouble(copy, make_names(m, "Labels"), Ntr, m + 1));
	RefForest* f = new RefForest;
	f->m = m; f->T = T; f->mtry = mtry; f->seed = seed; f->rfThr = rfThr;
	f->rf.reset(new rfRanger(m, false, std::move(input), T, mtry, rfThr, seed));
	f->rf->train(false);
	return f;
}

void ref_forest_destroy(void* h) { delete (RefForest*)h; }

uint64_t ref_forest_tree_nodes(void* h, uint32_t t) {
	RefForest* f = (RefForest*)h;
	return f->rf->forest->getSplitVarIDs()[t].size();
}

// ranger's layouts (SURVEY §7 "Tree export layout"): left/right child ids (0 = none), split var,
// split value, and per node {frac0, frac1} (NaN for internal nodes).
void ref_forest_export_tree(void* h, uint32_t t, uint64_t* left, uint64_t* right, uint64_t* var,
                            double* value, double* frac) {
	RefForest* f = (RefForest*)h;
	auto child = f->rf->forest->getChildNodeIDs()[t];
	auto vars = f->rf->forest->getSplitVarIDs()[t];
	auto vals = f->rf->forest->getSplitValues()[t];
	auto tcc = ((ForestProbability*)f->rf->forest)->getTerminalClassCounts()[t];
	size_t N = vars.size();
	for (size_t i = 0; i < N; i++) {
		left[i] = child[0][i]; right[i] = child[1][i]; var[i] = vars[i]; value[i] = vals[i];
		if (tcc[i].empty()) { frac[2 * i] = frac[2 * i + 1] = std::nan(""); }
		else { frac[2 * i] = tcc[i][0]; frac[2 * i + 1] = tcc[i].size() > 1 ? tcc[i][1] : 0.0; }
	}
}

// rfRanger predict ctor (deep copy of the trained forest) + predict() on a column-major test
// matrix [m+1][Nte] whose label column is 0 (src/hyperSMURF.cpp:316-333). preds = [Nte][2].
int ref_forest_predict(void* h, const double* colmajor, uint32_t Nte, double* preds) {
	RefForest* f = (RefForest*)h;
	CoutSilencer quiet;
	std::vector<double> copy(colmajor, colmajor + (size_t)Nte * (f->m + 1));
	std::unique_ptr<Data> test(new DataDouble(copy, make_names(f->m, "dependent"), Nte, f->m + 1));
	rfRanger rfTest(f->rf->forest, f->m, true, std::move(test), f->T, f->mtry, f->rfThr, 2 * f->seed);
	rfTest.predict(false);
	const auto& p = rfTest.forestPred->getPredictions();
	for (uint32_t i = 0; i < Nte; i++) { preds[2 * i] = p[0][i][0]; preds[2 * i + 1] = p[0][i].size() > 1 ? p[0][i][1] : 0.0; }
	return 0;
}

// rfRanger::saveForest (src/rfRanger.cpp:193-203 -> Forest::saveToFile, src/ranger/Forest/Forest.cpp:390-420):
// writes <dir>/<part>.out.forest in ranger's binary layout (train mode, src/hyperSMURF.cpp:349-360).
int ref_forest_save(void* h, uint32_t part, const char* dir) {
	RefForest* f = (RefForest*)h;
	CoutSilencer quiet;
	try { f->rf->saveForest(part, dir); } catch (...) { return -1; }
	return 0;
}

// predict mode (src/hyperSMURF.cpp:401-446): rfRanger(forestFilename, ...) loads the saved forest, predict() on a
// column-major test matrix [m+1][Nte] whose label column is 0.  preds = [Nte][2].
int ref_forest_file_predict(const char* forestFilename, const double* colmajor, uint32_t Nte, uint32_t m, uint32_t T,
                            uint32_t mtry, uint32_t seed, uint32_t rfThr, double* preds) {
	CoutSilencer quiet;
	std::vector<double> copy(colmajor, colmajor + (size_t)Nte * (m + 1));
	std::unique_ptr<Data> test(new DataDouble(copy, make_names(m, "dependent"), Nte, m + 1));
	rfRanger rfTest(std::string(forestFilename), m, true, std::move(test), T, mtry, rfThr, seed);
	rfTest.predict(false);
	const auto& p = rfTest.forestPred->getPredictions();
	if (p.empty() || p[0].size() != Nte) return -1;
	for (uint32_t i = 0; i < Nte; i++) { preds[2 * i] = p[0][i][0]; preds[2 * i + 1] = p[0][i].size() > 1 ? p[0][i][1] : 0.0; }
	return 0;
}

// Curves: AUROC first, then AUPRC (order matters, src/hyperSMURF.cpp:382-384). out = {auroc, auprc}.
void ref_curves(const uint32_t* labels, const double* preds, uint64_t N, double* out) {
	std::vector<uint32_t> l(labels, labels + N);
	Curves c(l, preds);
	out[0] = c.evalAUROC_ok();
	out[1] = c.evalAUPRC();
}

// The reference's own CV driver: Folds (from file vector) -> hyperSMURF::initStandard -> smurfIt
// (src/parSMURF.cpp:97-106).  class1/class2 must be zero-filled by the caller, length n.
// Returns wall seconds of smurfIt() (the reference's Timer region).
double ref_cv(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign,
              uint32_t nFolds, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees,
              uint32_t mtry, uint32_t seed, uint32_t nThr, uint32_t rfThr, int32_t minFold, int32_t maxFold,
              int reset_rand, double* class1, double* class2) {
	CoutSilencer quiet;
	std::vector<double> xx(x, x + (size_t)n * (m + 1));
	std::vector<uint32_t> yy(labels, labels + n);
	std::vector<uint32_t> ff(foldAssign, foldAssign + n);
	CommonParams cp;
	cp.nn = n; cp.mm = m; cp.nFolds = nFolds; cp.seed = seed; cp.verboseLevel = VERBSILENT;
	cp.nThr = nThr; cp.rfThr = rfThr; cp.wmode = MODE_CV; cp.woptimiz = OPT_NO; cp.rfVerbose = false;
	cp.verboseMPI = false; cp.noMtSender = false; cp.customCV = false;
	cp.minFold = (uint32_t)minFold; cp.maxFold = (uint32_t)maxFold;
	std::vector<GridParams> gp(1);
	gp[0].nParts = nParts; gp[0].fp = fp; gp[0].ratio = ratio; gp[0].k = k; gp[0].nTrees = nTrees; gp[0].mtry = mtry;
	gp[0].auroc = gp[0].auprc = 0;
	omp_set_num_threads(nThr);
	if (reset_rand) srand(1);
	Folds foldManager(nFolds, n, yy.data(), ff, false);
	foldManager.computeFolds();
	hyperSMURF smurfer(&cp, gp, &foldManager, xx, yy, class1, class2);
	smurfer.initStandard();
	Timer t; t.startTime();
	smurfer.smurfIt();
	t.endTime();
	return t.duration();
}

// exec.mode "train" exactly as parSMURF1's main drives it (src/parSMURF.cpp:97-106 with ArgHandler_new.cpp:153-176: one fold, the fold
// file ignored -> random fold generation, src/folds.cpp:27-77): Folds(1, n, y, ff, true) -> hyperSMURF::initStandard -> smurfIt, which
// writes <forestDir>/<p>.out.forest (+ .importance) for every partition.  Returns wall seconds of smurfIt.
double ref_mode_train(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k,
                      uint32_t nTrees, uint32_t mtry, uint32_t seed, uint32_t nThr, uint32_t rfThr, const char* forestDir, int reset_rand) {
	CoutSilencer quiet;
	std::vector<double> xx(x, x + (size_t)n * (m + 1));
	std::vector<uint32_t> yy(labels, labels + n);
	std::vector<uint32_t> ff(n, 0);
	double* class1 = new double[n](); double* class2 = new double[n]();
	CommonParams cp;
	cp.nn = n; cp.mm = m; cp.nFolds = 1; cp.seed = seed; cp.verboseLevel = VERBSILENT;
	cp.nThr = nThr; cp.rfThr = rfThr; cp.wmode = MODE_TRAIN; cp.woptimiz = OPT_NO; cp.rfVerbose = false;
	cp.verboseMPI = false; cp.noMtSender = false; cp.customCV = false;
	cp.minFold = (uint32_t)-1; cp.maxFold = (uint32_t)-1;
	cp.forestDirname = forestDir;
	std::vector<GridParams> gp(1);
	gp[0].nParts = nParts; gp[0].fp = fp; gp[0].ratio = ratio; gp[0].k = k; gp[0].nTrees = nTrees; gp[0].mtry = mtry;
	gp[0].auroc = gp[0].auprc = 0;
	omp_set_num_threads(nThr);
	if (reset_rand) srand(1);
	Folds foldManager(1, n, yy.data(), ff, true);
	foldManager.computeFolds();
	hyperSMURF smurfer(&cp, gp, &foldManager, xx, yy, class1, class2);
	smurfer.initStandard();
	Timer t; t.startTime();
	smurfer.smurfIt();
	t.endTime();
	delete[] class1; delete[] class2;
	return t.duration();
}

// Grid search (exec.optimizer = "grid", src/hyperSMURF.cpp:176-228): every outer fold runs an inner (nFolds-1)-fold CV per grid
// point, then the outer fold with the arg-max-AUPRC parameters.  The reference appends fold<k>.dat files to the CWD, so the
// caller should chdir to a scratch directory.  grid = [G][6] {nParts, fp, ratio, k, nTrees, mtry}; gridMetrics [G][2] receives
// {auroc, auprc} of the inner CVs of the LAST outer fold.
double ref_cv_grid(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
                   const uint32_t* grid, uint32_t G, uint32_t seed, uint32_t nThr, uint32_t rfThr, int reset_rand, double* class1,
                   double* class2, double* gridMetrics) {
	CoutSilencer quiet;
	std::vector<double> xx(x, x + (size_t)n * (m + 1));
	std::vector<uint32_t> yy(labels, labels + n);
	std::vector<uint32_t> ff(foldAssign, foldAssign + n);
	CommonParams cp;
	cp.nn = n; cp.mm = m; cp.nFolds = nFolds; cp.seed = seed; cp.verboseLevel = VERBSILENT;
	cp.nThr = nThr; cp.rfThr = rfThr; cp.wmode = MODE_CV; cp.woptimiz = OPT_GRID; cp.rfVerbose = false;
	cp.verboseMPI = false; cp.noMtSender = false; cp.customCV = false;
	cp.minFold = (uint32_t)-1; cp.maxFold = (uint32_t)-1;
	std::vector<GridParams> gp(G);
	for (uint32_t i = 0; i < G; i++) {
		gp[i].nParts = grid[6 * i]; gp[i].fp = grid[6 * i + 1]; gp[i].ratio = grid[6 * i + 2]; gp[i].k = grid[6 * i + 3];
		gp[i].nTrees = grid[6 * i + 4]; gp[i].mtry = grid[6 * i + 5]; gp[i].auroc = gp[i].auprc = 0;
	}
	omp_set_num_threads(nThr);
	if (reset_rand) srand(1);
	Folds foldManager(nFolds, n, yy.data(), ff, false);
	foldManager.computeFolds();
	hyperSMURF smurfer(&cp, gp, &foldManager, xx, yy, class1, class2);
	smurfer.initStandard();
	Timer t; t.startTime();
	smurfer.smurfIt();
	t.endTime();
	for (uint32_t i = 0; i < G; i++) { gridMetrics[2 * i] = gp[i].auroc; gridMetrics[2 * i + 1] = gp[i].auprc; }
	return t.duration();
}

// Bounded CPU-baseline sample: partitions [p0, p1) of one fold through the reference classes in the
// order of src/hyperSMURF.cpp:261-346, `nThr` OpenMP threads over partitions (as the reference does).
// Accumulates into class1/class2 (length n).  Returns wall seconds.
// `drawsOut` (optional): the rand() values each partition's SamplerKNN::sample() consumed, [p1-p0][perPartDraws] with
// perPartDraws = P*fp*(m+1) -- with several OpenMP threads the threads interleave on the one global rand() stream
// (SURVEY fact #5), so the draws a partition saw are only known by recording them (thread-local sink of the rand wrapper).
// `predsOut` (optional): the reference's raw forestPred->getPredictions() of each partition, [p1-p0][Nte][2], copied before
// anything is added to anything.  With both, the CUDA path can be given the same draws and compared bit for bit whatever
// the thread interleaving was (bench.py `parity`, tests/test_gpu_scale_parity.py).
double ref_run_partitions_ex(void* h, uint32_t fold, uint32_t p0, uint32_t p1, uint32_t nParts, uint32_t fp,
                             uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
                             uint32_t nThr, uint32_t rfThr, double* class1, double* class2, int32_t* drawsOut, uint64_t perPartDraws,
                             double* predsOut);
double ref_run_partitions(void* h, uint32_t fold, uint32_t p0, uint32_t p1, uint32_t nParts, uint32_t fp,
                          uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
                          uint32_t nThr, uint32_t rfThr, double* class1, double* class2) {
	return ref_run_partitions_ex(h, fold, p0, p1, nParts, fp, ratio, k, nTrees, mtry, seed, nThr, rfThr, class1, class2, nullptr, 0, nullptr);
}
double ref_run_partitions_ex(void* h, uint32_t fold, uint32_t p0, uint32_t p1, uint32_t nParts, uint32_t fp,
                             uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
                             uint32_t nThr, uint32_t rfThr, double* class1, double* class2, int32_t* drawsOut, uint64_t perPartDraws,
                             double* predsOut) {
	RefCtx* c = (RefCtx*)h;
	CoutSilencer quiet;
	Partition part(c->folds.get(), c->ttd.get(), nParts, fp, ratio, c->m, c->n);
	part.setFold(fold);
	omp_lock_t lock; omp_init_lock(&lock);
	omp_set_num_threads(nThr);
	const uint32_t mm = c->m;
	Timer t; t.startTime();
	#pragma omp parallel for schedule(dynamic, 1)
	for (int32_t cp = (int32_t)p0; cp < (int32_t)p1; cp++) {
		uint32_t seedCustom = seed + cp + fold * nParts;
		SamplerKNN samp(&part, MODE_CV, c->x.data(), mm, c->n, k);
		samp.setPartition(cp);
		std::vector<int32_t> sink;
		if (drawsOut) { sink.reserve(perPartDraws); g_rand_sink = &sink; }
		samp.sample();
		g_rand_sink = nullptr;
		if (drawsOut) {
			int32_t* dst = drawsOut + (size_t)(cp - (int32_t)p0) * perPartDraws;
			for (uint64_t i = 0; i < perPartDraws; i++) dst[i] = i < sink.size() ? sink[i] : -1;
			if (sink.size() != perPartDraws) dst[0] = -2;   // caller checks: draw count differs from P*fp*(m+1)
		}
		std::vector<double> trng((size_t)samp.lineLen * (mm + 1));
		transposeMatrix(trng.data(), samp.trngData, samp.lineLen, mm + 1);
		std::unique_ptr<Data> input(new DataDouble(trng, make_names(mm, "Labels"), samp.trngSize, mm + 1));
		rfRanger rf(mm, false, std::move(input), nTrees, mtry, rfThr, seedCustom);
		rf.train(false);
		samp.copyTestSet(c->x.data());
		std::vector<double> test((size_t)samp.testSize * (mm + 1));
		transposeMatrix(test.data(), samp.testData, samp.testSize, mm + 1);
		std::unique_ptr<Data> tdata(new DataDouble(test, make_names(mm, "dependent"), samp.testSize, mm + 1));
		rfRanger rfTest(rf.forest, mm, true, std::move(tdata), nTrees, mtry, rfThr, 2 * seedCustom);
		rfTest.predict(false);
		const auto& predictions = rfTest.forestPred->getPredictions();
		omp_set_lock(&lock);
		samp.accumulateAndDivideResInProbVect(predictions, class1, class2, nParts);
		omp_unset_lock(&lock);
		if (predsOut) {
			double* dst = predsOut + (size_t)(cp - (int32_t)p0) * samp.testSize * 2;
			for (uint32_t i = 0; i < samp.testSize; i++) { dst[2 * i] = predictions[0][i][0]; dst[2 * i + 1] = predictions[0][i][1]; }
		}
	}
	t.endTime();
	omp_destroy_lock(&lock);
	return t.duration();
}

int ref_num_procs(void) { return omp_get_num_procs(); }

// Importer::import (src/fileImport.cpp:5-100) on the three text files; sizes = {x.size(), y.size(), f.size(), nFolds}.
// Call once with x = NULL to get the sizes, then with buffers of those sizes.  (The reference's progress messages go to stdout.)
int ref_import(const char* dataFile, const char* labelFile, const char* foldFile, uint64_t* sizes, double* x, uint32_t* y, uint32_t* f) {
	try {
		std::vector<GridParams> gp;
		char prog[] = "ref"; char* argv[] = {prog, nullptr};
		ArgHandle ah(1, argv, gp);
		ah.dataFilename = dataFile; ah.labelFilename = labelFile; ah.foldFilename = foldFile ? foldFile : "";
		std::vector<double> xx; std::vector<uint32_t> yy, ff;
		uint32_t nFolds = 0;
		std::streambuf* old = std::cout.rdbuf(nullptr);          // silence the importer's chatter
		try { Importer::import(&ah, xx, yy, ff, &nFolds); } catch (...) { std::cout.rdbuf(old); throw; }
		std::cout.rdbuf(old);
		sizes[0] = xx.size(); sizes[1] = yy.size(); sizes[2] = ff.size(); sizes[3] = nFolds;
		if (x) std::memcpy(x, xx.data(), xx.size() * sizeof(double));
		if (y) std::memcpy(y, yy.data(), yy.size() * sizeof(uint32_t));
		if (f) std::memcpy(f, ff.data(), ff.size() * sizeof(uint32_t));
		return 0;
	} catch (...) { return -1; }
}

// ArgHandle::processCommandLine on `--cfg path` (src/ArgHandler_new.cpp:19-82: jsonImport + checkCommonConfig + checkConfig).
// out10 = {nFolds, seed, wmode, woptimiz, gridSize, ensThrd, rfThrd, simulate, n, m}; grid [gridSize][6] = {nParts, fp, ratio, k,
// nTrees, mtry}.  The reference exit()s on an invalid configuration, so tests call this in a child process.
int ref_config_parse(const char* path, uint32_t* out10, uint32_t* grid, uint32_t gridCap) {
	CoutSilencer quiet;
	std::vector<GridParams> gp;
	char prog[] = "ref", opt[] = "--cfg";
	std::string p(path);
	char* argv[] = {prog, opt, &p[0], nullptr};
	optind = 1;
	ArgHandle ah(3, argv, gp);
	ah.processCommandLine(1);     // rank 1: no logo / messages
	out10[0] = ah.nFolds; out10[1] = ah.seed; out10[2] = ah.wmode; out10[3] = ah.woptimiz; out10[4] = (uint32_t)gp.size();
	out10[5] = ah.ensThreads; out10[6] = ah.rfThreads; out10[7] = ah.simulate; out10[8] = ah.n; out10[9] = ah.m;
	for (size_t i = 0; i < gp.size() && i < gridCap; i++) {
		uint32_t* r = grid + 6 * i;
		r[0] = gp[i].nParts; r[1] = gp[i].fp; r[2] = gp[i].ratio; r[3] = gp[i].k; r[4] = gp[i].nTrees; r[5] = gp[i].mtry;
	}
	return 0;
}

// saveToFile (src/HyperSMURFUtils.cpp:66-86); labels == NULL -> the two-column form main() writes for data read from files
void ref_save(const double* c1, const double* c2, uint32_t n, const uint32_t* labels, const char* path) {
	std::vector<uint32_t> lab;
	if (labels) lab.assign(labels, labels + n);
	saveToFile(c1, c2, n, &lab, path);
}

// generateRandomSet (src/HyperSMURFUtils.cpp:40-63), the simulated data set of cfgEx/simulCV.json: x row-major [n][m+1], y [n]
void ref_generate_random_set(uint32_t n, uint32_t m, double prob, uint32_t seed, double* x, uint32_t* y) {
	std::vector<double> xx((size_t)n * (m + 1));
	std::vector<uint32_t> yy(n);
	generateRandomSet(n, m, xx, yy, prob, seed);
	memcpy(x, xx.data(), xx.size() * sizeof(double));
	memcpy(y, yy.data(), yy.size() * sizeof(uint32_t));
}

}  // extern "C"
