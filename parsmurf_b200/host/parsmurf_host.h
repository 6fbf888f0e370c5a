// parsmurf_b200 host layer -- C++ mirror of the reference's class boundaries for the partition pipeline.
//
// The reference (AnacletoLAB/parSMURF) has no FFI: hyperSMURF::smurfIt calls Partition, SamplerKNN, rfRanger
// and Curves directly (reference src/hyperSMURF.cpp:255-396).  These classes keep those names, constructor
// arguments, public fields and call order, but their bodies forward to the C ABI in include/parsmurf_b200.h
// (CUDA, sm_100a).  Differences a maintainer must know:
//   * `x` is uploaded once into a `Device` (the GPU context); samplers borrow the Device instead of `const double* x`.
//   * partition concurrency is a RANGE [p0,p1) handled by batched launches (setPartitionRange) instead of one
//     OpenMP thread per partition; setPartition(p) still works (range of one).
//   * rand(): the reference draws from the process-global glibc rand(), never seeded for seed != 0
//     (src/ArgHandler_new.cpp:257-262).  The host owns a private replica of that generator (GlibcRand, seed 1),
//     so a run equals a fresh single-threaded (ensThrd 1) reference process bit for bit and stays deterministic.
//   * errors throw parsmurf::Error (the reference abort()s or swallows exceptions, src/samplerKNN.cpp:38-48,
//     src/rfRanger.cpp:51-53); Error::code is the PSM_E_* value.
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/parsmurf_b200.h"

namespace parsmurf {

// reference src/HyperSMURFUtils.h:17-68
enum verbLvl { VERBSILENT = 0, VERBPROGR = 1, VERBRF = 2, VERBALL = 3 };
enum wmode_t { MODE_CV = 1, MODE_TRAIN = 2, MODE_PREDICT = 4 };
enum woptimizer { OPT_NO = 16, OPT_GRID = 32, OPT_AUTOGP = 64 };

struct GridParams {
	uint32_t nParts, fp, ratio, k, nTrees, mtry;
	double auroc, auprc;
};

struct CommonParams {
	uint32_t nn = 0, mm = 0, nFolds = 0, seed = 1, verboseLevel = 0;
	std::string outfilename, timeFilename, forestDirname, cfgFilename;
	uint32_t nThr = 1, rfThr = 1;
	uint8_t wmode = MODE_CV, woptimiz = OPT_NO;
	bool rfVerbose = false, verboseMPI = false, noMtSender = false, customCV = false;
	uint32_t minFold = (uint32_t)-1, maxFold = (uint32_t)-1;
};

// wall-clock phase accounting of the host layer (seconds since the last reset), for bench.py / profiling
enum Phase { PH_SETUP = 0, PH_FOLDS_TTD, PH_FOLD_BEGIN, PH_DRAWS, PH_ASSEMBLE, PH_TRAIN, PH_PREDICT, PH_REDUCE, PH_SCATTER, PH_CURVES, PH_TEARDOWN, PH_COUNT };
extern double g_phase[PH_COUNT];
struct PhaseScope {
	int ph; double t0;
	explicit PhaseScope(int p);
	~PhaseScope();
};

struct Error : std::runtime_error {
	int code;
	Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// glibc rand()/srand() replica (stdlib/random_r.c TYPE_3): the stream behind std::random_shuffle
// (src/folds.cpp:41-42, src/testtraindivider.cpp:90) and SMOTE (src/samplerKNN.cpp:95,99)
class GlibcRand {
public:
	explicit GlibcRand(uint32_t seed = 1) { srand(seed); }
	void srand(uint32_t seed);
	int32_t rand();
	void fill(int32_t* out, size_t count);
	void shuffle(uint32_t* a, size_t n);   // libstdc++ std::random_shuffle(first, last)
	void window(uint32_t* out31) const;    // last 31 state words, oldest first (what psm_batch_assemble_glibc takes)
	void discard(uint64_t n);              // skip n draws in O(log n) (linear recurrence jump-ahead)
private:
	int32_t tbl[31];
	int f, r;
};

// reference src/folds.h:10-40
class Device;
class Folds {
public:
	Folds(uint32_t numberOfFolds, uint32_t classSize, const uint32_t* labels, std::vector<uint32_t>& ff, bool randomFold, GlibcRand* rng);
	// the same with a borrowed fold-id array (no copy of n ids); `ids` must outlive the object
	Folds(uint32_t numberOfFolds, uint32_t classSize, const uint32_t* labels, const uint32_t* ids, size_t nIds, bool randomFold, GlibcRand* rng);
	~Folds();
	Folds(const Folds&) = delete;
	void computeFolds();
	// "Folds from a file" on the device (psm_folds_build_device): the per-class lists are built in HBM from the fold ids by one
	// stable sort and never exist on the host unless somebody asks (hostLists()); only the fold sizes come back.  `dev` must
	// hold the dataset and the labels.  Random folds (no fold file) keep the host path: they are a rand()-driven shuffle.
	void computeFoldsOnDevice(Device* dev);
	const Device* builtOn() const { return devBuilt; }
	void hostLists() const;                // make posList / negList valid on the host (a D2H copy after computeFoldsOnDevice)
	const uint32_t classSize, nFolds;
	const uint32_t* labels;
	std::vector<uint32_t>& fromFile;
	// `folds` (every fold's examples, src/folds.h) is only filled on request when the assignment comes from a file: nothing on
	// the GPU path reads it (the per-class lists below are what the train / test arrays are assembled from)
	bool keepFoldsArray = false;
	uint32_t totPos = 0, totNeg = 0;
	std::vector<uint32_t> nPos, nNeg, foldsIdx, folds;
	uint32_t maxPos = 0, maxNeg = 0;
	// new: folds[] split by class, fold by fold in the same order (positives of fold f at posList[posOff[f] .. posOff[f+1])),
	// so TestTrainDivider assembles its index arrays by block copies instead of re-reading the labels per fold
	std::vector<uint32_t> posOff, negOff, posList, negList;
	bool randomFold;
	GlibcRand* rng;
private:
	Device* devBuilt = nullptr;            // computeFoldsOnDevice: where the lists live
	mutable bool hostListsValid = false;
	mutable std::mutex hostListsMtx;
	std::vector<uint32_t> noIds;           // what `fromFile` refers to when the ids are borrowed
	const uint32_t* ids = nullptr;         // fold id per example (fromFile.data() or the borrowed array)
	size_t nIds = 0;
};

// reference src/testtraindivider.h:13-36 (both constructors)
class TestTrainDivider {
public:
	TestTrainDivider(const Folds* folds, uint8_t wmode, GlibcRand* rng);
	TestTrainDivider(const Folds* folds, uint8_t wmode, uint32_t foldToJump);
	~TestTrainDivider();
	// The counts are valid immediately; the index arrays of fold f after ensure(f) (test + shuffled training arrays) or
	// ensureTest(f) (test arrays only).  A fold is built by whoever asks first -- the calling thread (fold lanes build their
	// folds concurrently) or the prefetch thread, which walks the folds given to prefetch() in order, one fold ahead of a
	// caller that consumes them one after the other.  Only folds somebody asks for are ever built: a rank of a unit-sharded
	// run touches one or two of them.
	void ensure(uint32_t fold) const;
	void ensureTest(uint32_t fold) const;
	void prefetch(std::vector<uint32_t> foldsInOrder);
	bool builtOnDemand() const { return lazy; }            // outer-CV constructor (folds can also be built on the device)
	void shuffleWindow(uint32_t fold, uint32_t* w31) const;   // rand() state (psm_fold_begin_device layout) before fold's shuffle
	const Folds* const folds;
	const uint8_t wmode;
	uint32_t nFolds;
	std::vector<std::vector<uint32_t>> testPosIdx, testNegIdx, trngPosIdx, trngNegIdx;
	std::vector<uint32_t> testPosNum, testNegNum, trngPosNum, trngNegNum;
private:
	void buildTest(uint32_t i, uint32_t currFoldIdx);
	void buildTrain(uint32_t i, uint32_t currFoldIdx, int64_t foldToJump);
	std::thread worker;                       // prefetch
	mutable std::mutex mtx;
	mutable std::condition_variable cv;
	mutable std::vector<uint8_t> testState, trainState;   // per fold: 0 not built, 1 being built, 2 ready (guarded by mtx)
	bool lazy = false;                        // outer-CV constructor: folds are built on demand
	GlibcRand start;                          // rand() state before the first fold's shuffle
	std::vector<uint64_t> shuffleOff;         // where fold i's shuffle starts in that stream
};

// reference src/partition.h:7-36
class Partition {
public:
	Partition(const Folds* folds, const TestTrainDivider* ttd, uint32_t nPart, uint32_t fp, uint32_t ratio, uint32_t m, uint32_t n);
	void setFold(uint32_t currentFold);
	void setFoldCounts(uint32_t currentFold);              // counts only: the index arrays live on the device (Device::beginFoldOnDevice)
	const Folds* const folds;
	const TestTrainDivider* const ttd;
	uint32_t nPart, fp, ratio;
	const uint32_t m, n;
	uint32_t maxSize;
	uint32_t currentFold = 0;
	uint32_t trngPosNum = 0, trngNegNum = 0, testPosNum = 0, testNegNum = 0;
	const uint32_t *testPosIdx = nullptr, *testNegIdx = nullptr, *trngPosIdx = nullptr, *trngNegIdx = nullptr;
};

// The GPU context: owns psm_ctx and the device-resident copy of x (new; the reference shares `x` between threads)
class Device {
public:
	Device(int device, void* cudaStream = nullptr);
	~Device();
	Device(const Device&) = delete;
	void upload(const double* x, uint32_t n, uint32_t m);
	bool shardedUpload = true;      // ranks of one communicator upload a slice of x each and all-gather it (Device::commInit first)
	void attachDevice(const double* x_dev, uint32_t n, uint32_t m);
	// begins the fold on the device if it is not the current one (upload index arrays, zero fold scores, drop kNN)
	void ensureFold(const Partition* part);
	// the same from the device-resident fold lists (psm_folds_upload / psm_fold_begin_device): nothing is built on the host
	void uploadFolds(const Folds* folds);
	void attachFolds(const Device* src);
	void beginFoldOnDevice(const Partition* part);
	void ensureKnn(uint32_t k);
	void invalidateFold() { foldTag = nullptr; }
	void check(int rc) const;
	// cross-rank communicator of this GPU (NCCL inside the library, psm_comm_init): id128 from commUniqueId() of one rank,
	// handed to the others by the host's own bootstrap.  allReduce sums `count` doubles at `dev_ptr` over the ranks, in place.
	static void commUniqueId(void* id128);
	void commInit(const void* id128, uint32_t rank, uint32_t world);
	bool hasComm() const;
	void allReduce(double* dev_ptr, uint64_t count);
	// fold lanes: lane 0 is this context; lanes 1.. are extra contexts on the same GPU with private streams, created on
	// demand and kept for the life of the Device (their buffers are reused like this one's)
	void setLaneCount(uint32_t lanes);
	uint32_t laneCount() const { return 1 + (uint32_t)extra.size(); }
	Device* lane(uint32_t i) { return i == 0 ? this : extra[i - 1].get(); }
	psm_ctx* ctx = nullptr;
	int deviceId = 0;
	std::vector<std::unique_ptr<Device>> extra;
	uint32_t n = 0, m = 0;
	const void* foldTag = nullptr;
	uint32_t foldId = (uint32_t)-1, knnK = 0;
};

// reference src/sampler.h:11-73 + src/samplerKNN.h:13-35
class SamplerKNN {
public:
	SamplerKNN(const Partition* part, uint8_t wmode, Device* dev, uint32_t m, uint32_t n, uint32_t k, GlibcRand* rng);
	void setPartition(uint32_t currentPart) { setPartitionRange(currentPart, currentPart + 1); }
	void setPartitionRange(uint32_t p0, uint32_t p1);
	void sample();                                  // setup + minOversample + majUndersample for the whole range, on the device
	void copyTestSet() {}                           // test rows are gathered by index inside the predict kernel
	// scores are accumulated on the device by rfRanger::predict(); kept for call-order compatibility
	void accumulateAndDivideResInProbVect(uint32_t /*nPart*/) {}
	std::vector<double> trngData(uint32_t p);       // row-major [trngSize][m+1] host copy (debug / parity)
	bool deviceDraws = true;                        // replay rand() on the GPU instead of generating the draws on the host
	const Partition* const part;
	Device* const dev;
	const uint32_t m, n;
	const uint8_t wmode;
	uint32_t currentPartition = 0, currentFold = 0, p0 = 0, p1 = 0;
	uint32_t posForOverSmpl = 0, negForUndrSmpl = 0;
	uint32_t trngPos = 0, trngNeg = 0, trngSize = 0, lineLen = 0, testSize = 0;   // of partition p0
private:
	uint32_t k;
	GlibcRand* rng;
};

// reference src/rfRanger.h:18-67 -- ranger ForestProbability with parSMURF's fixed hyper-parameters
class rfRanger {
public:
	rfRanger(Device* dev, uint32_t m, bool prediction_mode, const SamplerKNN* samp, uint32_t numTrees, uint32_t mtry, uint32_t rfThrd, uint32_t seedBase);
	// predict mode (src/rfRanger.cpp:120-168): the forests of partitions [p0, p1) are read from <forestDirname>/<p>.out.forest
	// and installed on the device; predict() then accumulates pred * (1/nPartTotal) like the in-memory constructor
	rfRanger(Device* dev, const std::string& forestDirname, uint32_t p0, uint32_t p1, uint32_t nPartTotal, uint32_t m, uint32_t numTrees,
	         uint32_t mtry, uint32_t rfThrd, uint32_t seed);
	void train(bool verbose);     // forest seed of partition p = seedBase + p  (seedBase = seed + fold*nPart, src/hyperSMURF.cpp:273)
	void predict(bool verbose);   // predict the fold's test rows and accumulate pred * (1/nPart) into the fold score buffer
	// ranger getter layouts (src/ranger/Forest/Forest.h:82-101, ForestProbability.h:40-44)
	struct TreeExport { std::vector<uint64_t> left, right, splitVarIDs; std::vector<double> splitValues, terminalClassCounts; };
	TreeExport getTree(uint32_t p, uint32_t t) const;
	// train mode (src/rfRanger.cpp:193-203): writes <forestDirname>/<nPart>.out.forest in ranger's binary forest layout,
	// byte for byte what Forest::saveToFile produces for the same forest.  (saveImportance has no counterpart: permutation
	// importance is not computed on the GPU path)
	void saveForest(uint32_t nPart, const std::string& forestDirname) const;
	Device* const dev;
	const SamplerKNN* const samp;
	uint32_t mtry, rfSeed, min_node_size = 10, num_threads, numTrees;
	bool prediction_mode;
private:
	uint32_t m_ = 0, nPartTotal_ = 0;
};

// ranger's binary forest file for a probability forest (src/ranger/Forest/Forest.cpp:390-420, ForestProbability.cpp:262-328,
// Tree.cpp:234-243, TreeProbability.cpp:65-79, utility.h:58-140): dependent_varID, num_trees, is_ordered_variable,
// num_variables, treetype 9, class_values, then per tree child_nodeIDs[2], split_varIDs, split_values, terminal node ids and
// their class fractions.  All integers are 64-bit little endian, vectors are length-prefixed.
struct ForestFile {
	uint64_t dependentVarID = 0, numVariables = 0;
	std::vector<double> classValues;
	std::vector<rfRanger::TreeExport> trees;
	void write(const std::string& path) const;
	static ForestFile read(const std::string& path);
};

// The simulated data set of cfgEx/simulCV.json ("simulate" block): labels Bernoulli(prob), features N(5,1) for positives and
// N(0,1) for negatives, column m = 1.0 / 2.0 -- the same engine (std::default_random_engine), distribution objects and draw
// order as generateRandomSet (src/HyperSMURFUtils.cpp:40-63), so with the same libstdc++ the set is the reference's.
void generateRandomSet(uint32_t nn, uint32_t mm, std::vector<double>& xx, std::vector<uint32_t>& yy, double prob, uint32_t seed);

// reference src/curves.h:12-21
class Curves {
public:
	Curves(Device* dev, const std::vector<uint32_t>& labels, const double* predictions, int tieMode = 0);
	double evalAUROC_ok();
	double evalAUPRC();
private:
	void eval();
	Device* dev;
	std::vector<uint32_t> labls;
	const double* preds;
	int tieMode;
	bool done = false;
	double res[2] = {0, 0};
};

// reference src/hyperSMURF.h:24-79
class hyperSMURF {
public:
	hyperSMURF(const CommonParams* commonParams, std::vector<GridParams>& gridParams, Folds* foldManager, Device* dev,
	           const std::vector<uint32_t>& y, double* class1Prob, double* class2Prob, GlibcRand* rng);
	~hyperSMURF();
	void initStandard();
	void initInternalCV(uint32_t foldToJump);
	void setGridParams(uint32_t idx);
	void createPart();
	void smurfIt();     // modes "cv" and "train" (src/hyperSMURF.cpp:151-396); mode "predict" -> predictIt (:401-446)

	// multi-GPU: this rank handles partitions [rank*.., ..) of every fold (src/MPIHelper.cpp:29-37) and `reduce`
	// sums the fold score buffer [2*Nte] f64 across ranks in place (the MPI_Gather + master sum of
	// src/hyperSMURF_MPI.cpp:594-618).  Default: single rank, no reduce.
	void setSharding(uint32_t rank, uint32_t world, std::function<void(double* dev_ptr, uint64_t count)> reduce);
	// How the work of a plain CV is cut between the ranks.  false: the reference's MPI rule -- every rank takes its
	// contiguous chunk of the partitions of EVERY fold and the fold scores are summed once per fold.  true (unit sharding):
	// the nFolds x nParts (fold, partition) units are cut into `world` contiguous ranges, so a rank opens only the one or two
	// folds it has units of (their index upload, kNN and fold set-up are not repeated on every rank, and its partition
	// batches are nFolds times larger); the class1Prob/class2Prob arrays are summed ONCE at the end and the per-fold curves
	// are computed from the summed arrays, fold f on rank f % world.  Same sums up to the fp64 addition order across ranks.
	// A rank's share of a fold is further cut into pieces of at most max(32, share / foldLanes) partitions, one piece per
	// fold lane at a time, so that a rank that owns only one or two folds still keeps several lanes busy.
	bool shardUnits = false;
	// Grid search (exec.optimizer "grid", src/hyperSMURF.cpp:176-228) on several ranks: true -- every rank runs whole inner CVs,
	// its cost-balanced share of the grid points of each outer fold, and the [G][2] metric table is summed once per outer fold;
	// false -- every inner fold's partitions are chunked over the ranks (the MPI rule), one reduce per inner fold.
	bool shardGridPoints = true;
	// [first, last) of fold `fold`'s partitions that belong to `rank` under either rule (src/MPIHelper.cpp:29-37 for the first)
	static void partRange(bool units, uint32_t nFoldsRun, uint32_t nPart, uint32_t world, uint32_t rank, uint32_t foldIdx, uint32_t* first, uint32_t* last);
	// grid search on several ranks: the rank that runs the inner CV of each grid point (cost-balanced, the same on every rank)
	static std::vector<uint32_t> gridOwners(const std::vector<GridParams>& grid, uint32_t world);
	// unit sharding with fold lanes: how many pieces each fold share (partitions of fold f owned by this rank) is cut into
	static std::vector<uint32_t> laneCuts(const std::vector<uint32_t>& share, uint32_t lanes);
	int curvesTieMode = 0;
	bool evalFoldCurves = true;
	// Folds of a CV are independent until the score merge: with foldLanes > 1 that many folds are in flight at once, each
	// on its own Device lane (context + stream) driven by its own host thread, so one fold's host-side work and its
	// latency-bound deep tree levels overlap the other folds' kernels.  Results are identical to foldLanes = 1: the
	// SMOTE rand() position of every fold is known up front (jump-ahead) and cross-rank reduces stay in fold order.
	uint32_t foldLanes = 1;
	// Build every fold's index arrays (and shuffle its training negatives) on the device instead of on the host
	// (TestTrainDivider); applies to plain and grid-outer CV folds, the inner CVs and train/predict modes keep the host path.
	bool deviceFolds = true;
	// class1Prob / class2Prob are pure outputs of this run (nothing to accumulate onto): the final fetch overwrites them, so
	// the caller need not zero them (the reference's main() zero-fills and every partition adds, src/parSMURF.cpp:90-93)
	bool scoresAreOutputs = false;
	// multi-rank runs: only rank 0 copies the final class1Prob / class2Prob to its host arrays, like the master of the
	// reference's MPI build, which alone holds and writes the predictions (src/hyperSMURF_MPI.cpp:594-646, src/parSMURFn.cpp);
	// the other ranks' arrays are left untouched.  false: every rank fetches them (2n doubles over PCIe per rank).
	bool scoresOnRootOnly = false;
	std::vector<double> foldAuroc, foldAuprc;
	double auroc = 0, auprc = 0;
	// curvesTieMode 2 ("auto", psm_curves_info): over all the curves calls of this run, the adjacent (equal score, different
	// label) pairs the device sort met, and how many calls had to be redone with the host std::sort
	std::atomic<uint64_t> curvesMixedPairs{0}, curvesHostSorts{0};
	double curvesMaxSpread = 0;      // largest |max - min| AUROC / AUPRC over the tie orders, over all the calls that had mixed ties
	std::mutex spreadMtx;

private:
	struct FoldGate;
	void predictIt();
	void noteCurves(Device* d);
	bool willShardUnits() const;
	bool foldsOnDevice() const;
	void runFold(uint32_t currentFold, uint32_t first, uint32_t last, Device* d, Partition* prt, GlibcRand* r, FoldGate* gate);
	bool unitMode = false;      // this smurfIt call shards by (fold, partition) unit
	uint32_t foldBase = 0, foldsRun = 0;   // first fold and number of folds of this smurfIt call
	void runFoldsInLanes(uint32_t startingFold, uint32_t endingFold, uint32_t lanes);
	const CommonParams* const commonParams;
	std::vector<GridParams>& gridParams;
	Folds* const foldManager;
	Device* const dev;
	const std::vector<uint32_t>& y;
	double* const class1Prob;
	double* const class2Prob;
	GlibcRand* rng;
	std::unique_ptr<TestTrainDivider> ttd;
	std::unique_ptr<Partition> part;
	uint32_t nn = 0, mm = 0, nFolds = 0, nPart = 0, numTrees = 0, fp = 0, ratio = 0, k = 0, mtry = 0, seed = 0, verboseLevel = 0;
	uint8_t wmode = MODE_CV, woptimiz = OPT_NO;
	bool inInternalCV = false;
	uint32_t idxInGridParams = 0, foldToJump = 0;
	uint32_t rank = 0, world = 1;
	std::function<void(double*, uint64_t)> reduce;
	std::vector<double> tempcl1, tempcl2;
};

// reference src/fileImport.h:14-23 (Importer::import) and src/HyperSMURFUtils.h (saveToFile) -- ingest.cpp
struct Importer {
	// x comes back row-major [n][m+1] with the label column appended (1.0 for label > 0, else 2.0), y / f as read; *nFolds is only
	// written when a fold file is given (max id + 1).  threads == 0: as many as the file is worth (<= 16).  A data file that starts
	// with the magic "PSMB1" is read as the binary side channel (labels / folds inside; the other two names are ignored).
	static void import(const std::string& dataFilename, const std::string& labelFilename, const std::string& foldFilename, std::vector<double>& x,
	                   std::vector<uint32_t>& y, std::vector<uint32_t>& f, uint32_t* nFolds, uint32_t* m, unsigned threads = 0);
	static bool isBinary(const std::string& dataFilename);
	static void importBinary(const std::string& path, std::vector<double>& x, std::vector<uint32_t>& y, std::vector<uint32_t>& f, uint32_t* nFolds, uint32_t* m);
	static void exportBinary(const std::string& path, const double* xRowMajorWithLabel, uint32_t n, uint32_t m, const uint32_t* y, const uint32_t* folds);
};
// "cl1<TAB>cl2[<TAB>label]" per example, byte-identical to the reference's text output; a name ending in ".bin" gets raw f64 arrays
void saveToFile(const double* cl1, const double* cl2, uint32_t nn, const std::vector<uint32_t>* labels, const std::string& outFilename, unsigned threads = 0);

// JSON configuration surface of cfgEx/*.json (reference src/ArgHandler_new.cpp:84-141, 294-375)
struct Config {
	CommonParams common;
	std::vector<GridParams> grid;
	std::string dataFilename, foldFilename, labelFilename;
	bool simulate = false;
	double prob = 0.0;
	bool printCfg = false;
	bool seedFromClock = false;      // no exec.seed (or 0): seed = time(NULL) and the caller srand()s it (ArgHandler_new.cpp:257-262)
	static Config fromFile(const std::string& path);
	static Config fromString(const std::string& json);
	void processMtry(uint32_t mm);   // mtry 0 or > m -> (uint32_t)sqrt(m)
};

}  // namespace parsmurf
