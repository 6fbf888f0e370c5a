// parsmurf_b200 -- kernel launcher declarations shared by the .cu translation units and api.cu
#pragma once
#include "common.cuh"

namespace psm {

// counters[CNT_NEXT_ITEMS] = open nodes of the next level; counters[CNT_CLS0 + k] = how many of them fall in size class k
// (level_finalize appends their item ids to clsItems[next][k]); both are cleared by the host before the next level_finalize
constexpr int NCLS = 6;
enum { CNT_NEXT_ITEMS = 0, CNT_CLS0 = 1, CNT_ERROR = 7, CNT_COUNT = 8 };
// Size class of a work item by its segment length (+3: worst-case shift of the 16-byte aligned load grid):
//   0: <= 32 grid entries (8 lanes x 4)   1: <= 64 (16 lanes)   2: <= 128 (one warp, one chunk)
//   3, 4, 5: longer (one warp, chunk loop), cut at SS_CLS4_MIN and SS_CLS5_MIN entries only so that the merged split-search launch
//   can start the longest segments first (train.cu: split_search_merged_kernel)
#ifndef SS_CLS4_MIN
#define SS_CLS4_MIN 510u
#endif
#ifndef SS_CLS5_MIN
#define SS_CLS5_MIN 2046u
#endif
__host__ __device__ __forceinline__ uint32_t size_class(uint32_t cnt) {
	return cnt <= 29u ? 0u : cnt <= 61u ? 1u : cnt <= 125u ? 2u : cnt < SS_CLS4_MIN ? 3u : cnt < SS_CLS5_MIN ? 4u : 5u;
}
enum { ERR_COUNT_LIMIT = 1, ERR_POOL_OVERFLOW = 2, ERR_INJ_CAND = 4 };
enum { STAT_SPLIT_BYTES = 0, STAT_NODES = 1, STAT_COUNT = 4 };

// everything the level kernels need, passed by value
struct TrainArgs {
	uint32_t nTrees;      // trees in this batch = partitions * T
	uint32_t T;           // trees per partition
	uint32_t m;
	uint32_t mtry;
	uint32_t stride;      // list / flag / inbag stride per tree = max Ntr of the batch
	uint32_t NC;          // node capacity per tree
	uint32_t itemCap;
	uint32_t injCandNodes;
	int listBuf;          // which of lists[0/1] holds the current level's segments
	const PartDesc* parts;
	const double* Dall;
	const entry_t* Sall;
	void* lists[2];       // [tree][feature][stride] of entry_t (entry32 == 0) or uint32_t (entry32 == 1)
	int entry32;
	uint32_t* inbag;      // [tree][stride]
	unsigned long long* mt;   // [tree][312]
	TreeState* ts;
	Node16* nodes;        // [tree][NC]
	uint32_t* flagBits;   // [tree][flagWords] go-left flag per in-bag row, bit-packed (tree_flags_kernel -> partition_lists_kernel)
	uint32_t flagWords;
	Item* items[2];
	uint32_t* clsItems[2][NCLS];   // [buf][class][itemCap] item ids of each size class (same ping-pong as items)
	int useClasses;       // 0: every item goes to class 3 (one generic launch per level)
	uint32_t* cand;       // [itemCap][mtry]
	SearchRes* res;       // [itemCap][mtry]
	SplitDec* dec;        // [itemCap]
	PartTask* ptask;      // [itemCap]
	uint32_t* splitFeat;  // [itemCap] winning feature of each split node
	uint32_t* counters;   // [CNT_COUNT]
	unsigned long long* stats;   // [STAT_COUNT]
	const uint32_t* injCand;     // [tree][injCandNodes][mtry] or nullptr
};

// knn.cu
void launch_gather_rows(const double* x, uint32_t ld, uint32_t m, const uint32_t* idx, uint32_t count, double* out, cudaStream_t s);
int launch_knn(const double* Xp, int P, int m, int localk, double* Dscratch, size_t scratchElems, int32_t* nnIdx, double* nnDist, cudaStream_t s);
// knn_tc.cu (tcgen05 3xTF32 candidate search + exact fp64 rerank)
int knn_tensor_candidates(int localk);
size_t knn_tensor_workspace_bytes(int P, int m, int C);
int launch_knn_tensor(const double* Xp, int P, int m, int localk, void* workspace, int32_t* nnIdx, double* nnDist, uint32_t* nIncompleteDev,
                      cudaStream_t s);
// assemble.cu
void launch_assemble(const double* x, uint32_t m, const uint32_t* posIdx, uint32_t P, const uint32_t* negIdx, const int32_t* nn,
                     uint32_t localk, uint32_t fp, const int32_t* draws, const PartDesc* parts, uint32_t nPartsBatch,
                     uint32_t maxNtr, double* Dall, int* err, cudaStream_t s);
// sort.cu
void launch_feature_sort(const double* Dall, const PartDesc* parts, uint32_t nPartsBatch, uint32_t m, unsigned long long* keysA,
                         unsigned long long* keysB, uint32_t* valsA, uint32_t* valsB, entry_t* Sall, cudaStream_t s);
int launch_feature_sort_smem(const double* Dall, const PartDesc* parts, uint32_t nPartsBatch, uint32_t m, uint32_t maxNtr, entry_t* Sall,
                             uint32_t* slowFlags /*[nPartsBatch * m] scratch, or NULL: bitonic network only*/,
                             unsigned short* rankTab /*[element offsets of Sall] dense rank by row, or NULL*/, cudaStream_t s);
// train.cu
size_t tree_smem_bytes(uint32_t m, uint32_t mtry);
void launch_bootstrap(const TrainArgs& A, const uint32_t* forestSeeds, cudaStream_t s);
void launch_inbag_from_ids(const TrainArgs& A, const unsigned long long* ids, const uint64_t* idOffsets, cudaStream_t s);
// inbagT: scratch of nParts * stride * 16 bytes for the all-trees-at-once variant (compact entries, T <= 16), or NULL
void launch_build_lists(const TrainArgs& A, unsigned char* inbagT, cudaStream_t s);
int launch_init_trees(const TrainArgs& A, cudaStream_t s);
int launch_split_search(const TrainArgs& A, const uint32_t* clsCount, int buf, cudaStream_t s);
void launch_split_select(const TrainArgs& A, uint32_t nItems, int buf, cudaStream_t s);
int launch_tree_level(const TrainArgs& A, int buf, cudaStream_t s);   // go-left bitmaps + level bookkeeping, one CTA per tree
void launch_partition_lists(const TrainArgs& A, uint32_t nItems, int buf, cudaStream_t s);
// glibc_rand.cu
void glibc_window_jump(uint32_t* w31, uint64_t n);
int launch_glibc_rand_fill(const uint32_t* w31, unsigned long long count, uint32_t* scratchDev, int32_t* out, cudaStream_t s);
// predict.cu
void launch_pack_forest(const Node16* nodes, const TreeState* ts, uint32_t NC, uint32_t nTrees, void* nodes8, cudaStream_t s);
// > 0: the fold scores class12 were updated directly (skip launch_accumulate); 0: the partition means are in scratch; < 0: error
int launch_predict(const double* x, uint32_t m, const uint32_t* testIdx, uint32_t Nte, const Node16* nodes, const void* nodes8, uint32_t NC,
                   const TreeState* ts, uint32_t maxNodes, uint32_t nPartsBatch, uint32_t T, double* scratch, double divider, double* class12,
                   cudaStream_t s);
void launch_accumulate(const double* scratch, uint32_t Nte, uint32_t nPartsBatch, double divider, double* class12, cudaStream_t s);
// curves.cu
int launch_curves(const unsigned char* labelsSorted, uint64_t N, uint32_t totP, uint32_t totN, uint32_t* tpScratch,
                  double* partials, size_t partialsCap, double* out2, cudaStream_t s);
void launch_tie_reverse_labels(const unsigned long long* keysSorted, const unsigned char* labelsSorted, uint64_t N, unsigned char* out, cudaStream_t s);
void launch_tie_mixed_count(const unsigned long long* keysSorted, const unsigned char* labelsSorted, uint64_t N, unsigned long long* counter, cudaStream_t s);
void launch_gather_u32(const uint32_t* src, const uint32_t* idx, uint64_t N, uint32_t* out, cudaStream_t s);
void launch_check_ids(const uint32_t* ids, uint64_t N, uint32_t n, uint32_t* bad /*device word, OR-ed*/, cudaStream_t s);
void launch_gather_f64(const double* src, const uint32_t* idx, uint64_t N, double* out, cudaStream_t s);
void launch_add_f64(double* dst, const double* src, size_t count, cudaStream_t s);
void launch_scatter_scores(const double* class12, const uint32_t* testIdx, uint32_t Nte, uint32_t n, double* glob, cudaStream_t s);
void launch_gather_labels(const uint32_t* labels, const uint32_t* perm, uint64_t N, unsigned char* out, cudaStream_t s);
int launch_sort_scores_desc(const double* scores, uint64_t N, unsigned long long* keysA, unsigned long long* keysB, uint32_t* valsA,
                            uint32_t* valsB, uint32_t* hist, cudaStream_t s);

int launch_sort_scores_desc_labelties(const double* scores, const uint32_t* labels, uint64_t N, int posFirst, unsigned long long* keysA,
                                      unsigned long long* keysB, uint32_t* valsA, uint32_t* valsB, uint32_t* hist, cudaStream_t s);
// fold keys: keys[i] = (labels[i] > 0 ? 0 : nFolds) + foldIds[i], vals[i] = i, counts[key]++ (counts zeroed by the caller); *bad |= 1 if a fold id >= nFolds
void launch_fold_keys(const uint32_t* labels, const uint32_t* foldIds, uint32_t n, uint32_t nFolds, unsigned long long* keys, uint32_t* vals, uint32_t* counts,
                      uint32_t* bad, cudaStream_t s);
int launch_radix_sort_u64(unsigned long long* keysA, unsigned long long* keysB, uint32_t* valsA, uint32_t* valsB, uint64_t N, int passes,
                          uint32_t* hist, cudaStream_t s, int bit0 = 0);
// shuffle.cu
int shuffle_sort_passes(uint32_t n);   // radix passes of launch_device_shuffle (kernel launches: 2 + 3 * passes)
int launch_device_shuffle(const int32_t* draws, const uint32_t* in, uint32_t n, unsigned long long* keysA, unsigned long long* keysB, uint32_t* valsA,
                          uint32_t* valsB, uint32_t* hist, uint32_t* jArr, uint32_t* out, cudaStream_t s);

}  // namespace psm
