/* TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the hyperSMURF partition pipeline of
 * AnacletoLAB/parSMURF.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path (parsmurf_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares every function below with the
 * unmodified reference classes compiled into oracle/_ref/libparsmurf_ref.so (oracle/Makefile), and
 * tests/test_oracle_golden.py compares it with the committed fixtures in tests/golden/ that
 * scripts/make_golden.py generated from that same reference build.
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef PSM_ORACLE_H
#define PSM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* glibc rand()/srand() (TYPE_3 additive feedback generator, stdlib/random_r.c) -- the stream the
 * reference draws from at src/samplerKNN.cpp:95,99, src/folds.cpp:41-42, src/testtraindivider.cpp:90. */
void*   psmo_rand_create(uint32_t seed);
void    psmo_rand_destroy(void* st);
int32_t psmo_rand_next(void* st);
void    psmo_rand_fill(void* st, int32_t* out, uint64_t count);

/* Folds from a fold-assignment vector (src/folds.cpp:79-117) + TestTrainDivider (src/testtraindivider.cpp:5-106),
 * including std::random_shuffle of the training negatives driven by `rand_state`.
 * Call once with all index pointers NULL to get sizes[4] = {trngPos, trngNeg, testPos, testNeg} for `fold`. */
void* psmo_ttd_create(const uint32_t* labels, const uint32_t* foldAssign, uint32_t n, uint32_t nFolds, void* rand_state);
void  psmo_ttd_destroy(void* ttd);
void  psmo_ttd_fold_sizes(void* ttd, uint32_t fold, uint32_t* sizes);
void  psmo_ttd_fold_indices(void* ttd, uint32_t fold, uint32_t* trngPos, uint32_t* trngNeg, uint32_t* testPos, uint32_t* testNeg);

/* Sampler::setPartition + SamplerKNN::setup (src/sampler.cpp:62-71, src/samplerKNN.cpp:15-30).
 * out[6] = {chunk, negAvail, negOffset, nPosOut, nNegOut, Ntr}.  Returns 0, or <0 where the reference
 * would abort()/hit undefined behaviour (P<2, uint32 underflow of the last chunk). */
int psmo_partition_sizes(uint32_t P, uint32_t Nneg, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t p, uint32_t* out);

/* Exact kNN among `P` points [P][dim] row-major: squared L2 accumulated in ascending dimension,
 * zero distances excluded, ascending by (distance, index); missing slots -> index -1, distance +inf
 * (src/ann_1.1.2/src/kd_search.cpp:172-210, src/ann_1.1.2/src/brute.cpp:54-78, ANN.h:235). */
void psmo_knn(const double* pts, uint32_t P, uint32_t dim, uint32_t k, int32_t* nnIdx, double* nnDist);

/* SamplerKNN::minOversample + majUndersample (src/samplerKNN.cpp:32-141): row-major training matrix
 * [Ntr][m+1].  `nn` = [P][localk] neighbour table, `draws` = the rand() outputs in consumption order
 * (1 + m per synthetic row).  Returns 0, -2 if a needed neighbour slot is -1, -3 if localk < 2. */
int psmo_assemble(const double* x, uint32_t m, const uint32_t* posIdx, uint32_t P, const uint32_t* negIdx,
                  uint32_t nNegOut, const int32_t* nn, uint32_t localk, uint32_t fp, const int32_t* draws, double* out);

/* ranger ForestProbability training as rfRanger configures it (src/rfRanger.cpp:6-54;
 * src/ranger/Forest/Forest.cpp:423-525; Tree.cpp:77-135,245-346,413-446; TreeProbability.cpp:47-63,81-358;
 * utility.cpp:84-139): column-major data [m+1][Ntr], response = column m, min_node_size 10,
 * sample_fraction 1 with replacement, tree seed (uint32)((t+1)*seed). */
void*    psmo_forest_train(const double* colmajor, uint32_t Ntr, uint32_t m, uint32_t T, uint32_t mtry, uint32_t seed);
void     psmo_forest_destroy(void* forest);
uint64_t psmo_forest_tree_nodes(void* forest, uint32_t t);
/* ranger getter layouts: left/right child (0 = none), split var, split value, frac[2*node+c] (NaN for internal). */
void     psmo_forest_export_tree(void* forest, uint32_t t, uint64_t* left, uint64_t* right, uint64_t* var, double* value, double* frac);
/* algorithmic split-search bytes of SURVEY §8(d): sum over nodes that reach findBestSplit of n_node*(mtry*8+8) */
uint64_t psmo_forest_split_bytes(void* forest);
/* RNG replay exports (fact #6): bootstrap sampleIDs [Ntr] and candidate lists [nNodes][mtry] of tree t */
void     psmo_tree_draws(uint32_t tree_seed, uint32_t Ntr, uint32_t m, uint32_t mtry, uint64_t nNodes, uint64_t* bootstrap, uint32_t* cand);
/* Tree::predict + ForestProbability::predictInternal (Tree.cpp:137-193, ForestProbability.cpp:124-150).
 * test rows row-major [Nte][ld] (ld >= m; only features 0..m-1 are read); preds [Nte][2]. */
void     psmo_forest_predict(void* forest, const double* rows, uint32_t Nte, uint32_t ld, double* preds);

/* Curves ctor + evalAUROC_ok + evalAUPRC (src/curves.cpp:7-36,77-122; src/curves.h:119-128). out = {auroc, auprc}.
 * If perm_out != NULL it receives the descending std::sort permutation (size_t -> uint64). */
void psmo_curves(const uint32_t* labels, const double* preds, uint64_t N, double* out, uint64_t* perm_out);

/* The whole CV as `parSMURF1` runs it with ensThrd = 1 (src/hyperSMURF.cpp:151-496): glibc rand seeded 1,
 * folds from `foldAssign`, per fold/partition sample->train->predict->accumulate in ascending order.
 * class1/class2 length n, zero-filled by the caller.  fold range [fold0, fold1). nthreads > 1 runs partitions
 * of a fold concurrently (SMOTE draws are still taken in partition order, so results do not change). */
int psmo_cv(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
            uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
            uint32_t fold0, uint32_t fold1, uint32_t nthreads, double* class1, double* class2);

/* One rank's share of a partition-sharded CV: only partitions [part0, part1) of every fold are trained/accumulated. */
int psmo_cv_parts(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
                  uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
                  uint32_t part0, uint32_t part1, double* class1, double* class2);
int psmo_cv_ranges(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
                   uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
                   const uint32_t* first, const uint32_t* last, double* class1, double* class2);

#ifdef __cplusplus
}
#endif
#endif
