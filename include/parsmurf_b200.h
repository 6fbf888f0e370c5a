/* parsmurf_b200 -- C ABI of the B200-native hyperSMURF partition pipeline.
 *
 * The reference (AnacletoLAB/parSMURF) has no FFI layer: the boundary this library replaces is the set of
 * C++ class interfaces that hyperSMURF::smurfIt calls per fold and per partition
 * (src/hyperSMURF.cpp:255-396).  Each entry group below cites the reference interface it stands in for.
 * The C++ host mirror of those classes (parsmurf_b200/host/) and the Python binding (parsmurf_b200/_capi.py)
 * are both thin callers of exactly these symbols.
 *
 * Conventions: every call returns PSM_OK (0) or a negative PSM_E* code and never throws or aborts across
 * the ABI; psm_last_error() gives the text.  Host buffers are borrowed for the duration of the call.
 * All calls on one context are issued from one host thread; partition concurrency is expressed as batched
 * launches, not host threads.  There is NO CPU fallback: without a CUDA device every compute entry fails
 * with PSM_E_CUDA.
 */
#ifndef PARSMURF_B200_H
#define PARSMURF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSM_OK 0
#define PSM_E_CUDA (-1)        /* CUDA runtime error (no device, launch failure, out of memory) */
#define PSM_E_ARG (-2)         /* invalid argument / call order */
#define PSM_E_TOOFEW_POS (-3)  /* P < 2: the reference abort()s (src/samplerKNN.cpp:44-48) */
#define PSM_E_LOCALK (-4)      /* min(k,P) < 2 with fp > 0: the reference divides by zero (src/samplerKNN.cpp:95) */
#define PSM_E_NEIGHBOUR (-5)   /* a neighbour slot SMOTE needs is -1 (duplicates): the reference reads pos[-1] */
#define PSM_E_CHUNK (-6)       /* last negative chunk underflows uint32 in the reference (src/sampler.cpp:68) */
#define PSM_E_LIMIT (-7)       /* size outside this build's limits (Ntr >= 2^24, bootstrap count >= 2^15, k > 64) */
#define PSM_E_OVERFLOW (-8)    /* internal capacity exceeded (node / item pools) */

typedef struct psm_ctx psm_ctx;

/* ---- (1) context --------------------------------------------------------------------------------------- */
int psm_ctx_create(int device, psm_ctx** out);
int psm_ctx_destroy(psm_ctx* ctx);
const char* psm_last_error(psm_ctx* ctx);
/* Launch everything on `cuda_stream` (a cudaStream_t; NULL = the legacy default stream). */
int psm_set_stream(psm_ctx* ctx, void* cuda_stream);
/* Device-memory budget (bytes) for one partition batch of the training kernels; 0 = default (24 GiB). */
int psm_set_batch_budget(psm_ctx* ctx, uint64_t bytes);
int psm_sync(psm_ctx* ctx);
/* Fold lanes: folds of one cross-validation are independent until the final score merge (src/hyperSMURF.cpp:161-396 runs
 * them one after the other), so a host may drive several folds at once, one context per lane on the same device, each from
 * its own host thread.  psm_ctx_own_stream gives the context a private non-blocking stream (destroyed with the context);
 * psm_dataset_devptr returns the device copy of x so that the other lanes can psm_dataset_attach_device it instead of
 * uploading it again; psm_scores_merge adds src's device class1Prob/class2Prob[n] into dst's (test sets of different folds
 * are disjoint, so every element receives at most one non-zero term and the sum is exact) after src's stream has drained. */
int psm_ctx_own_stream(psm_ctx* ctx);
int psm_ctx_bind_thread(psm_ctx* ctx);   /* make the context's device current in the calling host thread */
int psm_dataset_devptr(psm_ctx* ctx, const double** x_dev);
int psm_scores_merge(psm_ctx* dst, psm_ctx* src);

/* ---- (2) dataset: `x` of hyperSMURF's ctor (src/hyperSMURF.h:26-28): row-major f64 [n][m+1], column m is
 *      1.0 for positives / 2.0 for negatives (src/fileImport.cpp:88-91).  Uploaded once. ------------------- */
int psm_dataset_upload(psm_ctx* ctx, const double* x_host, uint32_t n, uint32_t m);
int psm_dataset_attach_device(psm_ctx* ctx, const double* x_dev, uint32_t n, uint32_t m);

/* ---- (3) fold: the four index arrays Partition::setFold exposes (src/partition.cpp:25-37) ---------------- */
int psm_fold_begin(psm_ctx* ctx, const uint32_t* trngPosIdx, uint32_t trngPosNum, const uint32_t* trngNegIdx,
                   uint32_t trngNegNum, const uint32_t* testPosIdx, uint32_t testPosNum,
                   const uint32_t* testNegIdx, uint32_t testNegNum);

/* ---- (3b) folds on the device: replaces TestTrainDivider's per-fold index arrays and their std::random_shuffle
 *      (src/testtraindivider.cpp:5-106).  psm_folds_upload takes, once per CV, the examples of every fold split by class in
 *      the order Folds::computeFolds leaves them (src/folds.cpp:26-117): posList / negList with nFolds+1 offsets each.
 *      psm_fold_begin_device(fold, window) then opens `fold` like psm_fold_begin, but builds the four index arrays on the
 *      device: test = the fold's positives then its negatives; training = the other folds' blocks in ascending fold order
 *      (:57-80); and, when shuffle_window31 != NULL, the training negatives shuffled exactly like
 *      std::random_shuffle(trngNeg.begin(), trngNeg.end()) driven by glibc rand() whose last 31 state words (oldest first)
 *      are `shuffle_window31` (:90) -- the draws are replayed on the device and the swap chain is resolved per output slot
 *      (csrc/shuffle.cu), so nothing of size n is built, shuffled or uploaded by the host.
 *      psm_folds_attach shares src's device lists (fold lanes); psm_fold_indices_get copies the open fold's arrays back
 *      (sizes4 = {trngPos, trngNeg, testPos, testNeg}; pointers may be NULL); psm_scores_fold_curves_device is
 *      psm_scores_fold_curves for a fold of the uploaded lists. */
int psm_folds_upload(psm_ctx* ctx, uint32_t nFolds, const uint32_t* posList, const uint32_t* posOff, const uint32_t* negList,
                     const uint32_t* negOff);
int psm_folds_attach(psm_ctx* ctx, psm_ctx* src);
/* The per-class fold lists of Folds::computeFolds' "folds from a file" branch (src/folds.cpp:79-117) built on the device from
 * the per-example fold ids: one stable two-pass radix sort of the example ids by (class, fold), so every fold keeps ascending
 * example order, positives and negatives apart -- what psm_folds_upload would have been given.  Needs psm_labels_upload first
 * (labels > 0 = positive).  fold_ids (host, [n]) stay resident and are uploaded again only when they change; the lists are
 * rebuilt on every call.  nPos / nNeg [nFolds] receive the fold sizes; psm_folds_download copies the lists to the host
 * (posList [sum nPos], negList [sum nNeg]) for callers that still want them (inner CVs, train mode). */
int psm_folds_build_device(psm_ctx* ctx, const uint32_t* fold_ids, uint32_t nFolds, uint32_t* nPos, uint32_t* nNeg);
int psm_folds_download(psm_ctx* ctx, uint32_t* posList, uint32_t* negList);
int psm_fold_begin_device(psm_ctx* ctx, uint32_t fold, const uint32_t* shuffle_window31);
int psm_fold_indices_get(psm_ctx* ctx, uint32_t* sizes4, uint32_t* trngPos, uint32_t* trngNeg, uint32_t* testPos, uint32_t* testNeg);
int psm_scores_fold_curves_device(psm_ctx* ctx, uint32_t fold, int tie_mode, double* out);

/* ---- (4) kNN: replaces ANNkd_tree + annkSearch in SamplerKNN::minOversample (src/samplerKNN.cpp:60-93).
 *      Exact brute force over the fold's positives; zero distances excluded; ascending (distance, index);
 *      missing slots = -1 / +inf.  localk = min(k, P).  Computed once per fold (fact: pos is the same for
 *      every partition, src/sampler.cpp:69). ---------------------------------------------------------------- */
int psm_fold_knn(psm_ctx* ctx, uint32_t k);
/* How the last psm_fold_knn ran: *path = 0 exact fp64 kernel; 1 tensor-core candidate search (tcgen05 3xTF32 distance GEMM)
 * + exact fp64 rerank, proven complete for every query; 2 the tensor path left *incomplete queries unproven and the fold was
 * redone with the exact kernel.  The neighbour table is the same in all three cases.  The tensor path is taken when
 * P*P*m >= 2^33 and m >= 64 and k <= 28, or forced / forbidden with the environment variable PSM_KNN_TENSOR=1 / 0. */
int psm_fold_knn_path(psm_ctx* ctx, int* path, uint32_t* incomplete);
int psm_fold_knn_get(psm_ctx* ctx, int32_t* nnIdx /*[P][localk]*/, double* nnDist /*[P][localk] or NULL*/);
int psm_fold_knn_set(psm_ctx* ctx, const int32_t* nnIdx, uint32_t localk);   /* inject a neighbour table */

/* ---- (5) partition batch: Sampler::setPartition + SamplerKNN::sample (src/sampler.cpp:62-71,
 *      src/samplerKNN.cpp:8-141) + transposeMatrix (src/hyperSMURF.cpp:289-291) for partitions [p0,p1).
 *      `draws` = the rand() outputs SMOTE consumes, in consumption order, partitions concatenated
 *      (P*fp*(m+1) per partition; src/samplerKNN.cpp:95,99). ------------------------------------------------ */
int psm_batch_begin(psm_ctx* ctx, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t p0, uint32_t p1);
int psm_batch_assemble(psm_ctx* ctx, const int32_t* draws_host, uint64_t ndraws);
/* Same, but the draws are generated ON THE DEVICE by replaying glibc's rand() (TYPE_3 additive-feedback generator)
 * from `window31`: the generator's last 31 state words in chronological order (oldest first), positioned at the first
 * draw of partition p0.  psm_glibc_jump advances such a window by n draws on the host in O(log n). */
int psm_batch_assemble_glibc(psm_ctx* ctx, const uint32_t* window31);
int psm_glibc_jump(uint32_t* window31, uint64_t n);
/* The device replay by itself: generate the next `count` rand() outputs of the stream at `window31` on the GPU (count below
 * 248 * 2^24) and copy outputs [skip, skip + n_out) to the host -- what srand(seed); rand() ... would return on the reference's
 * platform (glibc stdlib/random_r.c TYPE_3).  Diagnostic entry: the CV path keeps the draws on the device. */
int psm_glibc_fill_device(psm_ctx* ctx, const uint32_t* window31, uint64_t count, uint64_t skip, int32_t* out_host, uint64_t n_out);
/* sizes[3] = {trngPos, trngNeg, Ntr}; out = row-major [Ntr][m+1] like Sampler::trngData (may be NULL) */
int psm_batch_get_training(psm_ctx* ctx, uint32_t p, uint32_t* sizes, double* out_rowmajor);

/* ---- (6) forest training: rfRanger train ctor + train() (src/rfRanger.cpp:6-54,177-183) -> ranger
 *      ForestProbability (min_node_size 10, sample_fraction 1 with replacement, no importance).
 *      forest_seeds[p-p0] = seed + p + fold*nParts (src/hyperSMURF.cpp:273).  Optional injection of the
 *      reference's draws: bootstrap[(p-p0)*T+t][Ntr] sample ids (stride = that partition's Ntr, partitions
 *      concatenated) and cand[(p-p0)*T+t][cand_nodes][mtry] split-variable candidates per node id. ----------- */
int psm_batch_train(psm_ctx* ctx, uint32_t nTrees, uint32_t mtry, const uint32_t* forest_seeds,
                    const uint64_t* inj_bootstrap, const uint32_t* inj_cand, uint32_t inj_cand_nodes);
int psm_batch_tree_nodes(psm_ctx* ctx, uint32_t p, uint32_t t, uint64_t* nNodes);
/* ranger getter layouts (src/ranger/Forest/Forest.h:82-101, ForestProbability.h:40-44): child ids (0 = none),
 * split var, split value, frac[2*node+c] (NaN for internal nodes; class 0 = positive). */
int psm_batch_export_tree(psm_ctx* ctx, uint32_t p, uint32_t t, uint64_t* left, uint64_t* right, uint64_t* var,
                          double* value, double* frac);
/* ---- (6b) forests that were grown earlier (predict mode, src/hyperSMURF.cpp:401-446): the rfRanger(forestFilename, ...)
 *      constructor (src/rfRanger.cpp:120-168 -> ForestProbability::loadFromFileInternal, ForestProbability.cpp:275-328).
 *      Installs the forests of partitions [p0, p1), nTrees each, as the current (trained) batch so that
 *      psm_batch_predict_accumulate can run on them; no training state is needed, only an open fold.  Trees are given in
 *      the layouts psm_batch_export_tree produces, concatenated tree after tree: nNodes[(p-p0)*nTrees + t] nodes each,
 *      frac holding 2 doubles per node.  Children must be consecutive (right == left + 1), as ranger creates them. --- */
int psm_batch_import_forests(psm_ctx* ctx, uint32_t p0, uint32_t p1, uint32_t nTrees, const uint64_t* nNodes,
                             const uint64_t* left, const uint64_t* right, const uint64_t* var, const double* value,
                             const double* frac);
/* bootstrap multiplicities of tree (p,t): counts[Ntr] (parity check of the device mt19937_64 replay) */
int psm_batch_get_inbag(psm_ctx* ctx, uint32_t p, uint32_t t, uint32_t* counts);

/* ---- (7) predict + accumulate: Sampler::copyTestSet + rfRanger predict ctor/predict() +
 *      accumulateAndDivideResInProbVect (src/sampler.cpp:81-98,116-133; src/rfRanger.cpp:56-118,185-191).
 *      Adds pred[i][c] * (1/nPartsTotal) of every partition of the batch, in ascending partition order,
 *      into the fold score buffer (rows = test positives then test negatives). ----------------------------- */
int psm_batch_predict_accumulate(psm_ctx* ctx, uint32_t nPartsTotal);

/* ---- (8) fold scores / reduce / scatter ------------------------------------------------------------------- */
int psm_fold_scores_get(psm_ctx* ctx, double* class1 /*[Nte]*/, double* class2 /*[Nte] or NULL*/);
/* device pointers of the fold score buffer [2][Nte] so the caller can ncclReduce / all_reduce it in place
 * (the MPI_Gather + master sum of src/hyperSMURF_MPI.cpp:594-618) */
int psm_fold_scores_devptr(psm_ctx* ctx, double** class12_dev, uint32_t* Nte);
/* class1Prob/class2Prob[testIdx[i]] += fold scores (scatter by original example id); host arrays length n */
int psm_fold_scatter_host(psm_ctx* ctx, double* class1Prob, double* class2Prob);

/* Device-resident variants (no host round trip of the scores): the 0/1 labels `y` of hyperSMURF's ctor are uploaded once;
 * fold scores are scattered into device class1Prob/class2Prob[n] (the arrays main() owns in src/parSMURF.cpp:90-93);
 * per-fold and global AUROC/AUPRC are computed from device memory; psm_scores_get ADDS the device arrays into the host ones. */
int psm_labels_upload(psm_ctx* ctx, const uint32_t* labels, uint32_t n);
int psm_scores_reset(psm_ctx* ctx);
int psm_fold_scatter_device(psm_ctx* ctx);
int psm_fold_curves(psm_ctx* ctx, int tie_mode, double* out);
int psm_scores_get(psm_ctx* ctx, double* class1Prob, double* class2Prob);
/* the same, but the host arrays are overwritten instead of added to (no zero-fill and no read-modify-write of 2n doubles) */
int psm_scores_read(psm_ctx* ctx, double* class1Prob, double* class2Prob);
int psm_scores_curves(psm_ctx* ctx, int tie_mode, double* out);
/* Unit-sharded multi-GPU runs (every rank owns a contiguous range of the nFolds x nParts (fold, partition) units and only
 * opens the folds it has units of): the per-rank partial class1Prob/class2Prob[n] arrays are summed ONCE, in place, through
 * psm_scores_devptr (device pointer of [2][n] f64), and the per-fold AUROC/AUPRC (src/hyperSMURF.cpp:365-396) are then
 * computed from the summed arrays for the fold whose test rows are given (positives first, then negatives).
 * psm_scratch_f64 returns a small zeroed device buffer that psm_scratch_write / psm_scratch_read fill and read back, so
 * that a few host doubles (per-fold metrics) can go through the same device all-reduce as the scores. */
int psm_scores_devptr(psm_ctx* ctx, double** class12Prob_dev, uint32_t* n);
int psm_scores_fold_curves(psm_ctx* ctx, const uint32_t* testPosIdx, uint32_t testPosNum, const uint32_t* testNegIdx,
                           uint32_t testNegNum, int tie_mode, double* out);
int psm_scratch_f64(psm_ctx* ctx, uint32_t count, double** dev);
int psm_scratch_write(psm_ctx* ctx, const double* host, uint32_t count);
int psm_scratch_read(psm_ctx* ctx, double* host, uint32_t count);

/* ---- (8b) cross-rank sum inside the library: the MPI_Gather + master sum of the score vectors
 *      (src/hyperSMURF_MPI.cpp:594-646; communicator set-up src/MPIHelper.cpp:29-37) as ONE NCCL all-reduce over
 *      NVLink / NVSwitch, issued on the context's stream.  One rank calls psm_comm_unique_id and hands the 128 bytes to the
 *      others by whatever bootstrap the host has (MPI_Bcast, torch.distributed, a file, shared memory of a multi-threaded
 *      process); every rank then calls psm_comm_init (collective: returns when all `world` ranks have joined).  NCCL is
 *      loaded at the first call (dlopen of libnccl.so.2 -- the copy the process already holds, if any), so a single-GPU
 *      host never needs it.  psm_comm_allreduce_sum_f64 sums `count` doubles at device pointer `dev` across the ranks, in
 *      place, and returns when the result is visible to the stream (it synchronises the stream). ------------------------- */
#define PSM_COMM_ID_BYTES 128
int psm_comm_unique_id(void* id128);
int psm_comm_init(psm_ctx* ctx, const void* id128, uint32_t rank, uint32_t world);
int psm_comm_info(psm_ctx* ctx, uint32_t* rank, uint32_t* world, int* nccl_version);
int psm_comm_allreduce_sum_f64(psm_ctx* ctx, double* dev, uint64_t count);
/* psm_dataset_upload for a replicated dataset on `world` ranks: every rank copies only its 1/world slice of the rows over
 * PCIe and one in-place ncclAllGather over NVLink completes all the copies.  Collective (same n, m on every rank); a rank reads
 * only rows [rank*ceil(n/world), ...) of x_host. */
int psm_dataset_upload_sharded(psm_ctx* ctx, const double* x_host, uint32_t n, uint32_t m);
int psm_comm_destroy(psm_ctx* ctx);

/* ---- (9) curves: Curves ctor + evalAUROC_ok + evalAUPRC (src/curves.cpp:7-36,77-122). out = {auroc, auprc}.
 *      tie_mode 0: descending order from the identical host std::sort call (bit-parity with the reference's
 *      tie order); tie_mode 1: stable device radix sort (ties in index order; may differ from the reference by
 *      more than 1e-6 on heavily tied scores); tie_mode 2: see psm_curves_info. -------------- */
int psm_curves(psm_ctx* ctx, const uint32_t* labels, const double* scores, uint64_t N, int tie_mode, double* out);
/* tie_mode 2 ("auto"): the stable device sort with the ties of equal score ordered positives first (the order that maximises
 * both curves), and a count of the adjacent pairs of EQUAL score and DIFFERENT label.  With none, every run of tied scores is
 * single-label, its internal order cannot change a TP/FP prefix, and the result is exactly the reference's whatever order its
 * std::sort leaves ties in.  Otherwise a second sort orders the ties negatives first (the minimum): the reference's value lies
 * between the two, and when they are within 1e-6 of each other their midpoint is returned (error <= 5e-7); only when the tie order matters
 * more than that is the permutation redone with the host std::sort of tie_mode 0.  psm_curves_info reports, for the last
 * curves call of the context, that pair count, whether the host sort ran, and tie_spread2 = {max - min AUROC, max - min AUPRC}. */
int psm_curves_info(psm_ctx* ctx, uint64_t* mixed_tie_pairs, int* used_host_sort, double* tie_spread2);
/* Host only (no device needed): the permutation std::sort leaves an index vector in under the reference's comparator
 * (src/curves.cpp:31-33: descending score, ties wherever libstdc++'s introsort puts them), computed on `threads` host threads
 * (0 = up to 16) with libstdc++'s own partition / insertion / heap steps.  This is what tie_mode 0 (and the fallback of
 * tie_mode 2) uploads. */
int psm_host_sort_permutation(const double* scores, uint64_t N, uint32_t* perm, int threads);

/* ---- convenience: steps (5)-(7) for partitions [p0,p1) of the current fold, batched under the memory budget.
 *      draws_host as in psm_batch_assemble for [p0,p1) (or NULL when fp == 0). ------------------------------ */
int psm_fold_run_partitions(psm_ctx* ctx, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t nTrees, uint32_t mtry,
                            uint32_t seed, uint32_t fold, uint32_t p0, uint32_t p1, const int32_t* draws_host);

/* largest number of partitions starting at p0 that one batch may hold under the memory budget (>= 1) */
int psm_batch_capacity(psm_ctx* ctx, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t nTrees, uint32_t mtry, uint32_t p0,
                       uint32_t* capacity);

/* ---- instrumentation: accumulated device time (ms, CUDA events on the context stream) and launch counts per
 *      kernel class since the last reset.  ids: 0 knn, 1 assemble, 2 feature_sort, 3 bootstrap, 4 build_lists,
 *      5 split_search, 6 split_select, 7 level_finalize, 8 partition_lists, 9 predict, 10 curves, 11 gather. --- */
#define PSM_NUM_KERNEL_CLASSES 12
int psm_profile_enable(psm_ctx* ctx, int on);
/* restrict the event timing to the classes whose bit (1u << id) is set (default: all); launch counts are always kept.
 * Timing one class costs two event records per launch of it, so a throughput run times only the kernel it reports on. */
int psm_profile_classes(psm_ctx* ctx, uint32_t class_mask);
int psm_profile_reset(psm_ctx* ctx);
int psm_profile_get(psm_ctx* ctx, double* ms /*[12]*/, uint64_t* launches /*[12]*/);
/* algorithmic split-search bytes (sum over searched nodes of n_node*(mtry*8+8)) since last reset */
int psm_stats_get(psm_ctx* ctx, uint64_t* split_bytes, uint64_t* nodes_total, uint64_t* levels_total);

const char* psm_version(void);

#ifdef __cplusplus
}
#endif
#endif
