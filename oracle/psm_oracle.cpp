// TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the hyperSMURF partition pipeline.
// See psm_oracle.h for scope, parity status (PINNED against oracle/_ref and tests/golden) and the
// rule that the product path never loads this library.  Written from the behaviour described in
// SURVEY.md §9; every function cites the reference file:line (relative to /root/reference) it follows.
// Build: oracle/Makefile (g++ -O2 -ffp-contract=off: no FMA contraction, like the reference's x86-64 build).

#include "psm_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <vector>
#include <omp.h>

namespace {

// ------------------------------------------------------------------ glibc rand()
// stdlib/random_r.c, TYPE_3: degree 31, separation 3; srandom_r seeds with the Park-Miller LCG
// (16807, Schrage's method) and discards 310 outputs; random_r: *f += *r; result = *f >> 1.
struct GlibcRand {
	int32_t tbl[31];
	int f, r;
	explicit GlibcRand(uint32_t seed) { reseed(seed); }
	void reseed(uint32_t seed) {
		if (seed == 0) seed = 1;
		tbl[0] = (int32_t)seed;
		int32_t word = (int32_t)seed;
		for (int i = 1; i < 31; i++) {
			long hi = word / 127773, lo = word % 127773;
			word = (int32_t)(16807 * lo - 2836 * hi);
			if (word < 0) word += 2147483647;
			tbl[i] = word;
		}
		f = 3; r = 0;
		for (int i = 0; i < 310; i++) next();
	}
	int32_t next() {
		uint32_t v = (uint32_t)tbl[f] + (uint32_t)tbl[r];
		tbl[f] = (int32_t)v;
		int32_t result = (int32_t)(v >> 1);
		if (++f >= 31) { f = 0; ++r; }
		else if (++r >= 31) r = 0;
		return result;
	}
};

// libstdc++ std::random_shuffle(first,last): for i = 1..n-1: j = rand() % (i+1); swap if i != j
// (bits/stl_algo.h) -- used at src/testtraindivider.cpp:90.
void random_shuffle_glibc(uint32_t* a, size_t n, GlibcRand& g) {
	for (size_t i = 1; i < n; i++) {
		size_t j = (size_t)(g.next() % (int64_t)(i + 1));
		if (i != j) std::swap(a[i], a[j]);
	}
}

// ------------------------------------------------------------------ folds / ttd
struct OTTD {
	uint32_t nFolds;
	std::vector<std::vector<uint32_t>> trngPos, trngNeg, testPos, testNeg;
};

OTTD* build_ttd(const uint32_t* labels, const uint32_t* ff, uint32_t n, uint32_t nFolds, GlibcRand* g) {
	// src/folds.cpp:79-110: folds from file = stable counting sort of example ids by fold id
	std::vector<uint32_t> foldsIdx(nFolds + 1, 0), hist(nFolds, 0), folds(n);
	for (uint32_t i = 0; i < n; i++) hist[ff[i]]++;
	for (uint32_t i = 1; i <= nFolds; i++) foldsIdx[i] = foldsIdx[i - 1] + hist[i - 1];
	std::fill(hist.begin(), hist.end(), 0);
	for (uint32_t i = 0; i < n; i++) { uint32_t b = ff[i]; folds[foldsIdx[b] + hist[b]++] = i; }
	// src/testtraindivider.cpp:20-91
	OTTD* t = new OTTD;
	t->nFolds = nFolds;
	t->trngPos.resize(nFolds); t->trngNeg.resize(nFolds); t->testPos.resize(nFolds); t->testNeg.resize(nFolds);
	for (uint32_t i = 0; i < nFolds; i++) {
		for (uint32_t j = foldsIdx[i]; j < foldsIdx[i + 1]; j++)
			(labels[folds[j]] < 1 ? t->testNeg[i] : t->testPos[i]).push_back(folds[j]);
		for (uint32_t kf = 0; kf < nFolds; kf++) {
			if (kf == i) continue;
			for (uint32_t j = foldsIdx[kf]; j < foldsIdx[kf + 1]; j++)
				(labels[folds[j]] < 1 ? t->trngNeg[i] : t->trngPos[i]).push_back(folds[j]);
		}
		if (g) random_shuffle_glibc(t->trngNeg[i].data(), t->trngNeg[i].size(), *g);
	}
	return t;
}

// ------------------------------------------------------------------ sizes
int partition_sizes(uint32_t P, uint32_t Nneg, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t p, uint32_t* out) {
	if (nParts == 0 || p >= nParts) return -4;
	uint32_t chunk = (uint32_t)std::ceil(Nneg / (double)nParts);              // src/sampler.cpp:65
	uint64_t before = (uint64_t)chunk * (nParts - 1);
	if (p == nParts - 1 && before > Nneg) return -5;                            // uint32 underflow in the reference
	uint32_t negAvail = (p != nParts - 1) ? chunk : (uint32_t)(Nneg - before); // src/sampler.cpp:68
	uint32_t nPosOut = (fp + 1) * P;                                            // src/samplerKNN.cpp:19
	uint32_t nNegOut = nPosOut * ratio;                                         // :20
	if (nNegOut == 0) nNegOut = negAvail;                                       // :22-23
	if (nNegOut > negAvail) nNegOut = negAvail;                                 // :28
	out[0] = chunk; out[1] = negAvail; out[2] = p * chunk; out[3] = nPosOut; out[4] = nNegOut; out[5] = nPosOut + nNegOut;
	if (P < 2) return -1;                                                       // :44-48 abort()
	return 0;
}

// ------------------------------------------------------------------ forest
struct OTree {
	std::vector<uint64_t> left, right, var;
	std::vector<double> val, frac0, frac1;
	std::vector<uint8_t> terminal;
};
struct OForest {
	uint32_t m, T, mtry;
	std::vector<OTree> trees;
	uint64_t split_bytes = 0;
};

// Tree::createPossibleSplitVarSubset -> drawWithoutReplacementSkip (Tree.cpp:245-258; utility.cpp:84-139).
// num_cols = m+1, the response column m is the single "no split" variable.
void draw_candidates(std::mt19937_64& rng, uint32_t m, uint32_t mtry, std::vector<size_t>& result) {
	size_t max = (size_t)m + 1;
	result.clear();
	if (mtry < max / 10) {
		// drawWithoutReplacementSimple (utility.cpp:94-117)
		std::vector<bool> temp(max, false);
		std::uniform_int_distribution<size_t> unif_dist(0, max - 1 - 1);
		for (size_t i = 0; i < mtry; ++i) {
			size_t draw;
			do {
				draw = unif_dist(rng);
				if (draw >= m) ++draw;      // skip value = m
			} while (temp[draw]);
			temp[draw] = true;
			result.push_back(draw);
		}
	} else {
		// drawWithoutReplacementFisherYates (utility.cpp:119-139)
		result.resize(max);
		std::iota(result.begin(), result.end(), 0);
		result.erase(result.begin() + m);
		std::uniform_real_distribution<double> distribution(0.0, 1.0);
		for (size_t i = 0; i < mtry; ++i) {
			size_t j = i + distribution(rng) * (max - 1 - i);
			std::swap(result[i], result[j]);
		}
		result.resize(mtry);
	}
}

void grow_tree(const double* D, uint32_t Ntr, uint32_t m, uint32_t mtry, uint32_t tree_seed,
               const std::vector<uint32_t>& classID, OTree& tr, uint64_t& split_bytes) {
	const size_t min_node_size = 10;                        // globals.h:86 via ForestProbability.cpp:63-65
	std::mt19937_64 rng;
	rng.seed(tree_seed);                                    // Tree.cpp:57
	// Tree::bootstrap (Tree.cpp:413-432)
	std::vector<std::vector<uint32_t>> nodeSamples(1);
	{
		std::uniform_int_distribution<size_t> unif_dist(0, Ntr - 1);
		nodeSamples[0].reserve(Ntr);
		for (size_t s = 0; s < Ntr; ++s) nodeSamples[0].push_back((uint32_t)unif_dist(rng));
	}
	auto new_node = [&]() {
		tr.left.push_back(0); tr.right.push_back(0); tr.var.push_back(0); tr.val.push_back(0);
		tr.frac0.push_back(std::nan("")); tr.frac1.push_back(std::nan("")); tr.terminal.push_back(0);
	};
	new_node();
	std::vector<size_t> cand;
	std::vector<std::pair<double, uint32_t>> vals;
	for (size_t i = 0; i < tr.var.size(); ++i) {            // node ids are BFS order (Tree.cpp:110-129, 294-302)
		draw_candidates(rng, m, mtry, cand);                // drawn before the terminal test (Tree.cpp:277-281)
		std::vector<uint32_t> S = std::move(nodeSamples[i]);
		const size_t n = S.size();
		size_t cnt[2] = {0, 0};
		for (uint32_t s : S) cnt[classID[s]]++;
		auto make_leaf = [&]() {                            // TreeProbability.cpp:47-63
			tr.terminal[i] = 1;
			double c0 = 0, c1 = 0;
			for (uint32_t s : S) { if (classID[s] == 0) ++c0; else ++c1; }
			tr.frac0[i] = c0 / (double)n; tr.frac1[i] = c1 / (double)n;
		};
		if (n <= min_node_size) { make_leaf(); continue; }  // TreeProbability.cpp:84-89
		if (cnt[0] == 0 || cnt[1] == 0) { make_leaf(); continue; }   // pure (:91-105)
		split_bytes += (uint64_t)n * ((uint64_t)mtry * 8 + 8);
		double best_decrease = -1, best_value = 0;
		size_t best_var = 0;
		for (size_t v : cand) {                             // TreeProbability.cpp:158-181 (SmallQ == LargeQ results)
			vals.clear();
			for (uint32_t s : S) vals.emplace_back(D[v * (size_t)Ntr + s], classID[s]);
			std::sort(vals.begin(), vals.end(), [](const std::pair<double, uint32_t>& a, const std::pair<double, uint32_t>& b) { return a.first < b.first; });
			size_t nl = 0, l[2] = {0, 0};
			size_t a = 0;
			while (a < n) {
				size_t b = a;
				while (b < n && vals[b].first == vals[a].first) { l[vals[b].second]++; ++b; }
				nl = b;
				if (b == n) break;                          // no split at the largest value
				size_t nr = n - nl;
				double sum_left = 0, sum_right = 0;
				for (int c = 0; c < 2; c++) {               // TreeProbability.cpp:322-334, weights 1.0
					size_t cr = cnt[c] - l[c];
					sum_left += 1.0 * l[c] * l[c];
					sum_right += 1.0 * cr * cr;
				}
				double decrease = sum_right / (double)nr + sum_left / (double)nl;    // :337
				if (decrease > best_decrease) {             // strict (:340)
					double lo = vals[a].first, hi = vals[b].first;
					best_value = (lo + hi) / 2;             // :348
					best_var = v;
					best_decrease = decrease;
					if (best_value == hi) best_value = lo;  // :353-355
				}
				a = b;
			}
		}
		if (best_decrease < 0) { make_leaf(); continue; }   // TreeProbability.cpp:184-186
		tr.var[i] = best_var; tr.val[i] = best_value;
		size_t lid = tr.var.size(); new_node();             // Tree.cpp:294-302
		size_t rid = tr.var.size(); new_node();
		tr.left[i] = lid; tr.right[i] = rid;
		nodeSamples.resize(tr.var.size());
		for (uint32_t s : S) {                              // Tree.cpp:305-318
			if (D[best_var * (size_t)Ntr + s] <= best_value) nodeSamples[lid].push_back(s);
			else nodeSamples[rid].push_back(s);
		}
	}
}

OForest* forest_train(const double* D, uint32_t Ntr, uint32_t m, uint32_t T, uint32_t mtry, uint32_t seed) {
	OForest* f = new OForest;
	f->m = m; f->T = T; f->mtry = mtry;
	// class ids by first appearance (ForestProbability.cpp:67-79)
	std::vector<double> class_values;
	std::vector<uint32_t> classID(Ntr);
	for (uint32_t i = 0; i < Ntr; i++) {
		double v = D[(size_t)m * Ntr + i];
		uint32_t id = (uint32_t)(std::find(class_values.begin(), class_values.end(), v) - class_values.begin());
		if (id == class_values.size()) class_values.push_back(v);
		classID[i] = id > 1 ? 1 : id;
	}
	f->trees.resize(T);
	for (uint32_t t = 0; t < T; t++) {
		uint32_t tree_seed = (uint32_t)((uint64_t)(t + 1) * seed);   // Forest.cpp:438
		grow_tree(D, Ntr, m, mtry, tree_seed, classID, f->trees[t], f->split_bytes);
	}
	return f;
}

void forest_predict(const OForest* f, const double* rows, uint32_t Nte, uint32_t ld, double* preds) {
	for (uint32_t i = 0; i < Nte; i++) {
		const double* r = rows + (size_t)i * ld;
		double s0 = 0, s1 = 0;
		for (const OTree& tr : f->trees) {
			size_t node = 0;
			while (!(tr.left[node] == 0 && tr.right[node] == 0))     // Tree.cpp:157-176
				node = (r[tr.var[node]] <= tr.val[node]) ? tr.left[node] : tr.right[node];
			s0 += tr.frac0[node]; s1 += tr.frac1[node];              // ForestProbability.cpp:136-141
		}
		preds[2 * i] = s0 / f->T; preds[2 * i + 1] = s1 / f->T;      // :145-149
	}
}

void knn(const double* pts, uint32_t P, uint32_t dim, uint32_t k, int32_t* nnIdx, double* nnDist) {
	std::vector<std::pair<double, int32_t>> best;
	for (uint32_t q = 0; q < P; q++) {
		best.clear();
		for (uint32_t j = 0; j < P; j++) {
			double d = 0;
			for (uint32_t c = 0; c < dim; c++) { double t = pts[(size_t)q * dim + c] - pts[(size_t)j * dim + c]; d += t * t; }
			if (d == 0) continue;                                    // ANN_ALLOW_SELF_MATCH = false (kd_search.cpp:200-201)
			best.emplace_back(d, (int32_t)j);
		}
		size_t kk = std::min<size_t>(k, best.size());
		std::partial_sort(best.begin(), best.begin() + kk, best.end());   // (dist, index) ascending
		for (uint32_t j = 0; j < k; j++) {
			nnIdx[(size_t)q * k + j] = j < kk ? best[j].second : -1;
			nnDist[(size_t)q * k + j] = j < kk ? best[j].first : std::numeric_limits<double>::infinity();
		}
	}
}

int assemble(const double* x, uint32_t m, const uint32_t* posIdx, uint32_t P, const uint32_t* negIdx, uint32_t nNegOut,
             const int32_t* nn, uint32_t localk, uint32_t fp, const int32_t* draws, double* out) {
	const size_t ld = (size_t)m + 1;
	for (uint32_t i = 0; i < P; i++) std::memcpy(out + i * ld, x + posIdx[i] * ld, ld * sizeof(double));   // samplerKNN.cpp:55-57
	if (fp > 0 && localk < 2) return -3;                            // rand() % (localk-1) with localk == 1
	const double randMax = 2147483647.0;
	size_t row = P, d = 0;
	for (uint32_t i = 0; i < P; i++) {
		const double* q = x + posIdx[i] * ld;
		for (uint32_t j = 0; j < fp; j++) {                         // samplerKNN.cpp:94-105
			uint32_t idx = (uint32_t)(draws[d++] % (int32_t)(localk - 1));
			int32_t nb = nn[(size_t)i * localk + idx];
			if (nb < 0) return -2;
			const double* t = x + posIdx[nb] * ld;
			double* o = out + row * ld;
			for (uint32_t l = 0; l < m; l++) {
				double alpha = draws[d++] / randMax;
				o[l] = t[l] * alpha + q[l] * (1 - alpha);
			}
			o[m] = 1.0;
			row++;
		}
	}
	for (uint32_t i = 0; i < nNegOut; i++) std::memcpy(out + (row + i) * ld, x + negIdx[i] * ld, ld * sizeof(double));   // :136-138
	return 0;
}

template <typename T>
double traps(const std::vector<T>& x, const std::vector<T>& y) {   // src/curves.h:119-128
	double area = 0;
	for (size_t i = 0; i + 1 < x.size(); i++) {
		if (std::isnan(y[i]) | std::isnan(y[i + 1])) continue;
		area += ((y[i] + y[i + 1]) * (x[i + 1] - x[i]) * 0.5);
	}
	return area;
}

void curves(const uint32_t* labels, const double* preds, uint64_t N, double* out, uint64_t* perm_out) {
	uint32_t totP = 0, totN = 0;
	for (uint64_t i = 0; i < N; i++) { totP += labels[i] == 1; totN += labels[i] == 0; }   // curves.cpp:16-17
	std::vector<size_t> idx(N);
	std::iota(idx.begin(), idx.end(), 0);
	std::sort(idx.begin(), idx.end(), [&](size_t i, size_t j) { return preds[i] > preds[j]; });   // curves.cpp:33
	if (perm_out) for (uint64_t i = 0; i < N; i++) perm_out[i] = idx[i];
	std::vector<uint32_t> TP(N), FP(N);
	uint32_t tp = 0, fpc = 0;
	for (uint64_t i = 0; i < N; i++) {
		uint8_t l = (uint8_t)labels[idx[i]];
		tp += l; fpc += (uint8_t)(1 - l);
		TP[i] = tp; FP[i] = fpc;
	}
	if (N) FP[N - 1] = totN;                                        // curves.cpp:87,112
	std::vector<float> fpr(N), tpr(N);
	for (uint64_t i = 0; i < N; i++) { fpr[i] = FP[i] / (float)totN; tpr[i] = TP[i] / (float)totP; }   // :116-119
	out[0] = traps<float>(fpr, tpr);
	std::vector<double> prec(N + 1), rec(N + 1);
	rec[0] = 0.0; prec[0] = 1.0;                                    // :92-93
	for (uint64_t i = 0; i < N; i++) { prec[i + 1] = TP[i] / (double)(TP[i] + FP[i]); rec[i + 1] = TP[i] / (double)totP; }
	out[1] = traps<double>(rec, prec);
}

}  // namespace

extern "C" {

void* psmo_rand_create(uint32_t seed) { return new GlibcRand(seed); }
void psmo_rand_destroy(void* st) { delete (GlibcRand*)st; }
int32_t psmo_rand_next(void* st) { return ((GlibcRand*)st)->next(); }
void psmo_rand_fill(void* st, int32_t* out, uint64_t count) { for (uint64_t i = 0; i < count; i++) out[i] = ((GlibcRand*)st)->next(); }

void* psmo_ttd_create(const uint32_t* labels, const uint32_t* foldAssign, uint32_t n, uint32_t nFolds, void* rand_state) {
	return build_ttd(labels, foldAssign, n, nFolds, (GlibcRand*)rand_state);
}
void psmo_ttd_destroy(void* t) { delete (OTTD*)t; }
void psmo_ttd_fold_sizes(void* h, uint32_t f, uint32_t* s) {
	OTTD* t = (OTTD*)h;
	s[0] = t->trngPos[f].size(); s[1] = t->trngNeg[f].size(); s[2] = t->testPos[f].size(); s[3] = t->testNeg[f].size();
}
void psmo_ttd_fold_indices(void* h, uint32_t f, uint32_t* a, uint32_t* b, uint32_t* c, uint32_t* d) {
	OTTD* t = (OTTD*)h;
	std::copy(t->trngPos[f].begin(), t->trngPos[f].end(), a); std::copy(t->trngNeg[f].begin(), t->trngNeg[f].end(), b);
	std::copy(t->testPos[f].begin(), t->testPos[f].end(), c); std::copy(t->testNeg[f].begin(), t->testNeg[f].end(), d);
}

int psmo_partition_sizes(uint32_t P, uint32_t Nneg, uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t p, uint32_t* out) {
	return partition_sizes(P, Nneg, nParts, fp, ratio, p, out);
}

void psmo_knn(const double* pts, uint32_t P, uint32_t dim, uint32_t k, int32_t* nnIdx, double* nnDist) { knn(pts, P, dim, k, nnIdx, nnDist); }

int psmo_assemble(const double* x, uint32_t m, const uint32_t* posIdx, uint32_t P, const uint32_t* negIdx, uint32_t nNegOut,
                  const int32_t* nn, uint32_t localk, uint32_t fp, const int32_t* draws, double* out) {
	return assemble(x, m, posIdx, P, negIdx, nNegOut, nn, localk, fp, draws, out);
}

void* psmo_forest_train(const double* D, uint32_t Ntr, uint32_t m, uint32_t T, uint32_t mtry, uint32_t seed) { return forest_train(D, Ntr, m, T, mtry, seed); }
void psmo_forest_destroy(void* f) { delete (OForest*)f; }
uint64_t psmo_forest_tree_nodes(void* f, uint32_t t) { return ((OForest*)f)->trees[t].var.size(); }
void psmo_forest_export_tree(void* h, uint32_t t, uint64_t* left, uint64_t* right, uint64_t* var, double* value, double* frac) {
	const OTree& tr = ((OForest*)h)->trees[t];
	for (size_t i = 0; i < tr.var.size(); i++) {
		left[i] = tr.left[i]; right[i] = tr.right[i]; var[i] = tr.var[i]; value[i] = tr.val[i];
		frac[2 * i] = tr.frac0[i]; frac[2 * i + 1] = tr.frac1[i];
	}
}
uint64_t psmo_forest_split_bytes(void* f) { return ((OForest*)f)->split_bytes; }

void psmo_tree_draws(uint32_t tree_seed, uint32_t Ntr, uint32_t m, uint32_t mtry, uint64_t nNodes, uint64_t* bootstrap, uint32_t* cand) {
	std::mt19937_64 rng;
	rng.seed(tree_seed);
	std::uniform_int_distribution<size_t> unif_dist(0, Ntr - 1);
	for (size_t s = 0; s < Ntr; ++s) bootstrap[s] = unif_dist(rng);
	std::vector<size_t> c;
	for (uint64_t i = 0; i < nNodes; i++) {
		draw_candidates(rng, m, mtry, c);
		for (uint32_t q = 0; q < mtry; q++) cand[i * mtry + q] = (uint32_t)c[q];
	}
}

void psmo_forest_predict(void* f, const double* rows, uint32_t Nte, uint32_t ld, double* preds) { forest_predict((OForest*)f, rows, Nte, ld, preds); }

void psmo_curves(const uint32_t* labels, const double* preds, uint64_t N, double* out, uint64_t* perm_out) { curves(labels, preds, N, out, perm_out); }

static int cv_impl(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
            uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
            uint32_t fold0, uint32_t fold1, uint32_t nthreads, uint32_t part0, uint32_t part1, double* class1, double* class2,
            const uint32_t* firstOfFold = nullptr, const uint32_t* lastOfFold = nullptr) {
	GlibcRand g(1);                                                  // rand() is never seeded for seed != 0 (fact #5)
	OTTD* ttd = build_ttd(labels, foldAssign, n, nFolds, &g);
	const size_t ld = (size_t)m + 1;
	int rc = 0;
	for (uint32_t f = fold0; f < fold1 && rc == 0; f++) {
		if (firstOfFold) { part0 = firstOfFold[f]; part1 = lastOfFold[f]; }   // per-fold share (unit-sharded ranks)
		const auto& tp = ttd->trngPos[f]; const auto& tn = ttd->trngNeg[f];
		const uint32_t P = tp.size();
		const uint32_t localk = std::min<uint32_t>(k, P);               // samplerKNN.cpp:52
		// kNN over the fold's positives, all m+1 columns (the label column adds 0) -- identical for every partition (fact #7)
		std::vector<double> pts((size_t)P * ld);
		for (uint32_t i = 0; i < P; i++) std::memcpy(&pts[i * ld], x + tp[i] * ld, ld * sizeof(double));
		std::vector<int32_t> nn((size_t)P * localk);
		std::vector<double> nd((size_t)P * localk);
		knn(pts.data(), P, m + 1, localk, nn.data(), nd.data());
		// test rows: positives then negatives (sampler.cpp:90-95)
		std::vector<uint32_t> testIdx(ttd->testPos[f]);
		testIdx.insert(testIdx.end(), ttd->testNeg[f].begin(), ttd->testNeg[f].end());
		const uint32_t Nte = testIdx.size();
		std::vector<double> test((size_t)Nte * ld);
		for (uint32_t i = 0; i < Nte; i++) std::memcpy(&test[i * ld], x + testIdx[i] * ld, ld * sizeof(double));
		// SMOTE draws are consumed in partition order from the one global stream
		std::vector<std::vector<int32_t>> draws(nParts);
		for (uint32_t p = 0; p < nParts; p++) {
			draws[p].resize((size_t)P * fp * (m + 1));
			for (auto& d : draws[p]) d = g.next();
		}
		std::vector<std::vector<double>> preds(nParts);
		#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
		for (int32_t p = (int32_t)part0; p < (int32_t)part1; p++) {
			uint32_t sz[6];
			int r = partition_sizes(P, tn.size(), nParts, fp, ratio, p, sz);
			if (r != 0) { rc = r; continue; }
			const uint32_t Ntr = sz[5];
			std::vector<double> rowm((size_t)Ntr * ld), colm((size_t)Ntr * ld);
			r = assemble(x, m, tp.data(), P, tn.data() + sz[2], sz[4], nn.data(), localk, fp, draws[p].data(), rowm.data());
			if (r != 0) { rc = r; continue; }
			for (size_t c = 0; c < ld; c++) for (size_t rr = 0; rr < Ntr; rr++) colm[c * Ntr + rr] = rowm[rr * ld + c];   // hyperSMURF.cpp:289-291
			uint32_t forestSeed = seed + p + f * nParts;                 // hyperSMURF.cpp:273
			OForest* fo = forest_train(colm.data(), Ntr, m, nTrees, mtry, forestSeed);
			preds[p].resize((size_t)Nte * 2);
			forest_predict(fo, test.data(), Nte, ld, preds[p].data());
			delete fo;
		}
		if (rc) break;
		const double divider = 1.0 / (double)nParts;                    // sampler.cpp:121
		for (uint32_t p = part0; p < part1; p++)
			for (uint32_t i = 0; i < Nte; i++) {
				class1[testIdx[i]] += preds[p][2 * i] * divider;
				class2[testIdx[i]] += preds[p][2 * i + 1] * divider;
			}
	}
	delete ttd;
	return rc;
}

int psmo_cv(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
            uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
            uint32_t fold0, uint32_t fold1, uint32_t nthreads, double* class1, double* class2) {
	return cv_impl(x, n, m, labels, foldAssign, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, fold0, fold1, nthreads, 0, nParts, class1, class2);
}

// only partitions [part0, part1) of every fold contribute (the share of one rank of a partition-sharded run);
// the SMOTE stream is still consumed for all partitions, so the sum over disjoint shares equals psmo_cv up to fp64 addition order
int psmo_cv_parts(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
                  uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
                  uint32_t part0, uint32_t part1, double* class1, double* class2) {
	return cv_impl(x, n, m, labels, foldAssign, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, 0, nFolds, 1, part0, part1, class1, class2);
}

// the same with a partition range per fold: [first[f], last[f]) of fold f (the share of one rank when the nFolds x nParts units
// are cut into contiguous ranges, parsmurf::hyperSMURF::partRange)
int psmo_cv_ranges(const double* x, uint32_t n, uint32_t m, const uint32_t* labels, const uint32_t* foldAssign, uint32_t nFolds,
                   uint32_t nParts, uint32_t fp, uint32_t ratio, uint32_t k, uint32_t nTrees, uint32_t mtry, uint32_t seed,
                   const uint32_t* first, const uint32_t* last, double* class1, double* class2) {
	return cv_impl(x, n, m, labels, foldAssign, nFolds, nParts, fp, ratio, k, nTrees, mtry, seed, 0, nFolds, 1, 0, 0, class1, class2, first, last);
}

}  // extern "C"
