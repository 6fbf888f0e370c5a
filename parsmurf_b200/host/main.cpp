// parsmurf_b200_cv -- command-line driver with parSMURF1's configuration surface:  parsmurf_b200_cv --cfg file.json
// (reference src/parSMURF.cpp:23-140, src/ArgHandler_new.cpp:19-82).  Data files are read by parsmurf::Importer (the reference's
// text format, or the PSMB1 binary side channel), predictions are written by parsmurf::saveToFile (ingest.cpp).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>

#include "parsmurf_host.h"

using namespace parsmurf;

int main(int argc, char** argv) {
	std::string cfgPath; int device = 0, tieMode = 2;   // 2 = auto: device sort, host std::sort only when a run of tied scores holds both labels
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--cfg") && i + 1 < argc) cfgPath = argv[++i];
		else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--tie-mode") && i + 1 < argc) tieMode = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--help")) { printf("usage: %s --cfg <file.json> [--device N] [--tie-mode 0|1|2]\n", argv[0]); return 0; }
	}
	if (cfgPath.empty()) { fprintf(stderr, "usage: %s --cfg <file.json>\n", argv[0]); return 2; }
	try {
		Config cfg = Config::fromFile(cfgPath);
		GlibcRand rng(1);
		if (cfg.seedFromClock) rng.srand(cfg.common.seed);                   // the only case in which the reference seeds rand() (ArgHandler_new.cpp:257-262)
		std::vector<double> x; std::vector<uint32_t> y, ff;
		uint32_t n = 0, m = 0;
		if (cfg.simulate) {
			// synthetic N(0,1) negatives / N(5,1) positives like generateRandomSet (src/HyperSMURFUtils.cpp:40-63); engine-specific stream
			n = cfg.common.nn; m = cfg.common.mm;
			std::mt19937_64 gen(cfg.common.seed);
			std::normal_distribution<> neg(0, 1), pos(5, 1);
			std::bernoulli_distribution lab(cfg.prob);
			y.resize(n); x.resize((size_t)n * (m + 1));
			for (auto& v : y) v = lab(gen);
			for (uint32_t i = 0; i < n; i++) {
				for (uint32_t j = 0; j < m; j++) x[(size_t)i * (m + 1) + j] = y[i] ? pos(gen) : neg(gen);
				x[(size_t)i * (m + 1) + m] = y[i] ? 1.0 : 2.0;
			}
		} else {
			uint32_t nFoldsFromFile = 0;
			Importer::import(cfg.dataFilename, cfg.labelFilename, cfg.foldFilename, x, y, ff, &nFoldsFromFile, &m);   // src/parSMURF.cpp:68-77
			n = (uint32_t)y.size();
			for (auto& v : y) v = v > 0;                                     // labels are used as 0 / 1 from here on
			if (!ff.empty()) cfg.common.nFolds = nFoldsFromFile;
		}
		cfg.common.nn = n; cfg.common.mm = m;
		cfg.processMtry(m);
		std::vector<double> class1(n, 0.0), class2(n, 0.0);
		Device dev(device);
		dev.upload(x.data(), n, m);
		Folds folds(cfg.common.nFolds, n, y.data(), ff, cfg.simulate || cfg.foldFilename.empty(), &rng);
		folds.computeFolds();
		hyperSMURF smurfer(&cfg.common, cfg.grid, &folds, &dev, y, class1.data(), class2.data(), &rng);
		smurfer.curvesTieMode = tieMode;
		smurfer.initStandard();
		auto t0 = std::chrono::high_resolution_clock::now();
		smurfer.smurfIt();
		auto t1 = std::chrono::high_resolution_clock::now();
		double secs = std::chrono::duration<double>(t1 - t0).count();
		if (cfg.common.wmode == MODE_TRAIN) {                                // src/parSMURF.cpp:128: only cv / predict write predictions
			printf("forests saved to %s  (%u partitions, %u examples, %u features, %.3f s)\n", cfg.common.forestDirname.c_str(), cfg.grid[0].nParts, n, m, secs);
			if (!cfg.common.timeFilename.empty()) { std::ofstream tf(cfg.common.timeFilename.c_str(), std::ios::app); tf << secs << std::endl; }
			return 0;
		}
		printf("AUROC: %.9g - AUPRC: %.9g  (%u examples, %u features, %.3f s)\n", smurfer.auroc, smurfer.auprc, n, m, secs);
		if (!cfg.common.timeFilename.empty()) { std::ofstream tf(cfg.common.timeFilename.c_str(), std::ios::app); tf << secs << std::endl; }
		std::vector<uint32_t> none;                                          // src/parSMURF.cpp:128-137: the label column only for simulated data
		saveToFile(class1.data(), class2.data(), n, cfg.simulate ? &y : &none, cfg.common.outfilename);
		return 0;
	} catch (const Error& e) {
		fprintf(stderr, "error (%d): %s\n", e.code, e.what());
		return 1;
	}
}
