// parsmurf_b200_cv -- command-line driver with parSMURF1's configuration surface:  parsmurf_b200_cv --cfg file.json
// (reference src/parSMURF.cpp:23-140, src/ArgHandler_new.cpp:19-82).  Data files are read by parsmurf::Importer (the reference's
// text format, or the PSMB1 binary side channel), predictions are written by parsmurf::saveToFile (ingest.cpp).
// --gpus N runs the job on N GPUs of the node, one host thread ("rank") per GPU, the way parSMURFn spreads it over MPI ranks
// (src/parSMURFn.cpp, src/hyperSMURF_MPI.cpp:594-646): the dataset is replicated, the partitions are sharded and the score
// vectors are summed by the library's own NCCL communicator (psm_comm_*); rank 0 plays the master and writes the predictions.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <thread>

#include "parsmurf_host.h"

using namespace parsmurf;

int main(int argc, char** argv) {
	std::string cfgPath; int device = 0, tieMode = 2, gpus = 1;   // 2 = auto: device sort, host std::sort only when a run of tied scores holds both labels
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--cfg") && i + 1 < argc) cfgPath = argv[++i];
		else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--tie-mode") && i + 1 < argc) tieMode = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::max(1, atoi(argv[++i]));
		else if (!strcmp(argv[i], "--help")) { printf("usage: %s --cfg <file.json> [--device N] [--gpus N] [--tie-mode 0|1|2]\n", argv[0]); return 0; }
	}
	if (cfgPath.empty()) { fprintf(stderr, "usage: %s --cfg <file.json>\n", argv[0]); return 2; }
	try {
		Config cfg = Config::fromFile(cfgPath);
		GlibcRand rng(1);
		if (cfg.seedFromClock) rng.srand(cfg.common.seed);                   // the only case in which the reference seeds rand() (ArgHandler_new.cpp:257-262)
		std::vector<double> x; std::vector<uint32_t> y, ff;
		uint32_t n = 0, m = 0;
		if (cfg.simulate) {
			n = cfg.common.nn; m = cfg.common.mm;
			generateRandomSet(n, m, x, y, cfg.prob, cfg.common.seed);          // src/parSMURF.cpp:77
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
		double auroc = 0, auprc = 0, secs = 0;
		const bool fromFile = !(cfg.simulate || cfg.foldFilename.empty());
		if (gpus <= 1) {
			Device dev(device);
			dev.upload(x.data(), n, m);
			Folds folds(cfg.common.nFolds, n, y.data(), ff, !fromFile, &rng);
			folds.computeFolds();
			hyperSMURF smurfer(&cfg.common, cfg.grid, &folds, &dev, y, class1.data(), class2.data(), &rng);
			smurfer.curvesTieMode = tieMode;
			smurfer.initStandard();
			auto t0 = std::chrono::high_resolution_clock::now();
			smurfer.smurfIt();
			secs = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
			auroc = smurfer.auroc; auprc = smurfer.auprc;
		} else {
			// one rank per GPU [device, device + gpus): every rank replays the same rand() stream on its own copy of the host
			// state (folds, grid table, score arrays), like the MPI ranks of parSMURFn each parsing the same configuration
			unsigned char id[PSM_COMM_ID_BYTES];
			Device::commUniqueId(id);
			std::vector<std::string> errs(gpus);
			std::vector<double> rankSecs(gpus, 0.0);
			auto rankMain = [&](int r) {
				try {
					GlibcRand myRng = rng;
					std::vector<GridParams> myGrid = cfg.grid;
					std::vector<uint32_t> myFf = ff;
					std::vector<double> c1(r == 0 ? 0 : n, 0.0), c2(r == 0 ? 0 : n, 0.0);
					Device dev(device + r);
					dev.commInit(id, (uint32_t)r, (uint32_t)gpus);           // collective: returns when every rank has joined
					dev.upload(x.data(), n, m);
					Folds folds(cfg.common.nFolds, n, y.data(), myFf, !fromFile, &myRng);
					folds.computeFolds();
					hyperSMURF smurfer(&cfg.common, myGrid, &folds, &dev, y, r == 0 ? class1.data() : c1.data(), r == 0 ? class2.data() : c2.data(), &myRng);
					smurfer.curvesTieMode = tieMode;
					smurfer.shardUnits = true;
					smurfer.scoresOnRootOnly = true;
					smurfer.setSharding((uint32_t)r, (uint32_t)gpus, [&dev](double* p, uint64_t c) { dev.allReduce(p, c); });
					smurfer.initStandard();
					auto t0 = std::chrono::high_resolution_clock::now();
					smurfer.smurfIt();
					rankSecs[r] = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
					if (r == 0) { auroc = smurfer.auroc; auprc = smurfer.auprc; }
				} catch (const std::exception& e) { errs[r] = e.what(); if (errs[r].empty()) errs[r] = "failed"; }
			};
			std::vector<std::thread> th;
			for (int r = 0; r < gpus; r++) th.emplace_back(rankMain, r);
			for (auto& t : th) t.join();
			for (int r = 0; r < gpus; r++) if (!errs[r].empty()) { fprintf(stderr, "error on rank %d: %s\n", r, errs[r].c_str()); return 1; }
			for (double v : rankSecs) secs = std::max(secs, v);
		}
		if (cfg.common.wmode == MODE_TRAIN) {                                // src/parSMURF.cpp:128: only cv / predict write predictions
			printf("forests saved to %s  (%u partitions, %u examples, %u features, %.3f s)\n", cfg.common.forestDirname.c_str(), cfg.grid[0].nParts, n, m, secs);
			if (!cfg.common.timeFilename.empty()) { std::ofstream tf(cfg.common.timeFilename.c_str(), std::ios::app); tf << secs << std::endl; }
			return 0;
		}
		printf("AUROC: %.9g - AUPRC: %.9g  (%u examples, %u features, %d GPU%s, %.3f s)\n", auroc, auprc, n, m, gpus, gpus > 1 ? "s" : "", secs);
		if (!cfg.common.timeFilename.empty()) { std::ofstream tf(cfg.common.timeFilename.c_str(), std::ios::app); tf << secs << std::endl; }
		std::vector<uint32_t> none;                                          // src/parSMURF.cpp:128-137: the label column only for simulated data
		saveToFile(class1.data(), class2.data(), n, cfg.simulate ? &y : &none, cfg.common.outfilename);
		return 0;
	} catch (const Error& e) {
		fprintf(stderr, "error (%d): %s\n", e.code, e.what());
		return 1;
	}
}
