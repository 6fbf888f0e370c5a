// parsmurf_b200_cv -- command-line driver with parSMURF1's configuration surface:  parsmurf_b200_cv --cfg file.json
// (reference src/parSMURF.cpp:23-140, src/ArgHandler_new.cpp:19-82).  Text ingest is deliberately minimal
// (whitespace/comma separated; one example per line) -- data import is outside the accelerated path.
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

static std::vector<double> read_numbers(const std::string& path, size_t* firstLineCols) {
	std::ifstream f(path.c_str());
	if (!f) throw Error(PSM_E_ARG, "cannot open " + path);
	std::vector<double> out;
	std::string line;
	bool first = true;
	while (std::getline(f, line)) {
		for (char& c : line) if (c == ',' || c == '\t') c = ' ';
		std::istringstream ss(line);
		double v; size_t cols = 0;
		while (ss >> v) { out.push_back(v); cols++; }
		if (first && cols) { if (firstLineCols) *firstLineCols = cols; first = false; }
	}
	return out;
}

int main(int argc, char** argv) {
	std::string cfgPath; int device = 0, tieMode = 0;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--cfg") && i + 1 < argc) cfgPath = argv[++i];
		else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--tie-mode") && i + 1 < argc) tieMode = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--help")) { printf("usage: %s --cfg <file.json> [--device N] [--tie-mode 0|1]\n", argv[0]); return 0; }
	}
	if (cfgPath.empty()) { fprintf(stderr, "usage: %s --cfg <file.json>\n", argv[0]); return 2; }
	try {
		Config cfg = Config::fromFile(cfgPath);
		GlibcRand rng(1);
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
			std::vector<double> lab = read_numbers(cfg.labelFilename, nullptr);
			n = (uint32_t)lab.size();
			y.resize(n);
			for (uint32_t i = 0; i < n; i++) y[i] = lab[i] > 0;
			size_t cols = 0;
			std::vector<double> d = read_numbers(cfg.dataFilename, &cols);
			if (n == 0 || d.size() % n) throw Error(PSM_E_ARG, "size mismatch between label and data file");
			m = (uint32_t)(d.size() / n);
			x.resize((size_t)n * (m + 1));
			for (uint32_t i = 0; i < n; i++) {
				memcpy(&x[(size_t)i * (m + 1)], &d[(size_t)i * m], m * sizeof(double));
				x[(size_t)i * (m + 1) + m] = y[i] ? 1.0 : 2.0;               // src/fileImport.cpp:88-91
			}
			if (!cfg.foldFilename.empty()) {
				std::vector<double> fv = read_numbers(cfg.foldFilename, nullptr);
				if (fv.size() != n) throw Error(PSM_E_ARG, "fold file size mismatch");
				ff.resize(n);
				for (uint32_t i = 0; i < n; i++) ff[i] = (uint32_t)fv[i];
				cfg.common.nFolds = *std::max_element(ff.begin(), ff.end()) + 1;
			}
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
		std::ofstream of(cfg.common.outfilename.c_str());                    // src/HyperSMURFUtils.cpp:66-86
		for (uint32_t i = 0; i < n; i++) of << class1[i] << "\t" << class2[i] << "\t" << y[i] << std::endl;
		return 0;
	} catch (const Error& e) {
		fprintf(stderr, "error (%d): %s\n", e.code, e.what());
		return 1;
	}
}
