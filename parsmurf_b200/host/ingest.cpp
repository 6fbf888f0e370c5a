// parsmurf_b200 host layer -- data ingest and predictions writer (once the partition pipeline runs in
// tens of milliseconds, parsing n*m decimal numbers and printing 2n of them is where a run spends its time).
//
//   Importer::import   == reference Importer::import (src/fileImport.cpp:5-100): label file, optional fold file, headerless
//                         data file whose FIRST LINE fixes the number of features (tokens split at " ,\t"); every value is read
//                         like `stream >> double` / `stream >> uint32_t` (white-space separated; reading stops at the first
//                         token that is not a number, as the reference's extraction loop does); after every m values the
//                         label column 1.0 (label > 0) / 2.0 is appended (:88-91).  Same values bit for bit, but the file is
//                         read in one piece and parsed by several threads with std::from_chars.
//   saveToFile         == reference saveToFile (src/HyperSMURFUtils.cpp:66-86): "cl1<TAB>cl2[<TAB>label]\n" with the default
//                         ostream formatting of a double (%g, 6 significant digits); byte-identical output, formatted in
//                         parallel and written with one call.
//   binary side channel: a data file that starts with the magic "PSMB1" holds x (features only, f64 row-major), labels and
//                         optional fold ids in binary; an output name ending in ".bin" gets the two score arrays as raw f64.
#include "parsmurf_host.h"

#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>

namespace parsmurf {

namespace {

std::string slurp(const std::string& path, const char* what) {
	std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
	if (!f) throw Error(PSM_E_ARG, std::string("Error opening ") + what + " file.");        // fileImport.cpp:17,32,56
	f.seekg(0, std::ios::end);
	const std::streamoff len = f.tellg();
	f.seekg(0, std::ios::beg);
	std::string s((size_t)std::max<std::streamoff>(len, 0), '\0');
	if (len > 0) f.read(&s[0], len);
	return s;
}

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// Parses the white-space separated numbers of text[b, e) (b and e on token boundaries) into out.  Returns false when a token
// is not entirely a number; the value of its numeric prefix, if any, is still appended (operator>> consumes the prefix, the
// NEXT extraction fails).
template <typename T>
bool parse_range(const char* p, const char* e, std::vector<T>& out) {
	while (p < e) {
		while (p < e && is_space(*p)) p++;
		if (p >= e) break;
		const char* t = p;
		while (p < e && !is_space(*p)) p++;
		const char* te = p;
		const char* q = t;
		if (q < te && *q == '+' && q + 1 < te && (q[1] == '.' || (q[1] >= '0' && q[1] <= '9'))) q++;   // from_chars takes no leading '+'
		T v{};
		const std::from_chars_result r = std::from_chars(q, te, v);
		if (r.ec == std::errc::result_out_of_range) {
			if constexpr (std::is_floating_point<T>::value) {              // strtod semantics, which operator>> uses: +-HUGE_VAL / 0
				v = (T)strtod(std::string(t, te).c_str(), nullptr);
				out.push_back(v);
				if (r.ptr != te) return false;
				continue;
			}
			return false;
		}
		if (r.ec != std::errc() || r.ptr == q) return false;
		out.push_back(v);
		if (r.ptr != te) return false;
	}
	return true;
}

unsigned pick_threads(size_t bytes, unsigned threads) {
	if (threads == 0) threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
	const size_t byWork = std::max<size_t>(1, bytes >> 20);              // about a megabyte of text per thread at least
	return (unsigned)std::min<size_t>(threads, byWork);
}

template <typename T>
std::vector<T> parse_all(const std::string& text, unsigned threads) {
	const unsigned nT = pick_threads(text.size(), threads);
	const char* base = text.data();
	const size_t len = text.size();
	std::vector<size_t> cut(nT + 1, len);
	cut[0] = 0;
	for (unsigned t = 1; t < nT; t++) {                                      // move every cut forward to the next white space
		size_t c = std::max(cut[t - 1], len * t / nT);
		while (c < len && !is_space(base[c])) c++;
		cut[t] = c;
	}
	std::vector<std::vector<T>> part(nT);
	std::vector<char> ok(nT, 1);
	auto work = [&](unsigned t) {
		part[t].reserve((cut[t + 1] - cut[t]) / 4 + 16);
		ok[t] = parse_range<T>(base + cut[t], base + cut[t + 1], part[t]) ? 1 : 0;
	};
	if (nT == 1) work(0);
	else {
		std::vector<std::thread> th;
		for (unsigned t = 0; t < nT; t++) th.emplace_back(work, t);
		for (auto& x : th) x.join();
	}
	size_t total = 0;
	unsigned last = nT;
	for (unsigned t = 0; t < nT; t++) { total += part[t].size(); if (!ok[t]) { last = t + 1; break; } }   // everything after the first bad token is ignored
	std::vector<T> out;
	out.reserve(total);
	for (unsigned t = 0; t < std::min(last, nT); t++) out.insert(out.end(), part[t].begin(), part[t].end());
	return out;
}

const char kMagicData[8] = {'P', 'S', 'M', 'B', '1', 0, 0, 0};
const char kMagicPred[8] = {'P', 'S', 'M', 'P', '1', 0, 0, 0};

}  // namespace

bool Importer::isBinary(const std::string& dataFilename) {
	std::ifstream f(dataFilename.c_str(), std::ios::in | std::ios::binary);
	char m[8] = {0};
	return f && f.read(m, 8) && memcmp(m, kMagicData, 8) == 0;
}

void Importer::import(const std::string& dataFilename, const std::string& labelFilename, const std::string& foldFilename, std::vector<double>& x,
                      std::vector<uint32_t>& y, std::vector<uint32_t>& f, uint32_t* nFolds, uint32_t* mOut, unsigned threads) {
	if (isBinary(dataFilename)) { importBinary(dataFilename, x, y, f, nFolds, mOut); return; }
	y = parse_all<uint32_t>(slurp(labelFilename, "label"), threads);          // fileImport.cpp:22-25
	const size_t n = y.size();
	f.clear();
	if (!foldFilename.empty()) {                                             // :31-48
		f = parse_all<uint32_t>(slurp(foldFilename, "fold"), threads);
		uint32_t mx = 0;
		for (uint32_t v : f) mx = std::max(mx, v);
		if (nFolds) *nFolds = mx + 1;
		if (f.size() != n) throw Error(PSM_E_ARG, "size mismatch between label and fold file");   // the reference only warns
	}
	const std::string text = slurp(dataFilename, "matrix");
	// number of features = tokens of the first line, split at blank, comma or tab (:62-76)
	size_t eol = text.find('\n');
	if (eol == std::string::npos) eol = text.size();
	size_t m = 0;
	for (size_t i = 0; i < eol;) {
		while (i < eol && (text[i] == ' ' || text[i] == ',' || text[i] == '\t')) i++;
		if (i >= eol) break;
		m++;
		while (i < eol && !(text[i] == ' ' || text[i] == ',' || text[i] == '\t')) i++;
	}
	if (m == 0) throw Error(PSM_E_ARG, "no features detected in the data file");
	const std::vector<double> d = parse_all<double>(text, threads);           // :81-92
	if (n == 0 || d.size() % n != 0 || d.size() / n != m) throw Error(PSM_E_ARG, "size mismatch between label and data file");   // the reference only warns
	x.resize(n * (m + 1));
	for (size_t i = 0; i < n; i++) {
		memcpy(&x[i * (m + 1)], &d[i * m], m * sizeof(double));
		x[i * (m + 1) + m] = y[i] > 0 ? 1.0 : 2.0;                           // :88-91
	}
	if (mOut) *mOut = (uint32_t)m;
}

// "PSMB1\0\0\0" | u32 n | u32 m | u32 hasFolds | u32 0 | f64 x[n][m] | u32 y[n] | u32 folds[n] (if hasFolds)
void Importer::exportBinary(const std::string& path, const double* xRowMajorWithLabel, uint32_t n, uint32_t m, const uint32_t* y, const uint32_t* folds) {
	std::ofstream o(path.c_str(), std::ios::out | std::ios::binary);
	if (!o) throw Error(PSM_E_ARG, "cannot write " + path);
	const uint32_t hdr[4] = {n, m, folds ? 1u : 0u, 0u};
	o.write(kMagicData, 8); o.write((const char*)hdr, 16);
	for (uint32_t i = 0; i < n; i++) o.write((const char*)(xRowMajorWithLabel + (size_t)i * (m + 1)), (std::streamsize)m * 8);
	o.write((const char*)y, (std::streamsize)n * 4);
	if (folds) o.write((const char*)folds, (std::streamsize)n * 4);
}

void Importer::importBinary(const std::string& path, std::vector<double>& x, std::vector<uint32_t>& y, std::vector<uint32_t>& f, uint32_t* nFolds, uint32_t* mOut) {
	std::ifstream in(path.c_str(), std::ios::in | std::ios::binary);
	char magic[8]; uint32_t hdr[4];
	if (!in || !in.read(magic, 8) || memcmp(magic, kMagicData, 8) != 0 || !in.read((char*)hdr, 16)) throw Error(PSM_E_ARG, path + " is not a PSMB1 data file");
	const size_t n = hdr[0], m = hdr[1];
	if (n == 0 || m == 0) throw Error(PSM_E_ARG, path + ": empty data set");
	std::vector<double> d(n * m);
	y.resize(n); f.clear();
	if (!in.read((char*)d.data(), (std::streamsize)(n * m * 8)) || !in.read((char*)y.data(), (std::streamsize)(n * 4))) throw Error(PSM_E_ARG, path + ": truncated");
	if (hdr[2]) {
		f.resize(n);
		if (!in.read((char*)f.data(), (std::streamsize)(n * 4))) throw Error(PSM_E_ARG, path + ": truncated");
		if (nFolds) *nFolds = *std::max_element(f.begin(), f.end()) + 1;
	}
	x.resize(n * (m + 1));
	for (size_t i = 0; i < n; i++) {
		memcpy(&x[i * (m + 1)], &d[i * m], m * sizeof(double));
		x[i * (m + 1) + m] = y[i] > 0 ? 1.0 : 2.0;
	}
	if (mOut) *mOut = (uint32_t)m;
}

void saveToFile(const double* cl1, const double* cl2, uint32_t nn, const std::vector<uint32_t>* labels, const std::string& outFilename, unsigned threads) {
	const bool withLabels = labels && !labels->empty();
	const bool binary = outFilename.size() > 4 && outFilename.compare(outFilename.size() - 4, 4, ".bin") == 0;
	FILE* fp = fopen(outFilename.c_str(), "wb");
	if (!fp) throw Error(PSM_E_ARG, "cannot write " + outFilename);
	if (binary) {                                                            // "PSMP1\0\0\0" | u32 n | u32 0 | f64 cl1[n] | f64 cl2[n]
		const uint32_t hdr[2] = {nn, 0};
		fwrite(kMagicPred, 1, 8, fp); fwrite(hdr, 4, 2, fp);
		fwrite(cl1, 8, nn, fp); fwrite(cl2, 8, nn, fp);
		fclose(fp);
		return;
	}
	const unsigned nT = pick_threads((size_t)nn * 24, threads);
	std::vector<std::string> part(nT);
	auto work = [&](unsigned t) {
		const uint32_t a = (uint32_t)((uint64_t)nn * t / nT), b = (uint32_t)((uint64_t)nn * (t + 1) / nT);
		std::string& s = part[t];
		s.reserve((size_t)(b - a) * 28);
		char buf[96];
		for (uint32_t i = a; i < b; i++) {                                   // outFile << cl1[i] << "\t" << cl2[i] [<< "\t" << label] << std::endl
			int k = withLabels ? snprintf(buf, sizeof buf, "%g\t%g\t%u\n", cl1[i], cl2[i], (*labels)[i]) : snprintf(buf, sizeof buf, "%g\t%g\n", cl1[i], cl2[i]);
			s.append(buf, (size_t)k);
		}
	};
	if (nT == 1) work(0);
	else {
		std::vector<std::thread> th;
		for (unsigned t = 0; t < nT; t++) th.emplace_back(work, t);
		for (auto& x : th) x.join();
	}
	for (const std::string& s : part) fwrite(s.data(), 1, s.size(), fp);
	fclose(fp);
}

}  // namespace parsmurf

// ------------------------------------------------------------------------------------------ C entry points (tests)
extern "C" {

// sizes = {x.size(), n, f.size(), nFolds, m}; first call with x == NULL for the sizes
int psmh_import(const char* dataFile, const char* labelFile, const char* foldFile, unsigned threads, uint64_t* sizes, double* x, uint32_t* y, uint32_t* f,
                char* errbuf, size_t errlen) {
	try {
		std::vector<double> xx; std::vector<uint32_t> yy, ff;
		uint32_t nFolds = 0, m = 0;
		parsmurf::Importer::import(dataFile, labelFile ? labelFile : "", foldFile ? foldFile : "", xx, yy, ff, &nFolds, &m, threads);
		sizes[0] = xx.size(); sizes[1] = yy.size(); sizes[2] = ff.size(); sizes[3] = nFolds; sizes[4] = m;
		if (x) memcpy(x, xx.data(), xx.size() * 8);
		if (y) memcpy(y, yy.data(), yy.size() * 4);
		if (f) memcpy(f, ff.data(), ff.size() * 4);
		return 0;
	} catch (const std::exception& e) {
		if (errbuf && errlen) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
		return -1;
	}
}

int psmh_save(const double* c1, const double* c2, uint32_t n, const uint32_t* labels, const char* path, unsigned threads) {
	try {
		std::vector<uint32_t> lab;
		if (labels) lab.assign(labels, labels + n);
		parsmurf::saveToFile(c1, c2, n, &lab, path, threads);
		return 0;
	} catch (...) { return -1; }
}

int psmh_export_binary(const char* path, const double* xWithLabel, uint32_t n, uint32_t m, const uint32_t* y, const uint32_t* folds) {
	try { parsmurf::Importer::exportBinary(path, xWithLabel, n, m, y, folds); return 0; } catch (...) { return -1; }
}

}  // extern "C"
