// GPLVMHaskell-exe twin: `GPLVMHaskell-exe <file> <iters> <True|False>` (reference app/Main.hs:15-44, README.md:66-75)
// on top of libgplvm_b200.so.  Same three positional arguments, same three stdout lines:
//   line 1  show (computeS W :: Array U DIM2 Double)   ->  AUnboxed ((Z :. D) :. M) [w00,w01,...]   (row-major)
//   line 2  show logLikelihood
//   line 3  Nothing | Just (AUnboxed ((Z :. D) :. N) [...])                                   (restored matrix)
// Differences, all opt-in: the reference hard-codes the shape 4 x 150 and M = 3 (Main.hs:24-26); here the shape is
// inferred from the file (rows = observations) and `--latent M` overrides 3.  The reference seeds W0 / sigma2_0 from
// the wall clock (getStdGen, Main.hs:23); here `--seed S` (default: time) drives a normal W0 and a StdGen-like integer
// sigma2_0 in [1, 2147483562] (PPCA.hs:42-43), or `--sigma2 v` fixes it.  The True/False flag selects the typed or
// untyped Haskell variant in the reference; both run the same arithmetic (TypeSafe/PPCA.hs:91-419 vs PPCA.hs:54-182),
// so here it only enables the dimension checks of makePPCATypeSafe (TypeSafe/PPCA.hs:58-66).
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../include/gplvm_b200.h"

// Haskell `show :: Double -> String` (GHC.Float.showFloat): shortest digits that round-trip; positional notation for
// 0.1 <= |x| < 10^7, otherwise d.ddde<exp>; always at least one digit after the point.
std::string show_double(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x < 0 ? "-Infinity" : "Infinity";
  if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof buf, std::fabs(x), std::chars_format::scientific);
  std::string s(buf, res.ptr);           // d[.ddd]e[+-]XX
  const size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  const int exp10 = std::atoi(s.c_str() + epos + 1);
  std::string digits;
  for (char ch : mant)
    if (ch != '.') digits.push_back(ch);
  std::string out = x < 0 ? "-" : "";
  if (exp10 >= -1 && exp10 < 7) {        // 0.1 <= |x| < 10^7
    if (exp10 < 0) {
      out += "0." + digits;              // exp10 == -1
    } else {
      const size_t int_len = (size_t)exp10 + 1;
      if (digits.size() <= int_len) {
        out += digits + std::string(int_len - digits.size(), '0') + ".0";
      } else {
        out += digits.substr(0, int_len) + "." + digits.substr(int_len);
      }
    }
  } else {
    out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : std::string("0")) + "e" + std::to_string(exp10);
  }
  return out;
}

static std::string show_array(const std::vector<double>& a, int64_t rows, int64_t cols) {
  std::string s = "AUnboxed ((Z :. " + std::to_string(rows) + ") :. " + std::to_string(cols) + ") [";
  for (size_t i = 0; i < a.size(); ++i) {
    if (i) s += ",";
    s += show_double(a[i]);
  }
  s += "]";
  return s;
}

// parseTestData (Main.hs:43-44): rows separated by "\n" | "\r\n", fields by spaces, a field is a double or "NaN".
static bool parse_file(const std::string& path, std::vector<std::vector<double>>& rows) {
  std::ifstream in(path);
  if (!in) return false;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::vector<double> r;
    std::string tok;
    while (ls >> tok) {
      if (tok == "NaN") {
        r.push_back(std::nan(""));
      } else {
        char* end = nullptr;
        const double v = std::strtod(tok.c_str(), &end);
        if (end == tok.c_str() || *end != '\0') return false;
        r.push_back(v);
      }
    }
    if (!r.empty()) rows.push_back(r);   // an empty trailing row is dropped, like Data.List.transpose does
  }
  return true;
}

int main(int argc, char** argv) {
  std::vector<std::string> pos;
  int64_t latent = 3;
  uint64_t seed = (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();
  double sigma2 = -1.0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--selftest-format") {
      const double t[] = {0.1, 1.0, 5.0, -1.5, 0.05, 1.0e-2, 123.456, 1.2345e7, 9999999.0, 1.0e7, 0.023676192353627, -381.92805336,
                          5.1, 3.0e-3, 1.0 / 3.0, 2147483562.0, 1e22, 5e-324};
      for (double v : t) std::printf("%s\n", show_double(v).c_str());
      return 0;
    }
    if (a == "--latent" && i + 1 < argc) latent = std::atoll(argv[++i]);
    else if (a == "--seed" && i + 1 < argc) seed = std::strtoull(argv[++i], nullptr, 10);
    else if (a == "--sigma2" && i + 1 < argc) sigma2 = std::atof(argv[++i]);
    else pos.push_back(a);
  }
  if (pos.size() != 3) {   // the reference dies on an incomplete pattern match here (Main.hs:17)
    std::fprintf(stderr, "usage: GPLVMHaskell-exe <file> <iters> <True|False> [--latent M] [--seed S] [--sigma2 v]\n");
    return 1;
  }
  char* end = nullptr;
  const long long iters = std::strtoll(pos[1].c_str(), &end, 10);
  if (end == pos[1].c_str() || *end) { std::fprintf(stderr, "Prelude.read: no parse\n"); return 1; }
  if (pos[2] != "True" && pos[2] != "False") { std::fprintf(stderr, "Prelude.read: no parse\n"); return 1; }
  const bool type_safe = pos[2] == "True";

  std::vector<std::vector<double>> rows;
  if (!parse_file(pos[0], rows) || rows.empty()) { std::fprintf(stderr, "cannot parse %s\n", pos[0].c_str()); return 1; }
  const int64_t N = (int64_t)rows.size(), D = (int64_t)rows[0].size();
  for (auto& r : rows)
    if ((int64_t)r.size() != D) { std::fprintf(stderr, "ragged input\n"); return 1; }
  std::vector<double> Y((size_t)D * N);   // transpose -> D x N feature-major (Main.hs:22-24)
  for (int64_t n = 0; n < N; ++n)
    for (int64_t d = 0; d < D; ++d) Y[(size_t)d * N + n] = rows[n][d];

  std::mt19937_64 gen(seed);
  std::normal_distribution<double> normal(0.0, 1.0);
  std::vector<double> W0((size_t)D * latent);
  for (double& w : W0) w = normal(gen);                                  // randomMatrixD (Util.hs:139-151)
  if (sigma2 <= 0.0) sigma2 = (double)(1 + gen() % 2147483562ULL);        // fromIntegral (fst (next gen)), PPCA.hs:43
  if (type_safe) {  // checkPPCAInputMatricesProp (TypeSafe/PPCA.hs:58-66,72-79): x2 == d, y1 == y2 hold by construction
    if (N < 1) { std::fprintf(stderr, "Input matrix should have at least one column\n"); return 1; }
  }
  std::vector<double> W((size_t)D * latent), restored((size_t)D * N);
  double s2 = 0, ll = 0;
  int no_missed = 0;
  int64_t steps = 0;
  const int rc = gplvm_ppca(Y.data(), D, N, latent, W0.data(), sigma2, GPLVM_STOP_ITERATIONS, iters, 0.0, W.data(), &s2, &ll,
                            restored.data(), &no_missed, &steps);
  if (rc != GPLVM_OK) { std::fprintf(stderr, "GPLVMHaskell-exe: %s\n", gplvm_last_error()); return 1; }
  std::printf("%s\n", show_array(W, D, latent).c_str());
  std::printf("%s\n", show_double(ll).c_str());
  if (no_missed) std::printf("Nothing\n");
  else std::printf("Just (%s)\n", show_array(restored, D, N).c_str());
  return 0;
}
