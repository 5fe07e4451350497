// TEST INFRASTRUCTURE.  Host-side analogue of the reference's cpp_thread_test/dgemm_thread_safety.cpp
// for the host-simulation build: THREADS threads call cblas_dgemm / cblas_dsyrk / cblas_dtrsm /
// cblas_dgemm_batch concurrently on identical inputs (sizes that take the packed, strided and
// pipelined staging paths); every thread's results must be byte-identical to thread 0's.  Built with
// -fsanitize=thread (tests/test_hostsim.py) it also checks the context pool, the launch counter and
// the pinned-slot pool for data races.
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "openblas_b200.h"

static void fill(std::vector<double> &v, unsigned seed) {
  unsigned long long s = seed * 2654435761ull + 12345;
  for (auto &x : v) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = (double)(s >> 11) / 9007199254740992.0 - 0.5; }
}

int main() {
  const int THREADS = 8, ROUNDS = 6;
  const int sizes[3][3] = {{20, 18, 16}, {70, 60, 50}, {150, 130, 40}};
  std::vector<std::vector<double>> results(THREADS);
  std::vector<std::thread> th;
  for (int t = 0; t < THREADS; t++)
    th.emplace_back([&, t] {
      std::vector<double> &out = results[t];
      for (int r = 0; r < ROUNDS; r++)
        for (auto &sz : sizes) {
          const int m = sz[0], n = sz[1], k = sz[2];
          std::vector<double> a((size_t)m * k), b((size_t)k * n), c((size_t)m * n), s((size_t)n * n), tri((size_t)m * m), x((size_t)m * n);
          fill(a, 1 + r); fill(b, 2 + r); fill(c, 3 + r); fill(s, 4 + r); fill(tri, 5 + r); fill(x, 6 + r);
          for (int i = 0; i < m; i++) tri[(size_t)i * m + i] += 4.0;
          cblas_dgemm(CblasColMajor, CblasNoTrans, CblasTrans, m, n, k, 0.7, a.data(), m, b.data(), n, 1.3, c.data(), m);
          cblas_dsyrk(CblasColMajor, CblasLower, CblasTrans, n, k, 0.7, b.data(), k, 1.3, s.data(), n);
          cblas_dtrsm(CblasColMajor, CblasLeft, CblasUpper, CblasNoTrans, CblasNonUnit, m, n, 0.7, tri.data(), m, x.data(), m);
          const double *ap[2] = {a.data(), a.data()}, *bp[2] = {b.data(), b.data()};
          std::vector<double> c2(c), c3(c);
          double *cp[2] = {c2.data(), c3.data()};
          enum CBLAS_TRANSPOSE ta[1] = {CblasNoTrans}, tb[1] = {CblasTrans};
          blasint mm[1] = {m}, nn[1] = {n}, kk[1] = {k}, lda[1] = {m}, ldb[1] = {n}, ldc[1] = {m}, gs[1] = {2};
          double al[1] = {0.5}, be[1] = {0.25};
          cblas_dgemm_batch(CblasColMajor, ta, tb, mm, nn, kk, al, ap, lda, bp, ldb, be, cp, ldc, 1, gs);
          out.insert(out.end(), c.begin(), c.end());
          out.insert(out.end(), s.begin(), s.end());
          out.insert(out.end(), x.begin(), x.end());
          out.insert(out.end(), c2.begin(), c2.end());
          out.insert(out.end(), c3.begin(), c3.end());
        }
    });
  for (auto &t : th) t.join();
  int bad = 0;
  for (int t = 1; t < THREADS; t++)
    if (results[t].size() != results[0].size() || memcmp(results[t].data(), results[0].data(), results[0].size() * sizeof(double))) bad++;
  double sum = 0;
  for (double v : results[0]) sum += v;
  printf(bad ? "THREADS: %d of %d differ from thread 0\n" : "THREADS OK (%d differ) of %d, checksum %.12g\n", bad, THREADS, sum);
  return bad != 0;
}
