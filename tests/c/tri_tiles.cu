// Host-side check of the closed-form triangle tile enumeration the GEMM kernels use for the SYRK
// family (gemm_common.cuh: tri_tile_count / tri_tile_coords / tri_outside / tri_partial / tri_keep).
#include <cstdio>
#include <vector>
#include "gemm_common.cuh"
using namespace b200;

int main() {
  int failures = 0;
  for (int tri = 1; tri <= 2; tri++)
    for (int64_t nt = 1; nt <= 200; nt++) {
      std::vector<char> seen((size_t)(nt * nt), 0);
      const int64_t T = tri_tile_count(nt);
      for (int64_t u = 0; u < T; u++) {
        int64_t bm, bn;
        tri_tile_coords(u, tri, bm, bn);
        const bool inside = bm >= 0 && bn >= 0 && bm < nt && bn < nt && (tri == 1 ? bm >= bn : bm <= bn);
        if (!inside || seen[(size_t)(bm * nt + bn)]) { failures++; continue; }
        seen[(size_t)(bm * nt + bn)] = 1;
        // a tile on the enumeration is never "outside"; it is "partial" exactly when it touches the diagonal
        if (tri_outside(tri, bm * 128, 128, bn * 128, 128)) failures++;
        if (tri_partial(tri, bm * 128, 128, bn * 128, 128) != (bm == bn)) failures++;
      }
      int64_t covered = 0;
      for (char c : seen) covered += c;
      if (covered != T) failures++;
    }
  // no rounding trouble for very large tile indices (sqrt in double, then integer correction)
  for (int64_t r : {(int64_t)1 << 20, ((int64_t)1 << 26) + 12345, ((int64_t)1 << 30) - 1})
    for (int64_t c : {(int64_t)0, r / 2, r}) {
      int64_t bm, bn;
      tri_tile_coords(r * (r + 1) / 2 + c, 1, bm, bn);
      if (bm != r || bn != c) failures++;
      tri_tile_coords(r * (r + 1) / 2 + c, 2, bm, bn);
      if (bm != c || bn != r) failures++;
    }
  // element mask
  if (!tri_keep(1, 5, 5) || !tri_keep(1, 6, 5) || tri_keep(1, 4, 5) || !tri_keep(2, 4, 5) || tri_keep(2, 6, 5) || !tri_keep(0, 0, 9)) failures++;
  // rectangular tiles (64 x 128, the ZGEMM kernel): a tile is skipped only if it has no element in the triangle
  for (int tri = 1; tri <= 2; tri++)
    for (int64_t bm = 0; bm < 40; bm++)
      for (int64_t bn = 0; bn < 20; bn++) {
        bool any = false, all = true;
        for (int64_t m = bm * 64; m < bm * 64 + 64; m++)
          for (int64_t n = bn * 128; n < bn * 128 + 128; n++) { const bool k = tri_keep(tri, m, n); any |= k; all &= k; }
        if (tri_outside(tri, bm * 64, 64, bn * 128, 128) != !any) failures++;
        if (any && tri_partial(tri, bm * 64, 64, bn * 128, 128) != !all) failures++;
      }
  // tiles twice as tall as wide (256 x 128): the compact enumeration visits exactly the tiles that are not "outside", once each
  for (int tri = 1; tri <= 2; tri++)
    for (int64_t n = 1; n <= 40 * 128; n += 61) {
      const int64_t tm = (n + 255) / 256, tn = (n + 127) / 128;
      std::vector<char> seen((size_t)(tm * tn), 0);
      const int64_t T = tri21_tile_count(tri, tm, tn);
      for (int64_t u = 0; u < T; u++) {
        int64_t bm, bn;
        tri21_tile_coords(u, tri, bm, bn);
        if (bm < 0 || bn < 0 || bm >= tm || bn >= tn || seen[(size_t)(bm * tn + bn)] || tri_outside(tri, bm * 256, 256, bn * 128, 128)) { failures++; continue; }
        seen[(size_t)(bm * tn + bn)] = 1;
      }
      for (int64_t bm = 0; bm < tm; bm++)
        for (int64_t bn = 0; bn < tn; bn++)
          if (!tri_outside(tri, bm * 256, 256, bn * 128, 128) && !seen[(size_t)(bm * tn + bn)]) failures++;
    }
  // tiles twice as wide as tall (64 x 128, the ZGEMM kernel): same property through the transposed enumeration
  for (int tri = 1; tri <= 2; tri++)
    for (int64_t n = 1; n <= 40 * 128; n += 53) {
      const int64_t tm = (n + 63) / 64, tn = (n + 127) / 128;
      std::vector<char> seen((size_t)(tm * tn), 0);
      const int64_t T = tri12_tile_count(tri, tm, tn);
      for (int64_t u = 0; u < T; u++) {
        int64_t bm, bn;
        tri12_tile_coords(u, tri, bm, bn);
        if (bm < 0 || bn < 0 || bm >= tm || bn >= tn || seen[(size_t)(bm * tn + bn)] || tri_outside(tri, bm * 64, 64, bn * 128, 128)) { failures++; continue; }
        seen[(size_t)(bm * tn + bn)] = 1;
      }
      for (int64_t bm = 0; bm < tm; bm++)
        for (int64_t bn = 0; bn < tn; bn++)
          if (!tri_outside(tri, bm * 64, 64, bn * 128, 128) && !seen[(size_t)(bm * tn + bn)]) failures++;
    }
  for (int64_t r : {(int64_t)1 << 20, ((int64_t)1 << 29) + 777}) {
    int64_t bm, bn;
    tri21_tile_coords(r * (r + 1) + 5, 1, bm, bn);
    if (bm != r || bn != 5) failures++;
    tri21_tile_coords((r + 1) * (r + 1) + 3, 2, bm, bn);
    if (bn != 2 * r + 1 || bm != 3) failures++;
  }
  printf(failures ? "TRI TILES: %d FAILURES\n" : "TRI TILES OK (%d failures)\n", failures);
  return failures != 0;
}
