#define PCU_EMUL 1
#include "bj.h"
#include <cstdio>
int main() {
  for (int w = 1; w <= 700; ++w) {
    const int W4 = (w + 3) / 4;
    long long sum = 0;
    for (int p = 0; p <= 60; ++p) {
      if (pcu::panel_cum(w, p) != sum) { std::printf("mismatch w=%d p=%d: %lld vs %lld\n", w, p, pcu::panel_cum(w, p), sum); return 1; }
      const int nkb = W4 < 8 * p + 8 ? W4 : 8 * p + 8;
      sum += 128ll * nkb;
    }
  }
  std::printf("panel_cum ok\n");
  return 0;
}
