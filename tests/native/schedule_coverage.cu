#include <cstdio>
#include <vector>
#include "cg_fused.cuh"
using namespace cg::fused;
int check(int G, int nfam, int npairs, int steal) {
  std::vector<int> cover(nfam * npairs, 0);
  int maxload = 0, minload = 1 << 30;
  for (int cta = 0; cta < G; ++cta) {
    Schedule s(cta, G, nfam, npairs, steal);
    int load = 0;
    for (int i = 0; i < s.nseg(); ++i) {
      Seg g = s.get(i);
      for (int m = 0; m < g.count; ++m) {
        int j = g.j0 + m * g.stride;
        if (j < 0 || j >= npairs || g.fam < 0 || g.fam >= nfam) { printf("OOB G=%d nfam=%d cta=%d seg=%d j=%d fam=%d\n", G, nfam, cta, i, j, g.fam); return 1; }
        cover[g.fam * npairs + j]++;
        ++load;
      }
    }
    if (load > maxload) maxload = load;
    if (load < minload) minload = load;
  }
  for (int i = 0; i < nfam * npairs; ++i) if (cover[i] != 1) { printf("COVER G=%d nfam=%d npairs=%d steal=%d fam=%d j=%d -> %d\n", G, nfam, npairs, steal, i / npairs, i % npairs, cover[i]); return 1; }
  printf("ok G=%d nfam=%d npairs=%d steal=%d load %d..%d\n", G, nfam, npairs, steal, minload, maxload);
  return 0;
}
int main() {
  int rc = 0;
  rc |= check(148, 20, 256, 0); rc |= check(148, 20, 256, 11); rc |= check(74, 10, 256, 11);
  rc |= check(148, 20, 384, 17); rc |= check(74, 10, 2048, 95); rc |= check(148, 32, 256, 5);
  rc |= check(74, 16, 256, 9); rc |= check(148, 20, 5, 3); rc |= check(30, 20, 100, 10); rc |= check(21, 20, 100, 10);
  rc |= check(8, 20, 100, 10); rc |= check(40, 20, 100, 10); rc |= check(148, 20, 1, 1);
  for (int G = 1; G <= 160; ++G) for (int nf : {1, 2, 10, 16, 20, 32}) for (int np : {1, 7, 64}) for (int st : {0, 1, 5, 16}) {
    std::vector<int> cover(nf * np, 0);
    for (int cta = 0; cta < G; ++cta) { Schedule s(cta, G, nf, np, st); for (int i = 0; i < s.nseg(); ++i) { Seg g = s.get(i); for (int m = 0; m < g.count; ++m) { int j = g.j0 + m * g.stride; if (j<0||j>=np||g.fam<0||g.fam>=nf) {printf("OOB2 %d %d %d %d\n",G,nf,np,st); return 1;} cover[g.fam * np + j]++; } } }
    for (int v : cover) if (v != 1) { printf("COVER2 G=%d nf=%d np=%d st=%d\n", G, nf, np, st); return 1; }
  }
  printf("sweep ok rc=%d\n", rc);
  return rc;
}
