// bj_symbolic.cpp -- see bj_symbolic.h.  Integer-only host code.
#include "bj_symbolic.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <numeric>

extern "C" {
// METIS 5 as shipped in the CUDA toolkit (libmetis_static.a): idx_t = int64, real_t = float.
int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options,
                 int64_t* perm, int64_t* iperm);
}

// METIS draws its random numbers from libc's rand(): one process-wide state behind a lock. Eight concurrent
// METIS_NodeND calls then run 3x slower than one alone (measured: 11.5 s against 3.6 s for a 64^3 block) and give
// a different ordering every run. These definitions are hidden, so only the METIS objects linked into this
// library bind to them: a thread-local 64-bit LCG, re-seeded by METIS at the start of every call -- concurrent
// calls neither wait for each other nor disturb each other's sequence, and an ordering depends on its graph only.
extern "C" {
static thread_local unsigned long long pcu_rand_state = 4321ULL;
__attribute__((visibility("hidden"))) void srand(unsigned int seed) { pcu_rand_state = seed; }
__attribute__((visibility("hidden"))) int rand(void) {
  pcu_rand_state = pcu_rand_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (int)((pcu_rand_state >> 33) & 0x7fffffffULL);
}
}

namespace pcu {
namespace {

// Liu's elimination-tree algorithm with path compression.  low[k] lists the i < k with a_ik != 0.
void etree(int n, const std::vector<int64_t>& lptr, const std::vector<int>& lidx, std::vector<int>& parent) {
  parent.assign(n, -1);
  std::vector<int> anc(n, -1);
  for (int k = 0; k < n; ++k)
    for (int64_t p = lptr[k]; p < lptr[k + 1]; ++p) {
      int i = lidx[p];
      while (i != -1 && i < k) {
        int nx = anc[i];
        anc[i] = k;
        if (nx == -1) parent[i] = k;
        i = nx;
      }
    }
}

// Post-order of a forest; children are visited in increasing label order.
// Children are visited smallest subtree first, except that tiny subtrees (<= kTinySubtree columns: the single
// columns eliminated early by the minimum-degree ordering of the dissection leaves) come last. A node is then
// numbered right after its tiny children and, before those, its largest child: contiguous with both, which is what
// lets the relaxed amalgamation absorb the tiny ones and still consider the top of the largest subtree.
constexpr int kTinySubtree = 4;
void postorder(int n, const std::vector<int>& parent, std::vector<int>& post) {
  std::vector<int> desc(n, 1);
  for (int j = 0; j < n; ++j) if (parent[j] != -1) desc[parent[j]] += desc[j];  // parent[j] > j in an etree
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  auto key = [&](int a) { return desc[a] <= kTinySubtree ? INT_MAX : desc[a]; };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) > key(b); });
  std::vector<int> head(n, -1), next(n, -1), stack;
  for (int j : order)  // each push goes to the front of its parent's list: the visiting order is the reverse
    if (parent[j] != -1) { next[j] = head[parent[j]]; head[parent[j]] = j; }
  post.clear();
  post.reserve(n);
  stack.reserve(n);
  for (int r = 0; r < n; ++r) {
    if (parent[r] != -1) continue;
    stack.push_back(r);
    while (!stack.empty()) {
      int p = stack.back();
      int c = head[p];
      if (c == -1) { post.push_back(p); stack.pop_back(); }
      else { head[p] = next[c]; stack.push_back(c); }
    }
  }
}

// Column counts of the Cholesky factor (diagonal included) by the Gilbert-Ng-Peyton
// skeleton/least-common-ancestor method.  up[j] lists the i > j with a_ij != 0.
void colcounts(int n, const std::vector<int64_t>& uptr, const std::vector<int>& uidx,
               const std::vector<int>& parent, const std::vector<int>& post, std::vector<int64_t>& cc) {
  std::vector<int> first(n, -1), maxfirst(n, -1), prevleaf(n, -1), anc(n);
  cc.assign(n, 0);
  std::iota(anc.begin(), anc.end(), 0);
  for (int k = 0; k < n; ++k) {
    int j = post[k];
    cc[j] = (first[j] == -1) ? 1 : 0;
    for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
  }
  for (int k = 0; k < n; ++k) {
    const int j = post[k];
    if (parent[j] != -1) cc[parent[j]]--;
    for (int64_t p = uptr[j]; p < uptr[j + 1]; ++p) {
      const int i = uidx[p];
      if (i <= j || first[j] <= maxfirst[i]) continue;  // j is not a leaf of the i-th row subtree
      maxfirst[i] = first[j];
      const int jprev = prevleaf[i];
      prevleaf[i] = j;
      cc[j]++;
      if (jprev != -1) {
        int q = jprev;
        while (q != anc[q]) q = anc[q];
        for (int s = jprev; s != q;) { int sp = anc[s]; anc[s] = q; s = sp; }
        cc[q]--;
      }
    }
    if (parent[j] != -1) anc[j] = parent[j];
  }
  for (int k = 0; k < n; ++k) {  // accumulate up the tree (post-order guarantees children first)
    const int j = post[k];
    if (parent[j] != -1) cc[parent[j]] += cc[j];
  }
}

inline int64_t tri(int64_t w) { return w * (w + 1) / 2; }

}  // namespace

int analyze(int n, const int* rowPtr, const int* colInd, const SymbolicOptions& opt, Symbolic* S) {
  *S = Symbolic();
  S->n = n;
  if (n <= 0) return -1;
  // ---- 1. symmetric adjacency without the diagonal
  std::vector<int64_t> xadj(n + 1, 0);
  for (int i = 0; i < n; ++i)
    for (int p = rowPtr[i]; p < rowPtr[i + 1]; ++p) {
      const int j = colInd[p];
      if (j < 0 || j >= n) return -2;
      if (j != i) { xadj[i + 1]++; xadj[j + 1]++; }
    }
  for (int i = 0; i < n; ++i) xadj[i + 1] += xadj[i];
  std::vector<int64_t> adj(xadj[n]);
  {
    std::vector<int64_t> pos(xadj.begin(), xadj.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int p = rowPtr[i]; p < rowPtr[i + 1]; ++p) {
        const int j = colInd[p];
        if (j != i) { adj[pos[i]++] = j; adj[pos[j]++] = i; }
      }
  }
  // ---- 2. fill-reducing ordering
  std::vector<int> mperm(n), miperm(n);
  if (opt.use_metis && n >= 8 && xadj[n] > 0) {
    std::vector<int64_t> p64(n), ip64(n);
    int64_t nv = n;
    const auto tm0 = std::chrono::steady_clock::now();
    int rc = METIS_NodeND(&nv, xadj.data(), adj.data(), nullptr, nullptr, p64.data(), ip64.data());
    if (rc != 1) return -3;
    S->ordering_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tm0).count();
    if (getenv("PREALPS_BJ_PROFILE")) fprintf(stderr, "  METIS_NodeND(n=%d): %.2f s\n", n, S->ordering_seconds);
    for (int i = 0; i < n; ++i) { mperm[i] = (int)p64[i]; miperm[i] = (int)ip64[i]; }
  } else {
    std::iota(mperm.begin(), mperm.end(), 0);
    std::iota(miperm.begin(), miperm.end(), 0);
  }
  // ---- 3. pattern of P A P^T: up[lo] = {hi > lo}, low[hi] = {lo < hi}
  auto build_lists = [&](const std::vector<int>& ip, std::vector<int64_t>& uptr, std::vector<int>& uidx,
                         std::vector<int64_t>& lptr, std::vector<int>& lidx) {
    uptr.assign(n + 1, 0);
    lptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i)
      for (int p = rowPtr[i]; p < rowPtr[i + 1]; ++p) {
        const int j = colInd[p];
        if (j == i) continue;
        const int a = ip[i], b = ip[j];
        uptr[std::min(a, b) + 1]++;
        lptr[std::max(a, b) + 1]++;
      }
    for (int i = 0; i < n; ++i) { uptr[i + 1] += uptr[i]; lptr[i + 1] += lptr[i]; }
    uidx.resize(uptr[n]);
    lidx.resize(lptr[n]);
    std::vector<int64_t> up(uptr.begin(), uptr.end() - 1), lp(lptr.begin(), lptr.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int p = rowPtr[i]; p < rowPtr[i + 1]; ++p) {
        const int j = colInd[p];
        if (j == i) continue;
        const int a = ip[i], b = ip[j];
        uidx[up[std::min(a, b)]++] = std::max(a, b);
        lidx[lp[std::max(a, b)]++] = std::min(a, b);
      }
  };
  std::vector<int64_t> uptr, lptr;
  std::vector<int> uidx, lidx, parent, post;
  build_lists(miperm, uptr, uidx, lptr, lidx);
  // ---- 4/5. elimination tree and its post-order, composed into the final permutation
  etree(n, lptr, lidx, parent);
  postorder(n, parent, post);
  if ((int)post.size() != n) return -4;
  S->perm.resize(n);
  S->iperm.resize(n);
  for (int k = 0; k < n; ++k) S->perm[k] = mperm[post[k]];
  for (int k = 0; k < n; ++k) S->iperm[S->perm[k]] = k;
  build_lists(S->iperm, uptr, uidx, lptr, lidx);
  etree(n, lptr, lidx, parent);  // now post-ordered: parent[k] > k and subtrees are contiguous
  std::vector<int> ident(n);
  std::iota(ident.begin(), ident.end(), 0);
  std::vector<int64_t> cc;
  colcounts(n, uptr, uidx, parent, ident, cc);
  for (int k = 0; k < n; ++k) S->nnzL_exact += cc[k];

  // ---- 6. fundamental supernodes: k+1 joins k when struct(k) = {k} U struct(k+1)
  struct SN { int first, last; int64_t h; int64_t zeros; };
  std::vector<SN> fund;
  for (int k = 0; k < n; ++k) {
    if (!fund.empty() && parent[k - 1] == k && cc[k - 1] == cc[k] + 1) { fund.back().last = k; }
    else fund.push_back({k, k, 0, 0});
  }
  for (auto& s : fund) s.h = cc[s.first];  // rows of the trapezoid = count of its first column
  // ---- 7a. whole-subtree merge: an elimination subtree with few columns becomes one dense supernode
  // desc[k] = number of columns in the subtree rooted at k (post-ordered => [k-desc+1, k])
  std::vector<int> desc(n, 1);
  for (int k = 0; k < n; ++k) if (parent[k] != -1) desc[parent[k]] += desc[k];
  std::vector<SN> st1;
  {
    // walk fundamental supernodes; a supernode whose LAST column has a subtree <= leaf_cols and whose
    // parent's subtree is larger absorbs its entire subtree
    size_t i = 0;
    std::vector<SN> tmp;
    for (const SN& s : fund) {
      const int k = s.last;
      const int par = parent[k];
      const bool small = desc[k] <= opt.leaf_cols;
      const bool top = small && (par == -1 || desc[par] > opt.leaf_cols);
      if (top && desc[k] > (s.last - s.first + 1)) {
        const int first = k - desc[k] + 1;
        while (!tmp.empty() && tmp.back().first >= first) tmp.pop_back();
        const int64_t w = desc[k];
        const int64_t h = w + (cc[k] - 1);
        // explicit zeros = dense trapezoid - exact entries of those columns
        int64_t exact = 0;
        for (int c = first; c <= k; ++c) exact += cc[c];
        tmp.push_back({first, k, h, tri(w) + (h - w) * w - exact});
      } else {
        tmp.push_back(s);
      }
      (void)i;
    }
    st1.swap(tmp);
  }
  // ---- 7b. relaxed last-child amalgamation (bottom-up, keeps column ranges contiguous)
  std::vector<SN> fin;
  for (const SN& s0 : st1) {
    SN p = s0;
    while (!fin.empty()) {
      const SN& c = fin.back();
      // c is the supernode numbered right before p; it can join p when its parent column lies inside p
      if (parent[c.last] < p.first || parent[c.last] > p.last) break;
      const int64_t wc = c.last - c.first + 1, wp = p.last - p.first + 1;
      const int64_t hc = c.h, hp = p.h;
      const int64_t wn = wc + wp, hn = wc + hp;
      const int64_t stor = tri(wn) + (hn - wn) * wn;
      // explicit zeros of the merged trapezoid = its storage - the exact entries of both parts
      const int64_t exact = (tri(wc) + (hc - wc) * wc - c.zeros) + (tri(wp) + (hp - wp) * wp - p.zeros);
      const int64_t zt = stor - exact;
      bool merge;
      if (wn <= opt.relax_small) merge = true;
      else if (wn <= 4 * opt.relax_small) merge = (double)zt < 1.5 * opt.relax_zero * (double)stor;
      else if (wn <= opt.relax_big_cols) merge = (double)zt < opt.relax_zero * (double)stor;
      else merge = (double)zt < opt.relax_big * (double)stor;
      if (!merge) break;
      p.first = c.first;
      p.h = hn;
      p.zeros = zt;
      fin.pop_back();
    }
    fin.push_back(p);
  }
  const int ns = (int)fin.size();
  S->nsuper = ns;
  S->sn_col.resize(ns + 1);
  S->col2sn.resize(n);
  for (int s = 0; s < ns; ++s) {
    S->sn_col[s] = fin[s].first;
    for (int c = fin[s].first; c <= fin[s].last; ++c) S->col2sn[c] = s;
  }
  S->sn_col[ns] = n;
  // ---- 8. supernodal tree, row structures (children before parents), levels
  S->sn_parent.assign(ns, -1);
  for (int s = 0; s < ns; ++s) {
    const int pc = parent[fin[s].last];
    S->sn_parent[s] = pc == -1 ? -1 : S->col2sn[pc];
  }
  std::vector<int> chead(ns, -1), cnext(ns, -1);
  for (int s = ns - 1; s >= 0; --s)
    if (S->sn_parent[s] != -1) { cnext[s] = chead[S->sn_parent[s]]; chead[S->sn_parent[s]] = s; }
  S->sn_rowptr.assign(ns + 1, 0);
  S->sn_rows.clear();
  std::vector<int> mark(n, -1), buf;
  S->sn_level.assign(ns, 0);
  for (int s = 0; s < ns; ++s) {
    const int a = S->sn_col[s], b = S->sn_col[s + 1];
    buf.clear();
    for (int c = a; c < b; ++c)
      for (int64_t p = uptr[c]; p < uptr[c + 1]; ++p) {
        const int r = uidx[p];
        if (r >= b && mark[r] != s) { mark[r] = s; buf.push_back(r); }
      }
    for (int c = chead[s]; c != -1; c = cnext[c]) {
      const int wc = S->sn_col[c + 1] - S->sn_col[c];
      for (int64_t p = S->sn_rowptr[c] + wc; p < S->sn_rowptr[c + 1]; ++p) {
        const int r = S->sn_rows[p];
        if (r >= b && mark[r] != s) { mark[r] = s; buf.push_back(r); }
      }
      S->sn_level[s] = std::max(S->sn_level[s], S->sn_level[c] + 1);
    }
    std::sort(buf.begin(), buf.end());
    for (int c = a; c < b; ++c) S->sn_rows.push_back(c);
    S->sn_rows.insert(S->sn_rows.end(), buf.begin(), buf.end());
    S->sn_rowptr[s + 1] = (int64_t)S->sn_rows.size();
    const int64_t w = b - a, h = S->sn_rowptr[s + 1] - S->sn_rowptr[s];
    S->nnzL_stored += tri(w) + (h - w) * w;
    for (int64_t c = 0; c < w; ++c) { const double cnt = (double)(h - c); S->flops += cnt * cnt; }
    S->nlevels = std::max(S->nlevels, S->sn_level[s] + 1);
  }
  return 0;
}

}  // namespace pcu
