// Fill-reducing orderings for the quasidefinite KKT matrix (host side, one-time).
//
// Replaces the reference's call into the third-party `amd` crate
// (/root/reference/src/qdldl/qdldl.rs:905-917, dense scale 1.5 from
// ldlsolvers/qdldl.rs:41).  Written from the published algorithm
// (Amestoy, Davis, Duff, "An approximate minimum degree ordering algorithm",
// SIMAX 1996): quotient graph, approximate external degrees, element
// absorption, mass elimination and supervariable detection by hashing.
//
// On top of it: a nested-dissection driver (George's automatic ND with BFS
// level structures) used to get a *short, bushy* elimination tree, which is
// what the level-scheduled device factorisation and solves want.  The
// reference has no equivalent; any valid permutation yields a valid LDL^T and
// the parity tests run the oracle on the very same permutation.
#include "symbolic.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>
#include <unordered_map>
#include <atomic>
#include <future>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace cb {
static double onow() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define OMARK(label) do { if (std::getenv("CB_TIMING")) { double t_ = onow(); std::fprintf(stderr, "[cb timing]   nd: %-22s %.3f s\n", label, t_ - ot_last); ot_last = t_; } } while (0)


namespace {

// Build the full symmetric adjacency (no diagonal, no duplicates) of a triu CSC.
void full_adjacency(int n, const int64_t* Ap, const int32_t* Ai,
                    std::vector<int64_t>& xadj, std::vector<int>& adj) {
  std::vector<int64_t> cnt(n + 1, 0);
  for (int j = 0; j < n; j++)
    for (int64_t p = Ap[j]; p < Ap[j + 1]; p++) {
      int i = Ai[p];
      if (i != j) { cnt[i + 1]++; cnt[j + 1]++; }
    }
  xadj.assign(n + 1, 0);
  for (int i = 0; i < n; i++) xadj[i + 1] = xadj[i] + cnt[i + 1];
  adj.resize(xadj[n]);
  std::vector<int64_t> pos(xadj.begin(), xadj.end() - 1);
  for (int j = 0; j < n; j++)
    for (int64_t p = Ap[j]; p < Ap[j + 1]; p++) {
      int i = Ai[p];
      if (i != j) { adj[pos[i]++] = j; adj[pos[j]++] = i; }
    }
  // sort + unique per row
  std::vector<int64_t> nx(n + 1, 0);
  int64_t w = 0;
  for (int i = 0; i < n; i++) {
    int64_t b = xadj[i], e = xadj[i + 1];
    std::sort(adj.begin() + b, adj.begin() + e);
    nx[i] = w;
    int last = -1;
    for (int64_t p = b; p < e; p++)
      if (adj[p] != last) { adj[w++] = adj[p]; last = adj[p]; }
  }
  nx[n] = w;
  adj.resize(w);
  xadj.swap(nx);
}

enum : uint8_t { ST_VAR = 0, ST_ELEM = 1, ST_DEAD_ELEM = 2, ST_ABSORBED = 3, ST_DENSE = 4 };

}  // namespace

// Approximate minimum degree on a general undirected graph given as CSR
// adjacency (xadj/adj).  `order` receives the elimination sequence (perm:
// order[k] = vertex eliminated k-th).
// `halo` (optional): vertices that take part in the degree computations and in the elements but are never
// eliminated and never emitted -- the boundary of a region whose interior is being ordered (halo-AMD: the separator
// vertices a nested-dissection leaf touches will be eliminated later, so a leaf vertex next to many of them is more
// expensive than its degree inside the leaf says).
void amd_graph(int n, const std::vector<int64_t>& xadj, const std::vector<int>& adjncy,
               double dense_scale, std::vector<int>& order, const std::vector<char>* forced_first,
               const std::vector<char>* halo, const std::atomic<bool>* cancel) {
  order.clear();
  order.reserve(n);
  if (n == 0) return;

  std::vector<std::vector<int>> adj(n), elems(n), Le(n);
  std::vector<int> nv(n, 1), degree(n), esize(n, 0);
  std::vector<uint8_t> status(n, ST_VAR);
  std::vector<int> sv_next(n, -1), sv_tail(n);  // absorbed-variable chains
  std::vector<int64_t> w(n, 0);
  int64_t wflg = 1;
  for (int i = 0; i < n; i++) sv_tail[i] = i;

  // "dense" rows are withheld and ordered last (AMD's dense-row rule:
  // threshold max(16, 10*scale*sqrt(n)))
  double dense = 10.0 * dense_scale * std::sqrt((double)n);
  if (dense < 16.0) dense = 16.0;
  std::vector<int> dense_nodes;
  for (int i = 0; i < n; i++) {
    int d = (int)(xadj[i + 1] - xadj[i]);
    if ((double)d > dense && !(forced_first && (*forced_first)[i]) && !(halo && (*halo)[i])) { status[i] = ST_DENSE; dense_nodes.push_back(i); }
  }
  for (int i = 0; i < n; i++) {
    if (status[i] == ST_DENSE) continue;
    auto& a = adj[i];
    a.reserve(xadj[i + 1] - xadj[i]);
    for (int64_t p = xadj[i]; p < xadj[i + 1]; p++)
      if (status[adjncy[p]] != ST_DENSE) a.push_back(adjncy[p]);
    degree[i] = (forced_first && (*forced_first)[i]) ? 0 : (int)a.size();   // forced vertices are eliminated first
  }

  // degree buckets
  std::vector<int> head(n + 1, -1), nxt(n, -1), prv(n, -1);
  auto bucket_insert = [&](int i, int d) {
    nxt[i] = head[d]; prv[i] = -1;
    if (head[d] >= 0) prv[head[d]] = i;
    head[d] = i;
  };
  auto bucket_remove = [&](int i, int d) {
    if (prv[i] >= 0) nxt[prv[i]] = nxt[i]; else head[d] = nxt[i];
    if (nxt[i] >= 0) prv[nxt[i]] = prv[i];
  };
  auto is_halo = [&](int i) { return halo && (*halo)[i]; };
  int nlive = 0, nhalo = 0;
  for (int i = n - 1; i >= 0; i--)
    if (status[i] == ST_VAR) { if (is_halo(i)) nhalo++; else { bucket_insert(i, degree[i]); nlive++; } }

  int nel = 0, mindeg = 0;
  const int ntot = nlive;
  std::vector<int> Lp, bucket_of_hash(n, -1), hnext(n, -1);
  std::vector<unsigned> hashv(n, 0);
  std::vector<int> touched_hash;
  std::vector<int> tmp;

  while (nel < ntot) {
    if (cancel && cancel->load(std::memory_order_relaxed)) { order.clear(); return; }   // the caller no longer wants it
    while (mindeg <= n && head[mindeg] < 0) mindeg++;
    int p = head[mindeg];
    bucket_remove(p, mindeg);
    int nvpiv = nv[p];
    nel += nvpiv;
    nv[p] = -nvpiv;

    // ---- form the new element Lp ----
    Lp.clear();
    int degme = 0;
    auto take = [&](int i) {
      if (status[i] == ST_VAR && nv[i] > 0) {
        degme += nv[i];
        nv[i] = -nv[i];
        Lp.push_back(i);
        if (!is_halo(i)) bucket_remove(i, degree[i]);
      }
    };
    for (int i : adj[p]) take(i);
    for (int e : elems[p])
      if (status[e] == ST_ELEM) {
        for (int i : Le[e]) take(i);
        status[e] = ST_DEAD_ELEM;
        std::vector<int>().swap(Le[e]);
      }
    status[p] = ST_ELEM;
    std::vector<int>().swap(adj[p]);
    std::vector<int>().swap(elems[p]);

    // ---- |Le \ Lp| for every element touching Lp ----
    if (wflg > (int64_t)4e18) { std::fill(w.begin(), w.end(), 0); wflg = 1; }
    for (int i : Lp) {
      int nvi = -nv[i];
      for (int e : elems[i]) {
        if (status[e] != ST_ELEM) continue;
        if (w[e] >= wflg) w[e] -= nvi;
        else w[e] = (int64_t)esize[e] + wflg - nvi;
      }
    }

    // ---- degree update, pruning, hashing ----
    touched_hash.clear();
    for (int i : Lp) {
      int nvi = -nv[i];
      int64_t deg = 0;
      unsigned h = 0;
      auto& ei = elems[i];
      size_t k = 0;
      for (size_t t = 0; t < ei.size(); t++) {
        int e = ei[t];
        if (status[e] != ST_ELEM) continue;
        int64_t dext = w[e] - wflg;
        if (dext > 0) { deg += dext; ei[k++] = e; h += (unsigned)e; }
        else { status[e] = ST_DEAD_ELEM; std::vector<int>().swap(Le[e]); }  // aggressive absorption
      }
      ei.resize(k);
      auto& ai = adj[i];
      k = 0;
      for (size_t t = 0; t < ai.size(); t++) {
        int j = ai[t];
        if (status[j] == ST_VAR && nv[j] > 0) { deg += nv[j]; ai[k++] = j; h += (unsigned)j; }
      }
      ai.resize(k);
      if (ei.empty() && ai.empty() && !is_halo(i)) {
        // mass elimination: i has no neighbours outside Lp
        degme -= nvi; nvpiv += nvi; nel += nvi;
        nv[i] = 0; status[i] = ST_ABSORBED;
        sv_next[sv_tail[p]] = i; sv_tail[p] = sv_tail[i];
      } else {
        degree[i] = (int)std::min<int64_t>(degree[i], deg);
        ei.push_back(p);
        h += (unsigned)p;
        h %= (unsigned)n;
        hashv[i] = h;
        if (!is_halo(i)) {     // halo vertices are never merged into supervariables
          if (bucket_of_hash[h] < 0) touched_hash.push_back((int)h);
          hnext[i] = bucket_of_hash[h];
          bucket_of_hash[h] = i;
        }
      }
    }
    wflg += (int64_t)n + 1;  // invalidate all w[e]

    // ---- supervariable detection within hash buckets ----
    for (int h : touched_hash) {
      int i = bucket_of_hash[h];
      bucket_of_hash[h] = -1;
      for (; i >= 0; i = hnext[i]) {
        if (nv[i] == 0) continue;
        // tag i's lists
        wflg++;
        for (int e : elems[i]) w[e] = wflg;
        for (int j : adj[i]) w[j] = wflg;
        int prev = i;
        for (int j = hnext[i]; j >= 0; j = hnext[j]) {
          bool same = nv[j] != 0 && elems[j].size() == elems[i].size() &&
                      adj[j].size() == adj[i].size();
          if (same) for (int e : elems[j]) if (w[e] != wflg) { same = false; break; }
          if (same) for (int v : adj[j]) if (w[v] != wflg) { same = false; break; }
          if (same) {
            nv[i] += nv[j];  // both negative here
            nv[j] = 0; status[j] = ST_ABSORBED;
            std::vector<int>().swap(adj[j]); std::vector<int>().swap(elems[j]);
            sv_next[sv_tail[i]] = j; sv_tail[i] = sv_tail[j];
            hnext[prev] = hnext[j];
          } else prev = j;
        }
      }
    }
    wflg += (int64_t)n + 1;

    // ---- finalise element p, restore degree lists ----
    auto& Lpe = Le[p];
    Lpe.clear();
    for (int i : Lp) {
      int nvi = -nv[i];
      if (nvi <= 0) continue;
      nv[i] = nvi;
      int64_t deg = (int64_t)degree[i] + degme - nvi;
      deg = std::min<int64_t>(deg, (int64_t)ntot + nhalo - nel - nvi);
      if (deg < 0) deg = 0;
      degree[i] = (int)deg;
      if (!is_halo(i)) {
        bucket_insert(i, degree[i]);
        if (degree[i] < mindeg) mindeg = degree[i];
      }
      Lpe.push_back(i);
    }
    nv[p] = nvpiv;
    esize[p] = degme;
    if (degme == 0) { status[p] = ST_DEAD_ELEM; std::vector<int>().swap(Le[p]); }

    // emit everything absorbed into p (mass-eliminated and indistinguishable variables), then p itself: the convention
    // of the AMD the reference calls, whose output lists the non-principal variables of a pivot before the pivot
    // (src/qdldl/test.rs:124-129 pins [3, 0, 1, 2] on its 4 x 4 matrix).  Same fill either way: these variables have no
    // neighbours outside the pivot's clique.
    for (int v = sv_next[p]; v >= 0; v = sv_next[v]) order.push_back(v);
    order.push_back(p);
  }
  // dense rows last, lowest degree first
  std::sort(dense_nodes.begin(), dense_nodes.end(), [&](int a, int b) {
    int64_t da = xadj[a + 1] - xadj[a], db = xadj[b + 1] - xadj[b];
    return da != db ? da < db : a < b;
  });
  for (int v : dense_nodes) order.push_back(v);
}

void amd_order(int n, const int64_t* Ap, const int32_t* Ai, double dense_scale,
               std::vector<int>& perm, const std::atomic<bool>* cancel) {
  std::vector<int64_t> xadj;
  std::vector<int> adj;
  full_adjacency(n, Ap, Ai, xadj, adj);
  amd_graph(n, xadj, adj, dense_scale, perm, nullptr, nullptr, cancel);
}

// ---------------------------------------------------------------------------
// Nested dissection (BFS level-structure bisection) with AMD on the leaves.
// Recursive; the two halves of the first few bisections are ordered by separate
// host threads (disjoint vertex sets, so the shared per-vertex scratch arrays
// are written without conflicts).
// ---------------------------------------------------------------------------
namespace {

struct NDShared {
  const std::vector<int64_t>* xadj;
  const std::vector<int>* adj;
  std::vector<int> part;   // region id per vertex (-1 = separator / withheld)
  std::vector<int> level;  // BFS scratch
  std::vector<int> local;  // global -> local index scratch for the leaf AMD
  std::atomic<int> next_region{1};
  int leaf_size = 200;
  int par_depth = 0;
  bool halo = false;       // halo-AMD on the leaves (CB_ND_HALO)
  bool hubs = true;        // hub separators for small-world regions (CB_ND_HUBS=0 turns them off)
  // a separator of k vertices is a dense k x k block at the top of its subtree: k^3 / 3 flops.  Beyond this cap the whole
  // dissection is given up (the caller falls back to minimum degree); 0 = never
  double sep_flop_cap = 0.0;
  std::atomic<bool> hopeless{false};
};

// BFS restricted to vertices with part[v]==region; returns eccentricity, fills queue and level[].
int nd_bfs(NDShared& W, int root, int region, const std::vector<int>& verts, std::vector<int>& q) {
  for (int v : verts) W.level[v] = -1;
  q.clear();
  q.push_back(root);
  W.level[root] = 0;
  size_t headp = 0;
  int maxl = 0;
  while (headp < q.size()) {
    const int v = q[headp++];
    const int lv = W.level[v];
    for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1]; p++) {
      const int u = (*W.adj)[p];
      if (W.part[u] == region && W.level[u] < 0) {
        W.level[u] = lv + 1;
        if (lv + 1 > maxl) maxl = lv + 1;
        q.push_back(u);
      }
    }
  }
  return maxl;
}

void nd_leaf(NDShared& W, const std::vector<int>& verts, std::vector<int>& out) {
  const int m = (int)verts.size();
  if (m == 0) return;
  if (m <= 2) { for (int v : verts) out.push_back(v); return; }
  for (int k = 0; k < m; k++) W.local[verts[k]] = k;
  if (W.halo) {
    // halo-AMD: the separator vertices this region touches join the graph as vertices that are never eliminated.
    // They get ids m, m+1, ... through a map private to this call (another thread may see the same separator vertex
    // as the halo of ITS region at the same time, so the shared `local` array must not be used for them).
    std::unordered_map<int, int> hid;
    std::vector<std::vector<int>> hadj;          // halo vertex -> leaf neighbours
    std::vector<int64_t> sx(m + 1, 0);
    std::vector<int> sa, lorder;
    for (int k = 0; k < m; k++) {
      const int v = verts[k];
      for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1]; p++) {
        const int u = (*W.adj)[p];
        if (W.local[u] >= 0) { sa.push_back(W.local[u]); continue; }
        if (W.part[u] != -1) continue;           // only separators (dense rows withheld by the driver also carry -1)
        auto it = hid.find(u);
        int h;
        if (it == hid.end()) { h = (int)hadj.size(); hid.emplace(u, h); hadj.emplace_back(); } else h = it->second;
        hadj[h].push_back(k);
        sa.push_back(m + h);
      }
      sx[k + 1] = (int64_t)sa.size();
    }
    const int nh = (int)hadj.size();
    sx.resize(m + nh + 1);
    for (int h = 0; h < nh; h++) { for (int k : hadj[h]) sa.push_back(k); sx[m + h + 1] = (int64_t)sa.size(); }
    std::vector<char> halo(m + nh, 0);
    for (int h = 0; h < nh; h++) halo[m + h] = 1;
    amd_graph(m + nh, sx, sa, 1e9, lorder, nullptr, &halo);
    for (int k : lorder) if (k < m) out.push_back(verts[k]);
    for (int k = 0; k < m; k++) W.local[verts[k]] = -1;
    return;
  }
  std::vector<int64_t> sx(m + 1, 0);
  std::vector<int> sa, lorder;
  for (int k = 0; k < m; k++) {
    const int v = verts[k];
    for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1]; p++) {
      const int u = (*W.adj)[p];
      if (W.local[u] >= 0) sa.push_back(W.local[u]);
    }
    sx[k + 1] = (int64_t)sa.size();
  }
  // neighbours outside `verts` have local == -1 only if they were never in a concurrently processed leaf:
  // leaves handled by different threads are vertex-disjoint AND separated, so no edge joins them.
  amd_graph(m, sx, sa, 1e9, lorder);
  for (int k : lorder) out.push_back(verts[k]);
  for (int k = 0; k < m; k++) W.local[verts[k]] = -1;
}


// Hub separator for small-world regions.  A BFS level structure has no thin level when a few vertices (linking rows of
// a block-angular problem, coupling constraints) join parts of the graph that are otherwise far apart.  Such vertices
// have above-typical degree, so: withhold the top fraction of the region by degree, label the components of the rest,
// and then give back every withheld vertex that touches at most one sizeable component (it joins that component; tiny
// components it touches are merged in).  What cannot be given back is a vertex separator: the real connectors.  The
// degree only proposes candidates; connectivity decides.  Returns true with `sep` filled when the remainder falls into
// components none of which holds more than 70 % of the region and the separator is below 5 % of it.
bool hub_separator(NDShared& W, const std::vector<int>& verts, int region, std::vector<int>& sep) {
  const int total = (int)verts.size();
  std::vector<int> deg(total);
  for (int k = 0; k < total; k++) W.local[verts[k]] = k;
  int dmax = 0;
  for (int k = 0; k < total; k++) {
    const int v = verts[k];
    int d = 0;
    for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1]; p++) if (W.part[(*W.adj)[p]] == region) d++;
    deg[k] = d;
    dmax = std::max(dmax, d);
  }
  // degree quantiles from a histogram (degrees are small integers)
  std::vector<int64_t> hist(dmax + 2, 0);
  for (int k = 0; k < total; k++) hist[deg[k]]++;
  auto quantile = [&](double f) {          // smallest degree d with  #{deg > d} <= f * total
    int64_t above = 0;
    int d = dmax;
    while (d > 0 && above + hist[d] <= (int64_t)(f * total)) { above += hist[d]; d--; }
    return d + 1;
  };
  int median = 0;
  { int64_t below = 0; while (median < dmax && below + hist[median] <= total / 2) { below += hist[median]; median++; } }
  const int tiny = std::max(64, W.leaf_size / 8);
  std::vector<int> thrs;
  const double fracs[] = {0.0005, 0.002, 0.005, 0.02, 0.05, 0.1, 0.2};
  for (double f : fracs) {
    int thr = std::max(quantile(f), median + 1);
    if (thr > dmax) continue;
    int64_t ns0 = 0;
    for (int d = thr; d <= dmax; d++) ns0 += hist[d];
    if (ns0 == 0 || ns0 * 10 > (int64_t)total * 3) continue;
    if (thrs.empty() || thrs.back() != thr) thrs.push_back(thr);
  }
  // Second proposal rule, for connectors of unremarkable degree (a linking row that touches 4 blocks has FEWER
  // neighbours than an ordinary row): locality.  In a graph with locality two neighbours of a vertex are usually
  // adjacent or share another neighbour (two variables of one constraint row appear together in other rows of the
  // same window); the neighbours of a connector lie in parts that only it joins.  Candidates = vertices of degree
  // >= 2 none of whose neighbour pairs (among the first 8 neighbours) is adjacent or has a common neighbour besides
  // the vertex itself.  Ordinary vertices leave the test at the first close pair, so the pass is cheap; whatever it
  // proposes wrongly is given back by the connectivity step.
  std::vector<char> cellcand;
  auto propose_by_locality = [&]() {
    std::vector<int> mark(total, -1);
    cellcand.assign(total, 0);
    int64_t nc = 0;
    int nb[8];
    for (int k = 0; k < total; k++) {
      const int v = verts[k];
      int cnt = 0;
      for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1] && cnt < 8; p++) {
        const int u = (*W.adj)[p];
        if (W.part[u] == region) nb[cnt++] = u;
      }
      if (cnt < 2) continue;
      bool close = false;
      for (int i = 0; i + 1 < cnt && !close; i++) {
        const int ui = nb[i];
        const int stamp = k * 8 + i;
        for (int64_t p = (*W.xadj)[ui]; p < (*W.xadj)[ui + 1]; p++) {
          const int w = (*W.adj)[p];
          if (W.part[w] == region) mark[W.local[w]] = stamp;
        }
        for (int j = i + 1; j < cnt && !close; j++) {
          const int uj = nb[j];
          if (mark[W.local[uj]] == stamp) { close = true; break; }
          for (int64_t p = (*W.xadj)[uj]; p < (*W.xadj)[uj + 1]; p++) {
            const int w = (*W.adj)[p];
            if (w != v && W.part[w] == region && mark[W.local[w]] == stamp) { close = true; break; }
          }
        }
      }
      if (!close) { cellcand[k] = 1; nc++; }
    }
    return nc;
  };
  // one attempt: withhold the candidates (deg >= thr, or the locality rule when thr < 0), label the components of the rest,
  // give back what is not a connector
  auto attempt = [&](int thr, std::vector<int>& out) -> bool {
    std::vector<int> comp(total), cpar, csize, stack;
    auto find = [&](int c) { while (cpar[c] != c) { cpar[c] = cpar[cpar[c]]; c = cpar[c]; } return c; };
    if (thr >= 0) for (int k = 0; k < total; k++) comp[k] = deg[k] >= thr ? -2 : -1;
    else for (int k = 0; k < total; k++) comp[k] = cellcand[k] ? -2 : -1;
    for (int k0 = 0; k0 < total; k0++) {
      if (comp[k0] != -1) continue;
      const int c = (int)cpar.size();
      cpar.push_back(c); csize.push_back(0);
      stack.clear(); stack.push_back(k0); comp[k0] = c;
      while (!stack.empty()) {
        const int k = stack.back(); stack.pop_back();
        csize[c]++;
        const int v = verts[k];
        for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1]; p++) {
          const int u = (*W.adj)[p];
          if (W.part[u] != region) continue;
          const int ku = W.local[u];
          if (comp[ku] == -1) { comp[ku] = c; stack.push_back(ku); }
        }
      }
      if ((int64_t)csize[c] * 10 > (int64_t)total * 6) {                // the rest still hangs together: withhold more
        if (thr < 0 && std::getenv("CB_ND_DEBUG")) std::fprintf(stderr, "[nd debug] locality rule: component of %d of %d after withholding\n", csize[c], total);
        return false;
      }
    }
    for (int pass = 0; pass < 8; pass++) {
      int64_t moved = 0;
      for (int k = 0; k < total; k++) {
        if (comp[k] != -2) continue;
        const int v = verts[k];
        int big = -1, small = -1;
        bool two = false;
        for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1] && !two; p++) {
          const int u = (*W.adj)[p];
          if (W.part[u] != region) continue;
          const int cu = comp[W.local[u]];
          if (cu < 0) continue;
          const int r = find(cu);
          if (csize[r] > tiny) { if (big >= 0 && big != r) two = true; big = r; }
          else small = r;
        }
        if (two) continue;
        const int target = big >= 0 ? big : small;
        if (target < 0) continue;                               // all neighbours withheld: decided in a later pass
        for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1]; p++) {
          const int u = (*W.adj)[p];
          if (W.part[u] != region) continue;
          const int cu = comp[W.local[u]];
          if (cu < 0) continue;
          const int r = find(cu);
          if (r != target) { cpar[r] = target; csize[target] += csize[r]; }
        }
        comp[k] = target; csize[target]++;
        moved++;
      }
      if (!moved) break;
    }
    int64_t nsep = 0, largest = 0;
    for (int k = 0; k < total; k++) if (comp[k] == -2) nsep++;
    for (size_t c = 0; c < cpar.size(); c++) if (cpar[c] == (int)c) largest = std::max<int64_t>(largest, csize[c]);
    if (thr < 0 && std::getenv("CB_ND_DEBUG")) std::fprintf(stderr, "[nd debug] locality rule: after give-back largest %lld, separator %lld of %d\n", (long long)largest, (long long)nsep, total);
    if (largest * 10 > (int64_t)total * 7 || nsep * 20 > (int64_t)total || nsep == 0) return false;
    out.clear();
    for (int k = 0; k < total; k++) if (comp[k] == -2) out.push_back(verts[k]);
    return true;
  };
  // the attempts are independent and read-only on the shared arrays: on a large region they run side by side on the
  // host threads (this is the top of the recursion, the other cores are idle), and the FIRST threshold of the list that
  // succeeds is taken -- the same answer as trying them one after the other
  bool ok = false;
  if (total >= 200000 && host_threads() > 1 && thrs.size() > 1) {
    std::vector<std::vector<int>> outs(thrs.size());
    std::vector<std::future<bool>> futs;
    for (size_t t = 0; t < thrs.size(); t++)
      futs.push_back(std::async(std::launch::async, [&, t]() { return attempt(thrs[t], outs[t]); }));
    std::vector<char> good(thrs.size(), 0);
    for (size_t t = 0; t < thrs.size(); t++) good[t] = futs[t].get() ? 1 : 0;
    for (size_t t = 0; t < thrs.size() && !ok; t++) if (good[t]) { sep.swap(outs[t]); ok = true; }
  } else {
    for (size_t t = 0; t < thrs.size() && !ok; t++) ok = attempt(thrs[t], sep);
  }
  if (!ok) {     // no degree threshold isolates the connectors: propose by locality
    const int64_t nc = propose_by_locality();
    if (nc > 0 && nc * 10 <= (int64_t)total * 3) ok = attempt(-1, sep);
    if (std::getenv("CB_TIMING")) std::fprintf(stderr, "[cb timing]   nd: hub separator, locality rule: %lld candidates of %d -> %s (%zu connectors)\n", (long long)nc, total, ok ? "ok" : "no", sep.size());
  }
  for (int v : verts) W.local[v] = -1;
  return ok;
}

void nd_rec(NDShared& W, std::vector<int>& verts, int depth, std::vector<int>& out) {
  if (W.hopeless.load(std::memory_order_relaxed)) return;
  const double t_enter = onow();
  const size_t total = verts.size();
  if ((int)total <= W.leaf_size) { nd_leaf(W, verts, out); return; }
  const int region = W.next_region.fetch_add(1);
  for (int v : verts) W.part[v] = region;
  std::vector<int> q;
  q.reserve(total);
  // one sweep to find a far vertex, then the level structure is rooted there
  int root = verts[0];
  int e = nd_bfs(W, root, region, verts, q);
  if (q.size() == total && W.hubs && (int)total >= 50 * W.leaf_size) {
    // small-world signature already in the first sweep (the two levels around the balance point hold more than 4 % of
    // the region each): no pseudo-peripheral root will make the levels thin -- go for the connectors at once and save
    // the second sweep (0.12 s on C4)
    std::vector<int64_t> c1(e + 1, 0);
    for (int v : q) c1[W.level[v]]++;
    int64_t acc = 0;
    int cut1 = 0;
    while (cut1 < e - 1 && acc + c1[cut1] < (int64_t)total / 2) acc += c1[cut1++];
    if (e >= 1 && c1[cut1] * 25 > (int64_t)total && c1[std::min(cut1 + 1, e)] * 25 > (int64_t)total) {
      std::vector<int> hsep;
      if (hub_separator(W, q, region, hsep) &&
          !(W.sep_flop_cap > 0.0 && (double)hsep.size() * (double)hsep.size() * (double)hsep.size() / 3.0 > W.sep_flop_cap)) {
        if (std::getenv("CB_TIMING")) std::fprintf(stderr, "[cb timing]   nd: depth %d hub separator of %zu: %zu connectors (first sweep: levels around the balance point hold %lld and %lld vertices), %.4f s\n", depth, total, hsep.size(), (long long)c1[cut1], (long long)c1[std::min(cut1 + 1, e)], onow() - t_enter);
        for (int v : hsep) W.part[v] = -1;
        std::vector<int> rest;
        rest.reserve(total - hsep.size());
        for (int v : q) if (W.part[v] == region) rest.push_back(v);
        std::vector<int>().swap(verts);
        std::vector<int>().swap(q);
        nd_rec(W, rest, depth, out);
        nd_leaf(W, hsep, out);
        return;
      }
    }
  }
  if (q.size() == total) {
    int best = q.back();
    int64_t bestdeg = INT64_MAX;
    for (size_t t = q.size(); t-- > 0;) {
      const int v = q[t];
      if (W.level[v] != e) break;
      const int64_t d = (*W.xadj)[v + 1] - (*W.xadj)[v];
      if (d < bestdeg) { bestdeg = d; best = v; }
    }
    root = best;
    e = nd_bfs(W, root, region, verts, q);
  }
  if (q.size() < total) {
    // the region is disconnected: label all components in one pass; big ones recurse, the small ones are
    // packed together into leaves (independent pieces cost AMD nothing extra)
    for (int v : verts) W.level[v] = -1;
    std::vector<int> pack, comp;
    std::vector<std::vector<int>> bigs;
    for (int sv : verts) {
      if (W.level[sv] >= 0) continue;
      comp.clear(); comp.push_back(sv); W.level[sv] = 0;
      for (size_t h = 0; h < comp.size(); h++) {
        const int v = comp[h];
        for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1]; p++) {
          const int u = (*W.adj)[p];
          if (W.part[u] == region && W.level[u] < 0) { W.level[u] = 0; comp.push_back(u); }
        }
      }
      if ((int)comp.size() > W.leaf_size) {
        bigs.emplace_back(comp);
      } else {
        pack.insert(pack.end(), comp.begin(), comp.end());
        if ((int)pack.size() >= 4 * W.leaf_size) { nd_leaf(W, pack, out); pack.clear(); }
      }
    }
    if (!pack.empty()) nd_leaf(W, pack, out);
    // the sizeable components are independent subproblems: ordered concurrently while the thread budget lasts
    // (k components count as log2(k) levels of the recursion's thread tree)
    int dd = depth + 1;
    for (size_t k = 1; k < bigs.size(); k <<= 1) dd++;
    if (bigs.size() > 1 && depth < W.par_depth) {
      std::vector<std::vector<int>> outs(bigs.size());
      std::vector<std::future<void>> futs;
      for (size_t b = 1; b < bigs.size(); b++)
        futs.push_back(std::async(std::launch::async, [&, b]() { nd_rec(W, bigs[b], dd, outs[b]); }));
      nd_rec(W, bigs[0], dd, outs[0]);
      for (auto& f : futs) f.get();
      for (auto& o : outs) out.insert(out.end(), o.begin(), o.end());
    } else {
      for (auto& b : bigs) nd_rec(W, b, depth + 1, out);
    }
    return;
  }
  if (e < 2) { nd_leaf(W, verts, out); return; }   // clique-like, cannot bisect
  // Candidate cuts: the levels around the one that balances the vertex counts.  For a cut between levels c and
  // c+1 every crossing edge joins the two boundary sets A (level c) and B (level c+1); the smallest vertex
  // separator inside A u B is a minimum vertex cover of that bipartite graph (Koenig: from a maximum matching,
  // Hopcroft-Karp).  It is never larger than the one-sided choice "all of A" and often much smaller.
  std::vector<int64_t> cntl(e + 1, 0);
  for (int v : q) cntl[W.level[v]]++;
  std::vector<int64_t> pre(e + 2, 0);
  for (int l = 0; l <= e; l++) pre[l + 1] = pre[l] + cntl[l];
  const int64_t half = (int64_t)total / 2;
  int cut = 0;
  while (cut < e - 1 && pre[cut + 1] < half) cut++;      // levels 0..cut hold at least half (or cut = e-1)
  // small-world signature: the levels around the balance point are fat (more than 4 % of the region each), so every
  // candidate cut below would be fat too -- look for connector vertices first and save the matchings on huge boundary
  // sets (C4: 1.1 s of Hopcroft-Karp on 3.3e5 boundary vertices for a cut that is then thrown away)
  if (W.hubs && (int)total >= 50 * W.leaf_size && cntl[cut] * 25 > (int64_t)total && cntl[std::min(cut + 1, e)] * 25 > (int64_t)total) {
    std::vector<int> hsep;
    if (hub_separator(W, q, region, hsep) &&
        !(W.sep_flop_cap > 0.0 && (double)hsep.size() * (double)hsep.size() * (double)hsep.size() / 3.0 > W.sep_flop_cap)) {
      if (std::getenv("CB_TIMING")) std::fprintf(stderr, "[cb timing]   nd: depth %d hub separator of %zu: %zu connectors (levels around the balance point hold %lld and %lld vertices), %.4f s\n", depth, total, hsep.size(), (long long)cntl[cut], (long long)cntl[std::min(cut + 1, e)], onow() - t_enter);
      for (int v : hsep) W.part[v] = -1;
      std::vector<int> rest;
      rest.reserve(total - hsep.size());
      for (int v : q) if (W.part[v] == region) rest.push_back(v);
      std::vector<int>().swap(verts);
      std::vector<int>().swap(q);
      nd_rec(W, rest, depth, out);
      nd_leaf(W, hsep, out);
      return;
    }
  }
  std::vector<int> sep, bestsep;
  int bestc = -1;
  double bestscore = 1e300;
  std::vector<int> A, B, matchA, matchB, dist, stk;
  std::vector<std::vector<int>> nbr;
  int cwin = 2;
  if (const char* ev = std::getenv("CB_ND_CUTWIN")) cwin = std::atoi(ev);
  for (int c = std::max(0, cut - cwin); c <= std::min(e - 1, cut + cwin); c++) {
    const int64_t below = pre[c + 1], above = (int64_t)total - below;
    if (std::min(below, above) * 5 < (int64_t)total && c != cut) continue;    // keep candidates roughly balanced
    A.clear(); B.clear();
    // local ids through W.local (reset below); boundary vertices only
    for (int v : q) {
      const int lv = W.level[v];
      if (lv != c && lv != c + 1) continue;
      bool touches = false;
      for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1] && !touches; p++) {
        const int u = (*W.adj)[p];
        if (W.part[u] == region && W.level[u] == (lv == c ? c + 1 : c)) touches = true;
      }
      if (!touches) continue;
      if (lv == c) { W.local[v] = (int)A.size(); A.push_back(v); }
      else { W.local[v] = (int)B.size(); B.push_back(v); }
    }
    const int na = (int)A.size(), nb = (int)B.size();
    nbr.assign(na, std::vector<int>());
    for (int i = 0; i < na; i++) {
      const int v = A[i];
      for (int64_t p = (*W.xadj)[v]; p < (*W.xadj)[v + 1]; p++) {
        const int u = (*W.adj)[p];
        if (W.part[u] == region && W.level[u] == c + 1) nbr[i].push_back(W.local[u]);
      }
    }
    for (int v : A) W.local[v] = -1;
    for (int v : B) W.local[v] = -1;
    // Hopcroft-Karp
    matchA.assign(na, -1); matchB.assign(nb, -1); dist.assign(na, 0);
    for (;;) {
      std::vector<int> bq;
      for (int i = 0; i < na; i++) { if (matchA[i] < 0) { dist[i] = 0; bq.push_back(i); } else dist[i] = -1; }
      bool found = false;
      for (size_t h = 0; h < bq.size(); h++) {
        const int i = bq[h];
        for (int j : nbr[i]) {
          const int i2 = matchB[j];
          if (i2 < 0) found = true;
          else if (dist[i2] < 0) { dist[i2] = dist[i] + 1; bq.push_back(i2); }
        }
      }
      if (!found) break;
      // iterative DFS along the layering
      std::vector<int> it(na, 0);
      for (int r = 0; r < na; r++) {
        if (matchA[r] >= 0) continue;
        stk.clear(); stk.push_back(r);
        while (!stk.empty()) {
          const int i = stk.back();
          if (it[i] >= (int)nbr[i].size()) { dist[i] = -2; stk.pop_back(); continue; }
          const int j = nbr[i][it[i]++];
          const int i2 = matchB[j];
          if (i2 < 0) {
            // augment along the stack
            int jj = j;
            for (size_t t = stk.size(); t-- > 0;) { const int ii = stk[t]; const int pj = matchA[ii]; matchA[ii] = jj; matchB[jj] = ii; jj = pj; }
            break;
          }
          if (dist[i2] == dist[i] + 1) stk.push_back(i2);
        }
      }
    }
    // Koenig: Z = reachable from unmatched A by alternating paths; cover = (A \ Z) u (B n Z)
    std::vector<char> za(na, 0), zb(nb, 0);
    {
      std::vector<int> bq;
      for (int i = 0; i < na; i++) if (matchA[i] < 0) { za[i] = 1; bq.push_back(i); }
      for (size_t h = 0; h < bq.size(); h++) {
        const int i = bq[h];
        for (int j : nbr[i]) {
          if (zb[j] || matchA[i] == j) continue;
          zb[j] = 1;
          const int i2 = matchB[j];
          if (i2 >= 0 && !za[i2]) { za[i2] = 1; bq.push_back(i2); }
        }
      }
    }
    sep.clear();
    for (int i = 0; i < na; i++) if (!za[i]) sep.push_back(A[i]);
    for (int j = 0; j < nb; j++) if (zb[j]) sep.push_back(B[j]);
    int64_t sl = 0;                                        // separator vertices taken from the lower side
    for (int i = 0; i < na; i++) if (!za[i]) sl++;
    const int64_t nl = below - sl, nrr = above - ((int64_t)sep.size() - sl);
    const double imb = (double)std::llabs(nl - nrr) / (double)total;
    const double score = (double)sep.size() * (1.0 + 2.0 * imb * imb) + 1e-9 * std::abs(c - cut);
    if (score < bestscore) { bestscore = score; bestc = c; bestsep = sep; }
  }
  if (bestc < 0) { nd_leaf(W, verts, out); return; }
  sep.swap(bestsep);
  for (int v : sep) W.local[v] = -2;                       // mark
  std::vector<int> L, R;
  for (int v : q) {
    if (W.local[v] == -2) continue;
    if (W.level[v] <= bestc) L.push_back(v); else R.push_back(v);
  }
  for (int v : sep) W.local[v] = -1;
  // a level-structure cut is only worth keeping when it is thin and roughly balanced; small-world graphs
  // (hub rows) give neither, and the region is then left to AMD as a whole
  const size_t smaller = std::min(L.size(), R.size());
  const bool poor_cut = sep.size() * 5 > total || smaller * 20 < total;
  const bool beyond_cap = W.sep_flop_cap > 0.0 && (double)sep.size() * (double)sep.size() * (double)sep.size() / 3.0 > W.sep_flop_cap;
  // a separator above 2 % of a large region is fat: look for connectors before accepting it
  if ((poor_cut || beyond_cap || sep.size() * 50 > total) && (int)total >= 50 * W.leaf_size && W.hubs) {
    std::vector<int> hsep;
    if (hub_separator(W, q, region, hsep) && (poor_cut || beyond_cap || hsep.size() * 2 < sep.size()) &&
        !(W.sep_flop_cap > 0.0 && (double)hsep.size() * (double)hsep.size() * (double)hsep.size() / 3.0 > W.sep_flop_cap)) {
      if (std::getenv("CB_TIMING")) std::fprintf(stderr, "[cb timing]   nd: depth %d hub separator of %zu: %zu connectors (level cut had %zu), %.4f s\n", depth, total, hsep.size(), sep.size(), onow() - t_enter);
      for (int v : hsep) W.part[v] = -1;
      std::vector<int> rest;
      rest.reserve(total - hsep.size());
      for (int v : q) if (W.part[v] == region) rest.push_back(v);
      std::vector<int>().swap(verts);
      std::vector<int>().swap(q);
      nd_rec(W, rest, depth, out);                           // falls into its components there
      nd_leaf(W, hsep, out);
      return;
    }
  }
  if (poor_cut) { nd_leaf(W, verts, out); return; }
  if (beyond_cap) {
    W.hopeless.store(true);
    return;
  }
  if (depth <= 2 && std::getenv("CB_TIMING")) std::fprintf(stderr, "[cb timing]   nd: depth %d bisection of %zu: %.4f s (sep %zu)\n", depth, total, onow() - t_enter, sep.size());
  for (int v : sep) W.part[v] = -1;
  std::vector<int>().swap(verts);
  std::vector<int>().swap(q);
  if (depth < W.par_depth) {
    std::vector<int> outL;
    auto fut = std::async(std::launch::async, [&]() { nd_rec(W, L, depth + 1, outL); });
    std::vector<int> outR;
    nd_rec(W, R, depth + 1, outR);
    fut.get();
    out.insert(out.end(), outL.begin(), outL.end());
    out.insert(out.end(), outR.begin(), outR.end());
  } else {
    nd_rec(W, L, depth + 1, out);
    nd_rec(W, R, depth + 1, out);
  }
  nd_leaf(W, sep, out);
}

}  // namespace

void nd_order(int n, const int64_t* Ap, const int32_t* Ai, double dense_scale,
              int leaf_size, std::vector<int>& perm, double sep_flop_cap) {
  double ot_last = onow();
  std::vector<int64_t> xadj;
  std::vector<int> adj;
  full_adjacency(n, Ap, Ai, xadj, adj);
  OMARK("adjacency");
  perm.clear();
  perm.reserve(n);
  if (n == 0) return;

  // withhold dense rows exactly like AMD does; they go last
  double dense = 10.0 * dense_scale * std::sqrt((double)n);
  if (dense < 16.0) dense = 16.0;
  NDShared W;
  W.xadj = &xadj; W.adj = &adj;
  W.part.assign(n, 0);
  W.level.assign(n, -1);
  W.local.assign(n, -1);
  W.leaf_size = leaf_size;
  W.sep_flop_cap = sep_flop_cap;
  if (const char* e = std::getenv("CB_ND_HUBS")) W.hubs = std::atoi(e) != 0;
  W.halo = std::getenv("CB_ND_HALO") != nullptr && std::atoi(std::getenv("CB_ND_HALO")) != 0;
  {
    unsigned hc = host_threads();
    int d = 0;
    // concurrent subtrees: a few per hardware thread (leaf sizes vary, so oversubscription balances the load;
    // measured on 8 cores: 0.21 s with 8 subtrees, 0.165 s with 32), at most 128
    while ((1u << (d + 1)) <= 4 * std::max(1u, hc) && d < (hc >= 64 ? 8 : 7)) d++;      // up to 256 concurrent subtrees on a 64+-thread host
    W.par_depth = (n >= 20000) ? d : 0;
    if (const char* e = std::getenv("CB_ND_THREADS_DEPTH")) W.par_depth = std::atoi(e);
  }
  std::vector<int> dense_nodes;
  for (int i = 0; i < n; i++)
    if ((double)(xadj[i + 1] - xadj[i]) > dense) { W.part[i] = -1; dense_nodes.push_back(i); }
  // connected components of the rest are independent roots
  {
    std::vector<char> seen(n, 0);
    for (int s = 0; s < n; s++) {
      if (seen[s] || W.part[s] != 0) continue;
      std::vector<int> comp{s};
      seen[s] = 1;
      for (size_t h = 0; h < comp.size(); h++) {
        const int v = comp[h];
        for (int64_t p = xadj[v]; p < xadj[v + 1]; p++) {
          const int u = adj[p];
          if (!seen[u] && W.part[u] == 0) { seen[u] = 1; comp.push_back(u); }
        }
      }
      nd_rec(W, comp, 0, perm);
      if (W.hopeless.load()) { perm.clear(); OMARK("dissection given up (separator beyond the flop cap)"); return; }
    }
  }
  OMARK("dissection + leaf AMD");
  std::sort(dense_nodes.begin(), dense_nodes.end(), [&](int a, int b) {
    int64_t da = xadj[a + 1] - xadj[a], db = xadj[b + 1] - xadj[b];
    return da != db ? da < db : a < b;
  });
  for (int v : dense_nodes) perm.push_back(v);
}

}  // namespace cb
