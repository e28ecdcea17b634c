// Contraction-tree refinement on the host (planning; no GPU work): subtree reconfiguration and per-leg slicing scores.
//
// The reference obtains its better-than-greedy paths from cotengra through rustengra
// (tnc/src/contractionpath/paths/tree_reconfiguration.rs:54-58 `cotengra_optimized_greedy(.., subtree_size)`,
// hyperoptimization.rs:69-76); neither is part of /root/reference.  This file restates the published technique
// (Gray & Kourtis, "Hyper-optimized tensor network contraction", 2021, section "subtree reconfiguration"): pick a
// connected piece of the contraction tree with at most `subtree_size` frontier nodes, find the optimal contraction
// order of that frontier by dynamic programming over subsets, splice it back when it is cheaper, sweep until no
// piece improves.  Legs follow the reference's tensor model: a leg joins exactly two tensors and disappears when they
// meet (Tensor::symmetric_difference, tensor.rs:463-479), so the legs of any set of tensors is the XOR of its members.
//
// cost(node) = 2^w(legs(l) | legs(r)) + size_weight * 2^w(legs(l) ^ legs(r))       (w = sum of log2 dims)
#include "internal.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

namespace tncb {
namespace {

typedef unsigned __int128 u128;

// Objective of one pair step from the log2 sizes of its operands (w1, w2) and of their shared legs (wk):
//   flops mode:  2^(w1 + w2 - wk) + size_weight * 2^(w1 + w2 - 2 wk)
//   time mode :  the device-time model of contractionpath/contraction_cost.py gpu_time_mnk (same constants, passed in)
struct Objective {
  bool time = false;
  double size_weight = 0.0;
  double crt = 160e12, k_half = 600.0, dmma = 34e12, hbm = 5e12, launch = 5e-6, dmma_k_half = 24.0, crt_k_max = 1048576.0, conv = 40.0;
  double pair(double w1, double w2, double wk) const {
    if (!time) return std::exp2(w1 + w2 - wk) + (size_weight != 0.0 ? size_weight * std::exp2(w1 + w2 - 2.0 * wk) : 0.0);
    // contraction_cost.gpu_time_mnk, term by term
    const double m = std::exp2(w1 - wk), n = std::exp2(w2 - wk), k = std::exp2(wk);
    const double mnk = m * n * k, flops = 8.0 * mnk;
    const double t_mem = 16.0 * (m * k + n * k + m * n) / hbm;
    double t = std::max(flops / (dmma * k / (k + dmma_k_half)), t_mem);
    if (m >= 128.0 && n >= 128.0 && k >= 256.0 && k <= crt_k_max && mnk >= 268435456.0) {
      const double t_crt = flops / (crt * k / (k + k_half)) + conv * (m * k + n * k) / hbm;
      t = std::min(t, std::max(t_crt, t_mem));
    }
    return t + launch;
  }
  void load(double sw, const double* tm) {
    size_weight = sw; time = tm != nullptr;
    if (tm) { crt = tm[0]; k_half = tm[1]; dmma = tm[2]; hbm = tm[3]; launch = tm[4]; dmma_k_half = tm[5]; crt_k_max = tm[6]; conv = tm[7]; }
  }
};

struct Tree {
  int n = 0, W = 0;                       // leaves, words per leg set
  std::vector<uint64_t> legs;             // (2n-1) * W
  std::vector<int> left, right, parent;   // -1 for leaves / the root
  std::vector<double> lw;                 // log2 dim per leg id
  int root = -1;
  uint64_t* L(int v) { return legs.data() + (size_t)v * W; }
  const uint64_t* L(int v) const { return legs.data() + (size_t)v * W; }
  double weight(const uint64_t* a) const {
    double s = 0;
    for (int w = 0; w < W; w++) { uint64_t x = a[w]; while (x) { s += lw[w * 64 + __builtin_ctzll(x)]; x &= x - 1; } }
    return s;
  }
  double union_weight(int a, int b) const {
    double s = 0;
    for (int w = 0; w < W; w++) { uint64_t x = L(a)[w] | L(b)[w]; while (x) { s += lw[w * 64 + __builtin_ctzll(x)]; x &= x - 1; } }
    return s;
  }
  double node_flops(int v) const { return std::exp2(union_weight(left[v], right[v])); }
  double shared_weight(int a, int b) const {
    double s = 0;
    for (int w = 0; w < W; w++) { uint64_t x = L(a)[w] & L(b)[w]; while (x) { s += lw[w * 64 + __builtin_ctzll(x)]; x &= x - 1; } }
    return s;
  }
  double node_cost(int v, const Objective& ob) const { return ob.pair(weight(L(left[v])), weight(L(right[v])), shared_weight(left[v], right[v])); }
  double node_size(int v) const { return std::exp2(weight(L(v))); }
};

// ssa pairs -> tree; returns false on a malformed path
bool build_tree(Tree& T, int n, int W, const uint64_t* leaf_legs, const double* leg_log2, const int32_t* ssa) {
  T.n = n; T.W = W;
  const int N = 2 * n - 1;
  T.legs.assign((size_t)N * W, 0); T.left.assign(N, -1); T.right.assign(N, -1); T.parent.assign(N, -1);
  T.lw.assign(leg_log2, leg_log2 + (size_t)W * 64);
  std::memcpy(T.legs.data(), leaf_legs, (size_t)n * W * sizeof(uint64_t));
  for (int t = 0; t < n - 1; t++) {
    const int a = ssa[2 * t], b = ssa[2 * t + 1], v = n + t;
    if (a < 0 || b < 0 || a >= v || b >= v || a == b || T.parent[a] != -1 || T.parent[b] != -1) return false;
    T.left[v] = a; T.right[v] = b; T.parent[a] = v; T.parent[b] = v;
    for (int w = 0; w < W; w++) T.L(v)[w] = T.L(a)[w] ^ T.L(b)[w];
  }
  T.root = N - 1;
  return true;
}

struct Work {   // scratch of one subtree optimisation
  std::vector<u128> L;        // legs of every frontier subset (super-leg masks)
  std::vector<double> cost, wL;   // best cost of the subset, log2 size of its legs
  std::vector<uint32_t> split;
  double tab[16][256];
};

inline double mask_weight(const Work& wk, u128 m) {
  double s = 0;
  for (int b = 0; b < 16; b++) { s += wk.tab[b][(unsigned)(m & 0xff)]; m >>= 8; if (!m) break; }
  return s;
}

// One subtree rooted at r: returns the improvement (old - new cost, > 0 when the tree was changed).
double reconfigure_at(Tree& T, int r, int subtree_size, const Objective& ob, std::mt19937_64& rng, int select, Work& wk) {
  if (T.left[r] < 0) return 0.0;
  // ---- grow the piece: expand the costliest (select 0) or a random (select 1) internal frontier node ----
  std::vector<int> frontier = {T.left[r], T.right[r]}, inner = {r};
  while ((int)frontier.size() < subtree_size) {
    int pick = -1; double best = -1.0; int seen = 0;
    for (int i = 0; i < (int)frontier.size(); i++) {
      const int v = frontier[i];
      if (T.left[v] < 0) continue;
      if (select == 0) { const double c = T.node_cost(v, ob); if (c > best) { best = c; pick = i; } }
      else { seen++; if ((rng() % seen) == 0) pick = i; }
    }
    if (pick < 0) break;
    const int v = frontier[pick];
    inner.push_back(v);
    frontier[pick] = T.left[v]; frontier.push_back(T.right[v]);
  }
  const int n = (int)frontier.size();
  if (n < 3) return 0.0;
  double old_cost = 0.0;
  for (int v : inner) old_cost += T.node_cost(v, ob);
  // ---- super legs: legs with the same membership over the frontier are one weighted leg ----
  // a leg sits in one or two frontier nodes (sym-diff model); signature = (i, j) with j = i for a single owner
  std::vector<int> sig_id((size_t)n * n, -1);
  std::vector<double> sw; std::vector<u128> fl(n, 0);
  {
    std::vector<int> owner((size_t)T.W * 64, -1);
    for (int i = 0; i < n; i++) {
      const uint64_t* a = T.L(frontier[i]);
      for (int w = 0; w < T.W; w++) {
        uint64_t x = a[w];
        while (x) {
          const int leg = w * 64 + __builtin_ctzll(x); x &= x - 1;
          if (owner[leg] < 0) owner[leg] = i;
          else if (owner[leg] < n) owner[leg] = owner[leg] * n + i + n * n;   // second owner: encode the pair
          else return 0.0;                                                    // not a two-owner leg: leave this piece alone
        }
      }
    }
    for (int leg = 0; leg < T.W * 64; leg++) {
      const int o = owner[leg];
      if (o < 0) continue;
      int i, j;
      if (o >= n * n) { const int p = o - n * n; i = p / n; j = p % n; } else { i = j = o; }
      int& id = sig_id[(size_t)i * n + j];
      if (id < 0) { id = (int)sw.size(); sw.push_back(0.0); }
      sw[id] += T.lw[leg];
      fl[i] |= (u128)1 << id; if (j != i) fl[j] |= (u128)1 << id;
    }
    if (sw.size() > 128) return 0.0;
  }
  for (int b = 0; b < 16; b++)
    for (int x = 0; x < 256; x++) {
      double s = 0;
      for (int k = 0; k < 8; k++) if ((x >> k) & 1) { const int id = b * 8 + k; if (id < (int)sw.size()) s += sw[id]; }
      wk.tab[b][x] = s;
    }
  // ---- optimal order of the frontier: DP over subsets ----
  const uint32_t full = (1u << n) - 1;
  wk.L.assign((size_t)full + 1, 0); wk.cost.assign((size_t)full + 1, 0.0); wk.wL.assign((size_t)full + 1, 0.0); wk.split.assign((size_t)full + 1, 0);
  for (uint32_t S = 1; S <= full; S++) {
    const uint32_t low = S & (~S + 1), rest = S ^ low;
    if (!rest) { wk.L[S] = fl[__builtin_ctz(S)]; wk.wL[S] = mask_weight(wk, wk.L[S]); continue; }
    wk.L[S] = wk.L[rest] ^ wk.L[low];
    wk.wL[S] = mask_weight(wk, wk.L[S]);
    double best = INFINITY; uint32_t bs = 0;
    for (uint32_t sub = (rest - 1) & rest;; sub = (sub - 1) & rest) {     // S1 = low | sub, S2 = rest \ sub (non-empty)
      const uint32_t S1 = low | sub, S2 = S ^ S1;
      const double base = wk.cost[S1] + wk.cost[S2];
      if (base < best) {
        const double c = base + ob.pair(wk.wL[S1], wk.wL[S2], mask_weight(wk, wk.L[S1] & wk.L[S2]));
        if (c < best) { best = c; bs = S1; }
      }
      if (!sub) break;
    }
    wk.cost[S] = best; wk.split[S] = bs;
  }
  const double new_cost = wk.cost[full];
  if (!(new_cost < old_cost * (1.0 - 1e-9))) return 0.0;
  // ---- splice: the inner node ids are reused (r stays the root of the piece) ----
  std::vector<int> ids(inner.begin() + 1, inner.end());
  struct Item { uint32_t S; int id; };
  std::vector<Item> stack = {{full, r}};
  std::vector<std::pair<int, uint32_t>> order;   // (node, subset) parents before children
  while (!stack.empty()) {
    const Item it = stack.back(); stack.pop_back();
    const uint32_t S1 = wk.split[it.S], S2 = it.S ^ S1;
    auto child = [&](uint32_t S) {
      if ((S & (S - 1)) == 0) return frontier[__builtin_ctz(S)];
      const int id = ids.back(); ids.pop_back();
      stack.push_back({S, id});
      return id;
    };
    const int a = child(S1), b = child(S2);
    T.left[it.id] = a; T.right[it.id] = b; T.parent[a] = it.id; T.parent[b] = it.id;
    order.push_back({it.id, it.S});
  }
  for (int i = (int)order.size() - 1; i >= 0; i--) {     // legs bottom-up
    const int v = order[i].first;
    for (int w = 0; w < T.W; w++) T.L(v)[w] = T.L(T.left[v])[w] ^ T.L(T.right[v])[w];
  }
  return old_cost - new_cost;
}

void totals(const Tree& T, const Objective& ob, double* flops, double* max_size, double* objective) {
  double f = 0, m = 0, o = 0;
  for (int v = T.n; v < 2 * T.n - 1; v++) {
    const double c = T.node_flops(v), s = T.node_size(v);
    f += c; m = std::max(m, s); o += T.node_cost(v, ob);
  }
  for (int v = 0; v < T.n; v++) m = std::max(m, T.node_size(v));
  if (flops) *flops = f;
  if (max_size) *max_size = m;
  if (objective) *objective = o;
}

void emit_ssa(const Tree& T, int32_t* ssa) {
  // post-order over the internal nodes; children before parents, new ssa ids in that order
  std::vector<int> newid(2 * T.n - 1, -1);
  for (int v = 0; v < T.n; v++) newid[v] = v;
  std::vector<std::pair<int, int>> st = {{T.root, 0}};
  int next = T.n, t = 0;
  while (!st.empty()) {
    auto& top = st.back();
    const int v = top.first;
    if (T.left[v] < 0) { st.pop_back(); continue; }
    if (top.second == 0) { top.second = 1; st.push_back({T.left[v], 0}); }
    else if (top.second == 1) { top.second = 2; st.push_back({T.right[v], 0}); }
    else {
      const int a = newid[T.left[v]], b = newid[T.right[v]];       // (smaller id first, as cotengra writes its ssa paths)
      ssa[2 * t] = std::min(a, b); ssa[2 * t + 1] = std::max(a, b);
      newid[v] = next++; t++;
      st.pop_back();
    }
  }
}

}  // namespace
}  // namespace tncb

using namespace tncb;

extern "C" {

int tncb_path_reconfigure(int n_leaves, int n_words, const uint64_t* leaf_legs, const double* leg_log2, int32_t* ssa_pairs,
                          int subtree_size, int max_sweeps, double size_weight, const double* time_model, uint64_t seed,
                          double* flops, double* max_size, double* objective) {
  if (n_leaves < 1 || n_words < 1 || !leaf_legs || !leg_log2 || (n_leaves > 1 && !ssa_pairs)) return fail(TNCB_ERR_INVALID, "null / empty argument");
  if (subtree_size < 2 || subtree_size > 15) return fail(TNCB_ERR_INVALID, "subtree_size must be in [2, 15]");
  Tree T;
  if (!build_tree(T, n_leaves, n_words, leaf_legs, leg_log2, ssa_pairs)) return fail(TNCB_ERR_INVALID, "ssa path is not a binary tree over the leaves");
  std::mt19937_64 rng(seed);
  Work wk;
  Objective ob; ob.load(size_weight, time_model);
  double obj; totals(T, ob, nullptr, nullptr, &obj);
  for (int sweep = 0; sweep < max_sweeps && n_leaves > 2; sweep++) {
    // costliest pieces first; odd sweeps grow the pieces at random for variety
    std::vector<int> order(n_leaves - 1);
    std::iota(order.begin(), order.end(), n_leaves);
    std::vector<double> key(2 * n_leaves - 1, 0.0);
    for (int v : order) key[v] = T.node_cost(v, ob) * (1.0 + 1e-3 * (double)(rng() % 1000));
    std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] > key[b]; });
    double gained = 0.0;
    for (int v : order) gained += reconfigure_at(T, v, subtree_size, ob, rng, sweep & 1, wk);
    const double before = obj;
    totals(T, ob, nullptr, nullptr, &obj);
    if (!(gained > 0.0) || before - obj < 1e-4 * before) { if (sweep & 1) break; }
  }
  if (n_leaves > 1) emit_ssa(T, ssa_pairs);
  totals(T, ob, flops, max_size, objective);
  return TNCB_OK;
}

// Per-leg slicing scores of a tree: cost_without[l] = the objective of the tree (flops, or model seconds with time_model)
// once leg l is fixed (ONE slice), size_without[l] = its largest tensor then.
int tncb_path_leg_scores(int n_leaves, int n_words, const uint64_t* leaf_legs, const double* leg_log2, const int32_t* ssa_pairs,
                         double size_weight, const double* time_model,
                         double* cost_without, double* size_without, double* cost, double* max_size) {
  if (n_leaves < 1 || n_words < 1 || !leaf_legs || !leg_log2 || (n_leaves > 1 && !ssa_pairs) || !cost_without || !size_without)
    return fail(TNCB_ERR_INVALID, "null / empty argument");
  Tree T;
  if (!build_tree(T, n_leaves, n_words, leaf_legs, leg_log2, ssa_pairs)) return fail(TNCB_ERR_INVALID, "ssa path is not a binary tree over the leaves");
  Objective ob; ob.load(size_weight, time_model);
  const int NL = n_words * 64;
  std::vector<double> sizes(2 * n_leaves - 1), big_with(NL, 0.0), big_without(NL, 0.0), delta(NL, 0.0);
  for (int v = 0; v < 2 * n_leaves - 1; v++) sizes[v] = T.node_size(v);
  for (int v = 0; v < 2 * n_leaves - 1; v++)
    for (int l = 0; l < NL; l++) {
      const bool has = (T.L(v)[l >> 6] >> (l & 63)) & 1;
      if (has) big_with[l] = std::max(big_with[l], sizes[v]); else big_without[l] = std::max(big_without[l], sizes[v]);
    }
  double total = 0.0;
  for (int v = n_leaves; v < 2 * n_leaves - 1; v++) {
    const int a = T.left[v], b = T.right[v];
    const double w1 = T.weight(T.L(a)), w2 = T.weight(T.L(b)), wk = T.shared_weight(a, b);
    const double c = ob.pair(w1, w2, wk);
    total += c;
    for (int w = 0; w < n_words; w++) {
      uint64_t x = T.L(a)[w] | T.L(b)[w];
      while (x) {
        const int bit = __builtin_ctzll(x); x &= x - 1;
        const int l = w * 64 + bit;
        const bool ina = (T.L(a)[w] >> bit) & 1, inb = (T.L(b)[w] >> bit) & 1;
        const double d = leg_log2[l];
        delta[l] += c - ob.pair(w1 - (ina ? d : 0.0), w2 - (inb ? d : 0.0), wk - (ina && inb ? d : 0.0));
      }
    }
  }
  for (int l = 0; l < NL; l++) {
    cost_without[l] = total - delta[l];
    size_without[l] = std::max(big_without[l], big_with[l] / std::exp2(leg_log2[l]));
  }
  double f, m;
  totals(T, ob, &f, &m, nullptr);
  if (cost) *cost = total;
  if (max_size) *max_size = m;
  return TNCB_OK;
}

}  // extern "C"
