// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// A from-scratch C++17 restatement of the *algorithm* of coitrees 0.4.0, the
// third-party crate that sequila-native's default `Coitrees` arm delegates to
// (Cargo.lock pins coitrees 0.4.0; a source-identical copy is vendored in the
// reference under sequila/sequila-core/superintervals/test/3rd-party/coitrees/src/,
// abbreviated CT/ below).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use this file.
//
// What is restated (reference file:line):
//   * sort key (first,last)                         CT/nosimd.rs:882, CT/avx.rs:466
//   * implicit balanced BST, subtree_last = max end CT/nosimd.rs:707-801
//   * expected-hit heuristic (f32, layout only)     CT/nosimd.rs:793-798
//   * van-Emde-Boas relayout + "simple subtree"
//     runs (cut-off 64 scalar / 8 chunks AVX,
//     density 0.2)                                  CT/nosimd.rs:24-29, 869-1123; CT/avx.rs:18-23
//   * query descent and pruning rules               CT/nosimd.rs:343-384; CT/avx.rs:679-733
//   * closed-interval predicate                     CT/nosimd.rs:647-649
//   * 8-interval chunk nodes, padding (MAX,MIN),
//     stored first-1 / last+1, cmpgt compare        CT/avx.rs:50-106, 307-318, 403-433
//
// LANES = 1 restates BasicCOITree (nosimd.rs); LANES = 8 restates AVXCOITree
// (avx.rs), which is what the reference runs when built with
// RUSTFLAGS=-Ctarget-cpu=native (its CI setting).
//
// Parity status: pinned against the reference's own golden tables
// (tests/golden/*.json, extracted from the reference's test sources) and
// differentially against the reference-shipped C++ superintervals library
// compiled from /root/reference into oracle/_ref/ (see oracle/Makefile).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

namespace orc {

struct Interval {
  int32_t first;
  int32_t last;
  uint64_t meta;  // `Position = usize` in the reference (IJ:46)
};

static constexpr uint32_t NIL = std::numeric_limits<uint32_t>::max();  // I::MAX, I = u32 (IJ:781)

template <int LANES>
struct alignas(LANES == 8 ? 32 : 8) Node;

// Scalar node: one interval per node (CT/nosimd.rs:33-51).
template <>
struct alignas(8) Node<1> {
  int32_t subtree_last;
  int32_t first;
  int32_t last;
  uint32_t left;
  uint32_t right;
  uint64_t meta;

  int32_t min_first() const { return first; }
  int32_t max_last() const { return last; }
};

// Chunk node: eight intervals, stored as first-1 / last+1 so that the strict
// compare (the only one AVX2 has) implements <= / >= (CT/avx.rs:62-84).
template <>
struct alignas(32) Node<8> {
  int32_t firsts_m1[8];
  int32_t lasts_p1[8];
  uint64_t meta[8];
  int32_t subtree_last;
  uint32_t left;
  uint32_t right;

  int32_t min_first() const { return int32_t(uint32_t(firsts_m1[0]) + 1u); }
  int32_t max_last() const {
    int32_t m = lasts_p1[0];
    for (int i = 1; i < 8; ++i) m = std::max(m, lasts_p1[i]);
    return int32_t(uint32_t(m) - 1u);
  }
};

struct NodeInfo {
  uint32_t depth;
  uint32_t inorder;
  uint32_t subtree_size;
  uint32_t parent;
  uint32_t max_depth;  // deepest node in this subtree (absolute depth)
  float hit_proportion;
};

template <int LANES>
class Tree {
 public:
  static constexpr uint32_t kSimpleCutoff = (LANES == 8) ? 8 : 64;  // CT/avx.rs:18, CT/nosimd.rs:24
  static constexpr float kDensityCutoff = 0.2f;                    // CT/nosimd.rs:29

  Tree() = default;

  explicit Tree(std::vector<Interval> ivs) {
    n_intervals_ = ivs.size();
    if (ivs.empty()) return;
    std::sort(ivs.begin(), ivs.end(), [](const Interval& a, const Interval& b) {
      return a.first != b.first ? a.first < b.first : a.last < b.last;
    });
    std::vector<Node<LANES>> sorted = make_nodes(ivs);
    ivs.clear();
    ivs.shrink_to_fit();
    relayout(std::move(sorted));
  }

  size_t size() const { return n_intervals_; }
  bool empty() const { return nodes_.empty(); }
  size_t bytes() const { return nodes_.size() * sizeof(Node<LANES>); }

  // Visit every stored interval b with b.first <= last && b.last >= first.
  template <class F>
  void query(int32_t first, int32_t last, F&& visit) const {
    if (nodes_.empty()) return;
    if constexpr (LANES == 8) {
#if defined(__AVX2__)
      __m256i qf = _mm256_set1_epi32(first), ql = _mm256_set1_epi32(last);
      descend8(root_, first, last, qf, ql, visit);
#else
      descend8_scalar(root_, first, last, visit);
#endif
    } else {
      descend1(root_, first, last, visit);
    }
  }

  size_t count(int32_t first, int32_t last) const {
    size_t c = 0;
    query(first, last, [&](uint64_t) { ++c; });
    return c;
  }

 private:
  std::vector<Node<LANES>> nodes_;
  size_t n_intervals_ = 0;
  uint32_t root_ = 0;

  // ---- node construction --------------------------------------------------
  static std::vector<Node<LANES>> make_nodes(const std::vector<Interval>& ivs) {
    std::vector<Node<LANES>> out;
    if constexpr (LANES == 1) {
      out.resize(ivs.size());
      for (size_t i = 0; i < ivs.size(); ++i) {
        out[i].subtree_last = ivs[i].last;
        out[i].first = ivs[i].first;
        out[i].last = ivs[i].last;
        out[i].left = out[i].right = NIL;
        out[i].meta = ivs[i].meta;
      }
    } else {
      const size_t nchunks = (ivs.size() + 7) / 8;
      out.resize(nchunks);
      for (size_t c = 0; c < nchunks; ++c) {
        Node<8>& nd = out[c];
        int32_t mx = std::numeric_limits<int32_t>::min();
        for (int j = 0; j < 8; ++j) {
          size_t i = c * 8 + j;
          // padding lanes are (MAX, MIN) (CT/avx.rs:410-427); the +-1 wraps like a release build
          int32_t f = i < ivs.size() ? ivs[i].first : std::numeric_limits<int32_t>::max();
          int32_t l = i < ivs.size() ? ivs[i].last : std::numeric_limits<int32_t>::min();
          nd.firsts_m1[j] = int32_t(uint32_t(f) - 1u);
          nd.lasts_p1[j] = int32_t(uint32_t(l) + 1u);
          nd.meta[j] = i < ivs.size() ? ivs[i].meta : 0;
          mx = std::max(mx, l);
        }
        nd.subtree_last = mx;
        nd.left = nd.right = NIL;
      }
    }
    return out;
  }

  // ---- pass 1: implicit BST over the sorted node array --------------------
  struct Walk {
    uint32_t root;
    int32_t subtree_first;
    float expected_hits;
  };

  static Walk annotate(std::vector<Node<LANES>>& nd, std::vector<NodeInfo>& info, uint32_t lo,
                       uint32_t hi, uint32_t depth, uint32_t parent, uint32_t& inorder) {
    const uint32_t mid = lo + (hi - lo) / 2;
    NodeInfo& me = info[mid];
    me.depth = depth;
    me.parent = parent;
    me.subtree_size = hi - lo;
    me.max_depth = depth;

    int32_t sub_first = nd[mid].min_first();
    float lhits = 0.f, rhits = 0.f;
    int32_t lspan = 0, rspan = 0;

    if (mid > lo) {
      Walk w = annotate(nd, info, lo, mid, depth + 1, mid, inorder);
      lhits = w.expected_hits;
      lspan = int32_t(uint32_t(nd[w.root].subtree_last) - uint32_t(w.subtree_first) + 1u);
      sub_first = w.subtree_first;
      nd[mid].subtree_last = std::max(nd[mid].subtree_last, nd[w.root].subtree_last);
      nd[mid].left = w.root;
      me.max_depth = std::max(me.max_depth, info[w.root].max_depth);
    }
    me.inorder = inorder++;
    if (mid + 1 < hi) {
      Walk w = annotate(nd, info, mid + 1, hi, depth + 1, mid, inorder);
      rhits = w.expected_hits;
      rspan = int32_t(uint32_t(nd[w.root].subtree_last) - uint32_t(w.subtree_first) + 1u);
      nd[mid].subtree_last = std::max(nd[mid].subtree_last, nd[w.root].subtree_last);
      nd[mid].right = w.root;
      me.max_depth = std::max(me.max_depth, info[w.root].max_depth);
    }

    const int32_t span = int32_t(uint32_t(nd[mid].subtree_last) - uint32_t(sub_first) + 1u);
    const int32_t own = int32_t(uint32_t(nd[mid].max_last()) - uint32_t(nd[mid].min_first()) + 1u);
    const float hits = (float(own) + float(lspan) * lhits + float(rspan) * rhits) / float(span);
    me.hit_proportion = hits / float(me.subtree_size);
    return Walk{mid, sub_first, hits};
  }

  // ---- pass 2: van Emde Boas order ----------------------------------------
  // `place(r, lo, hi, dmin, dmax)` appends, in memory order, the nodes of the
  // subtree rooted at sorted index r (covering sorted range [lo,hi)) whose
  // depth is <= dmax.  Split at the middle depth: bottom subtrees left of the
  // root first, then the top piece, then bottom subtrees right of the root.
  struct Layout {
    std::vector<Node<LANES>>* nd;
    const std::vector<NodeInfo>* info;
    std::vector<uint32_t> order;  // memory slot -> sorted index
  };

  static void bottoms(const Layout& L, uint32_t lo, uint32_t hi, uint32_t depth, uint32_t want,
                      std::vector<std::pair<uint32_t, uint32_t>>& out) {
    if (lo >= hi) return;
    if (depth == want) {
      out.emplace_back(lo, hi);
      return;
    }
    const uint32_t mid = lo + (hi - lo) / 2;
    bottoms(L, lo, mid, depth + 1, want, out);
    bottoms(L, mid + 1, hi, depth + 1, want, out);
  }

  static void place(Layout& L, uint32_t lo, uint32_t hi, uint32_t dmin, uint32_t dmax) {
    auto& nd = *L.nd;
    const auto& info = *L.info;
    const uint32_t r = lo + (hi - lo) / 2;
    const bool whole = info[r].max_depth <= dmax;  // piece has no descendants below it

    if (whole && (info[r].subtree_size <= kSimpleCutoff || info[r].hit_proportion >= kDensityCutoff)) {
      // simple subtree: a sorted run, each node records how many remain
      nd[lo].subtree_last = nd[r].subtree_last;
      uint32_t remaining = hi - lo;
      for (uint32_t i = lo; i < hi; ++i, --remaining) {
        nd[i].left = nd[i].right = remaining;
        L.order.push_back(i);
      }
      const uint32_t p = info[r].parent;
      if (p != NIL) {
        if (nd[p].left == r) nd[p].left = lo;
        else nd[p].right = lo;
      }
      return;
    }
    if (dmin == dmax || hi - lo == 1) {
      L.order.push_back(r);
      return;
    }
    const uint32_t pivot = dmin + (dmax - dmin) / 2;
    std::vector<std::pair<uint32_t, uint32_t>> lefts, rights;
    bottoms(L, lo, r, dmin + 1, pivot + 1, lefts);
    bottoms(L, r + 1, hi, dmin + 1, pivot + 1, rights);
    for (auto [a, b] : lefts) {
      const uint32_t rr = a + (b - a) / 2;
      place(L, a, b, pivot + 1, std::min(dmax, info[rr].max_depth));
    }
    place_top(L, lo, hi, dmin, pivot);
    for (auto [a, b] : rights) {
      const uint32_t rr = a + (b - a) / 2;
      place(L, a, b, pivot + 1, std::min(dmax, info[rr].max_depth));
    }
  }

  // The top piece is the same subtree truncated at depth `dmax`; it can only be
  // "whole" if nothing lies below the cut, which `place` re-checks.
  static void place_top(Layout& L, uint32_t lo, uint32_t hi, uint32_t dmin, uint32_t dmax) {
    place(L, lo, hi, dmin, dmax);
  }

  void relayout(std::vector<Node<LANES>> sorted) {
    const uint32_t n = uint32_t(sorted.size());
    std::vector<NodeInfo> info(n);
    uint32_t inorder = 0;
    annotate(sorted, info, 0, n, 0, NIL, inorder);

    Layout L{&sorted, &info, {}};
    L.order.reserve(n);
    const uint32_t r = n / 2;
    place(L, 0, n, 0, info[r].max_depth);

    std::vector<uint32_t> slot_of(n);
    for (uint32_t s = 0; s < n; ++s) slot_of[L.order[s]] = s;

    nodes_.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
      Node<LANES> nd = sorted[i];
      if (nd.left != nd.right) {
        if (nd.left != NIL) nd.left = slot_of[nd.left];
        if (nd.right != NIL) nd.right = slot_of[nd.right];
      }
      nodes_[slot_of[i]] = nd;
    }
    // a truncated top piece keeps its root; a simple run that replaced the whole tree starts at 0
    root_ = (sorted[0].left == sorted[0].right && sorted[0].left == n) ? slot_of[0] : slot_of[r];
  }

  // ---- queries --------------------------------------------------------------
  template <class F>
  void descend1(uint32_t at, int32_t first, int32_t last, F& visit) const {
    const Node<1>& nd = nodes_[at];
    if (nd.left == nd.right) {
      const Node<1>* p = &nodes_[at];
      const Node<1>* e = p + nd.right;
      for (; p != e; ++p) {
        if (last < p->first) break;
        if (first <= p->last) visit(p->meta);
      }
      return;
    }
    if (nd.first <= last && nd.last >= first) visit(nd.meta);
    if (nd.left != NIL && nodes_[nd.left].subtree_last >= first) descend1(nd.left, first, last, visit);
    if (nd.right != NIL && nd.first <= last && nodes_[nd.right].subtree_last >= first)
      descend1(nd.right, first, last, visit);
  }

  template <class F>
  static inline void chunk_scalar(const Node<8>& nd, int32_t first, int32_t last, F& visit) {
    for (int j = 0; j < 8; ++j)
      if (last > nd.firsts_m1[j] && nd.lasts_p1[j] > first) visit(nd.meta[j]);
  }

#if defined(__AVX2__)
  template <class F>
  static inline void chunk_avx(const Node<8>& nd, __m256i qf, __m256i ql, F& visit) {
    const __m256i f = _mm256_load_si256(reinterpret_cast<const __m256i*>(nd.firsts_m1));
    const __m256i l = _mm256_load_si256(reinterpret_cast<const __m256i*>(nd.lasts_p1));
    const __m256i hit = _mm256_and_si256(_mm256_cmpgt_epi32(ql, f), _mm256_cmpgt_epi32(l, qf));
    uint32_t m = uint32_t(_mm256_movemask_ps(_mm256_castsi256_ps(hit)));
    while (m) {
      const int j = __builtin_ctz(m);
      visit(nd.meta[j]);
      m &= m - 1;
    }
  }

  template <class F>
  void descend8(uint32_t at, int32_t first, int32_t last, __m256i qf, __m256i ql, F& visit) const {
    const Node<8>& nd = nodes_[at];
    if (nd.left == nd.right) {
      const Node<8>* p = &nodes_[at];
      const Node<8>* e = p + nd.right;
      for (; p != e; ++p) {
        if (last < p->min_first()) break;
        chunk_avx(*p, qf, ql, visit);
      }
      return;
    }
    chunk_avx(nd, qf, ql, visit);
    if (nd.left != NIL && nodes_[nd.left].subtree_last >= first)
      descend8(nd.left, first, last, qf, ql, visit);
    if (nd.right != NIL && nd.min_first() <= last && nodes_[nd.right].subtree_last >= first)
      descend8(nd.right, first, last, qf, ql, visit);
  }
#endif

  template <class F>
  void descend8_scalar(uint32_t at, int32_t first, int32_t last, F& visit) const {
    const Node<8>& nd = nodes_[at];
    if (nd.left == nd.right) {
      const Node<8>* p = &nodes_[at];
      const Node<8>* e = p + nd.right;
      for (; p != e; ++p) {
        if (last < p->min_first()) break;
        chunk_scalar(*p, first, last, visit);
      }
      return;
    }
    chunk_scalar(nd, first, last, visit);
    if (nd.left != NIL && nodes_[nd.left].subtree_last >= first)
      descend8_scalar(nd.left, first, last, visit);
    if (nd.right != NIL && nd.min_first() <= last && nodes_[nd.right].subtree_last >= first)
      descend8_scalar(nd.right, first, last, visit);
  }
};

}  // namespace orc
