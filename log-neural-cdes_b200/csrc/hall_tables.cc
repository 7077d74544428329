// K0 – Hall basis and tensor->Lie (Dynkin) projection tables, host side.
//
// Native replacement for data_dir/hall_set.py:60-256 of the reference: the
// basis is grown degree by degree (pairs (i,j), i<j, left(j) <= i), brackets
// of basis elements are rewritten into the basis with the Jacobi identity, and
// the projection of a tensor word is its right-nested bracket divided by its
// length.  Tables are immutable once built and cached per (width, depth).
#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

#include "lncde_b200.h"

namespace {

using Lin = std::vector<std::pair<int, long long>>;  // sparse integer combination of basis keys

struct HallTable {
  int width = 0, depth = 0;
  std::vector<std::pair<int, int>> basis;         // (left, right)
  std::vector<std::pair<int, int>> ranges;        // [lo, hi) per degree
  std::map<std::pair<int, int>, int> index;       // (left,right) -> key
  std::map<std::pair<int, int>, Lin> memo;        // bracket cache
  // t2l CSR (columns include the scalar slot)
  std::vector<int32_t> rowptr, cols;
  std::vector<float> vals;

  static void accumulate(std::map<int, long long>& acc, const Lin& v, long long s) {
    for (auto& kc : v) acc[kc.first] += s * kc.second;
  }
  static Lin compress(const std::map<int, long long>& acc) {
    Lin out;
    for (auto& kc : acc)
      if (kc.second != 0) out.push_back(kc);
    return out;
  }

  // [a, b] expressed in the basis
  const Lin& bracket(int a, int b) {
    static const Lin empty;
    if (a == b) return empty;
    auto key = std::make_pair(a, b);
    auto it = memo.find(key);
    if (it != memo.end()) return it->second;
    Lin out;
    if (b < a) {
      for (auto& kc : bracket(b, a)) out.push_back({kc.first, -kc.second});
    } else {
      auto hit = index.find(key);
      if (hit != index.end()) {
        out.push_back({hit->second, 1});
      } else {
        // [a,[l,r]] = [[a,l],r] - [[a,r],l]
        const int l = basis[b].first, r = basis[b].second;
        std::map<int, long long> acc;
        const Lin al = bracket(a, l);
        for (auto& kc : al) accumulate(acc, bracket(kc.first, r), kc.second);
        const Lin ar = bracket(a, r);
        for (auto& kc : ar) accumulate(acc, bracket(kc.first, l), -kc.second);
        out = compress(acc);
      }
    }
    return memo.emplace(key, std::move(out)).first->second;
  }

  void build(int w, int d) {
    width = w;
    depth = d;
    basis.push_back({0, 0});
    for (int i = 1; i <= w; ++i) {
      index[{0, i}] = (int)basis.size();
      basis.push_back({0, i});
    }
    ranges = {{0, 1}, {1, w + 1}};
    for (int deg = 2; deg <= d; ++deg) {
      const int lo = (int)basis.size();
      for (int ld = 1; 2 * ld <= deg; ++ld) {
        const auto ir = ranges[ld], jr = ranges[deg - ld];
        for (int i = ir.first; i < ir.second; ++i)
          for (int j = std::max(jr.first, i + 1); j < jr.second; ++j)
            if (basis[j].first <= i) {
              index[{i, j}] = (int)basis.size();
              basis.push_back({i, j});
            }
      }
      ranges.push_back({lo, (int)basis.size()});
    }
    build_t2l();
  }

  // Walk all words level by level carrying the right-nested bracket of the
  // suffix: rb(l1 l2..ln) = [l1, rb(l2..ln)].
  void build_t2l() {
    const int n = (int)basis.size();
    std::vector<std::map<int, double>> rows(n);
    // level 1
    std::vector<Lin> prev(width), cur;
    for (int i = 1; i <= width; ++i) prev[i - 1] = Lin{{i, 1}};
    long long level_off = 1;  // flat index of first word of this level
    for (int len = 1; len <= depth; ++len) {
      const long long count = (long long)prev.size();
      for (long long wi = 0; wi < count; ++wi)
        for (auto& kc : prev[wi]) rows[kc.first][(int)(level_off + wi)] += (double)kc.second / len;
      if (len == depth) break;
      // words of length len+1: first letter a (major), suffix = a word of length len
      cur.assign((size_t)count * width, Lin{});
      for (int a = 1; a <= width; ++a)
        for (long long wi = 0; wi < count; ++wi) {
          std::map<int, long long> acc;
          for (auto& kc : prev[wi]) accumulate(acc, bracket(a, kc.first), kc.second);
          cur[(size_t)(a - 1) * count + wi] = compress(acc);
        }
      level_off += count;
      prev.swap(cur);
    }
    rowptr.assign(n + 1, 0);
    for (int k = 0; k < n; ++k) {
      for (auto& cv : rows[k])
        if (cv.second != 0.0) {
          cols.push_back(cv.first);
          vals.push_back((float)cv.second);
        }
      rowptr[k + 1] = (int32_t)cols.size();
    }
  }
};

std::mutex g_mu;
std::map<std::pair<int, int>, std::unique_ptr<HallTable>> g_cache;

}  // namespace

namespace lncde {

// Largest table we are willing to build (tensor algebra dimension).
static bool supported(int width, int depth) {
  if (width < 1 || depth < 1 || depth > 6 || width > 512) return false;
  double dim = 0, p = 1;
  for (int k = 0; k <= depth; ++k) {
    dim += p;
    p *= width;
  }
  return dim <= 4.0e6;
}

const HallTable* get_table(int width, int depth) {
  if (!supported(width, depth)) return nullptr;
  std::lock_guard<std::mutex> lock(g_mu);
  auto key = std::make_pair(width, depth);
  auto it = g_cache.find(key);
  if (it == g_cache.end()) {
    std::unique_ptr<HallTable> t(new HallTable());
    t->build(width, depth);
    it = g_cache.emplace(key, std::move(t)).first;
  }
  return it->second.get();
}

}  // namespace lncde

extern "C" {

int lncde_hall_size(int width, int depth) {
  const HallTable* t = lncde::get_table(width, depth);
  return t ? (int)t->basis.size() : LNCDE_ERR_UNSUPPORTED;
}

int lncde_hall_basis(int width, int depth, int32_t* pairs) {
  if (!pairs) return LNCDE_ERR_NULL;
  const HallTable* t = lncde::get_table(width, depth);
  if (!t) return LNCDE_ERR_UNSUPPORTED;
  for (size_t k = 0; k < t->basis.size(); ++k) {
    pairs[2 * k] = t->basis[k].first;
    pairs[2 * k + 1] = t->basis[k].second;
  }
  return LNCDE_OK;
}

int64_t lncde_tensor_dim(int width, int depth) {
  if (width < 1 || depth < 0) return LNCDE_ERR_SHAPE;
  int64_t dim = 1;
  for (int k = 0; k < depth; ++k) dim = dim * width + 1;
  return dim;
}

int lncde_hall_t2l_nnz(int width, int depth) {
  const HallTable* t = lncde::get_table(width, depth);
  return t ? (int)t->cols.size() : LNCDE_ERR_UNSUPPORTED;
}

int lncde_hall_t2l_csr(int width, int depth, int32_t* rowptr, int32_t* cols, float* vals) {
  if (!rowptr || !cols || !vals) return LNCDE_ERR_NULL;
  const HallTable* t = lncde::get_table(width, depth);
  if (!t) return LNCDE_ERR_UNSUPPORTED;
  std::copy(t->rowptr.begin(), t->rowptr.end(), rowptr);
  std::copy(t->cols.begin(), t->cols.end(), cols);
  std::copy(t->vals.begin(), t->vals.end(), vals);
  return LNCDE_OK;
}

}  // extern "C"
