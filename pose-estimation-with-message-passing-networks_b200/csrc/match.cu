// Host side of the training-time label construction: the detection-to-ground-truth matching of every image of a
// batch (ConstructGraph.py:626-686 EDGE_LABEL_METHOD 4, :769-942 method 6, USE_NEIGHBOURS :704-727 / :890-911).
//
// The similarity matrix exp(-d^2 / factor) comes from the caller (float32, computed with the reference's own tensor
// operations so that every value carries the reference's bits); everything after it runs here, one std::thread per
// image, no Python in the loop: type masks, radius thresholds, the one or two linear sum assignments, the fill-in rule of
// method 6, and the neighbour / ambiguity rule.
//
// The assignment is scipy.optimize.linear_sum_assignment's algorithm restated -- the shortest augmenting path method of
// D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016, as implemented in SciPy's
// rectangular_lsap (SciPy 1.18, the version the reference fixtures were generated with): float64 costs, maximisation by
// negation, a tall matrix is transposed, the remaining columns are scanned in reverse index order and among equal
// shortest path costs an unassigned column wins.  The thresholded matrices are full of ties (zeros), so the tie rule is
// part of the result; tests/test_labels.py checks this restatement against scipy itself on tie-heavy matrices.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

#include "common.cuh"

namespace pgmp {
namespace {

// rows a[k] -> columns b[k], k < min(nr, nc); a ascending.  cost: row-major [nr][nc] float64.  Returns false when infeasible.
bool lsap(int64_t nr, int64_t nc, const double* cost_in, bool maximize, std::vector<int64_t>& a, std::vector<int64_t>& b) {
  a.clear();
  b.clear();
  if (nr == 0 || nc == 0) return true;
  const bool transpose = nc < nr;
  std::vector<double> temp((size_t)(nr * nc));
  if (transpose) {
    for (int64_t i = 0; i < nr; ++i)
      for (int64_t j = 0; j < nc; ++j) temp[(size_t)(j * nr + i)] = cost_in[i * nc + j];
    std::swap(nr, nc);
  } else {
    std::copy(cost_in, cost_in + nr * nc, temp.begin());
  }
  if (maximize)
    for (double& v : temp) v = -v;
  const double* cost = temp.data();
  const double inf = std::numeric_limits<double>::infinity();
  std::vector<double> u((size_t)nr, 0.0), v((size_t)nc, 0.0), spc((size_t)nc);
  std::vector<int64_t> path((size_t)nc, -1), col4row((size_t)nr, -1), row4col((size_t)nc, -1), remaining((size_t)nc);
  std::vector<char> SR((size_t)nr), SC((size_t)nc);
  for (int64_t cur = 0; cur < nr; ++cur) {
    double min_val = 0.0;
    int64_t num_remaining = nc;
    for (int64_t it = 0; it < nc; ++it) remaining[(size_t)it] = nc - it - 1;
    std::fill(SR.begin(), SR.end(), 0);
    std::fill(SC.begin(), SC.end(), 0);
    std::fill(spc.begin(), spc.end(), inf);
    int64_t sink = -1, i = cur;
    while (sink == -1) {
      int64_t index = -1;
      double lowest = inf;
      SR[(size_t)i] = 1;
      const double* row = cost + i * nc;
      const double ui = u[(size_t)i];
      for (int64_t it = 0; it < num_remaining; ++it) {
        const int64_t j = remaining[(size_t)it];
        const double r = min_val + row[j] - ui - v[(size_t)j];
        if (r < spc[(size_t)j]) {
          path[(size_t)j] = i;
          spc[(size_t)j] = r;
        }
        if (spc[(size_t)j] < lowest || (spc[(size_t)j] == lowest && row4col[(size_t)j] == -1)) {
          lowest = spc[(size_t)j];
          index = it;
        }
      }
      min_val = lowest;
      if (min_val == inf) return false;
      const int64_t j = remaining[(size_t)index];
      if (row4col[(size_t)j] == -1) sink = j; else i = row4col[(size_t)j];
      SC[(size_t)j] = 1;
      remaining[(size_t)index] = remaining[(size_t)--num_remaining];
    }
    u[(size_t)cur] += min_val;
    for (int64_t r = 0; r < nr; ++r)
      if (SR[(size_t)r] && r != cur) u[(size_t)r] += min_val - spc[(size_t)col4row[(size_t)r]];
    for (int64_t j = 0; j < nc; ++j)
      if (SC[(size_t)j]) v[(size_t)j] -= min_val - spc[(size_t)j];
    int64_t j = sink;
    while (true) {
      const int64_t r = path[(size_t)j];
      row4col[(size_t)j] = r;
      std::swap(col4row[(size_t)r], j);
      if (r == cur) break;
    }
  }
  a.resize((size_t)nr);
  b.resize((size_t)nr);
  if (transpose) {      // rows of the transposed problem are the caller's columns: report sorted by the caller's row
    std::vector<int64_t> order((size_t)nr);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return col4row[(size_t)x] < col4row[(size_t)y]; });
    for (int64_t k = 0; k < nr; ++k) {
      a[(size_t)k] = col4row[(size_t)order[(size_t)k]];
      b[(size_t)k] = order[(size_t)k];
    }
  } else {
    for (int64_t k = 0; k < nr; ++k) {
      a[(size_t)k] = k;
      b[(size_t)k] = col4row[(size_t)k];
    }
  }
  return true;
}

struct ImageOut {
  int n = 0;
  bool ok = true;
};

// one image; returns the number of (row, col) pairs written
ImageOut match_one(const pgmp_match_params& p, int b) {
  ImageOut res;
  const int G = p.num_gt[b], n = p.num_det[b];
  int32_t* out_r = p.match_row + (int64_t)b * p.cap;
  int32_t* out_c = p.match_col + (int64_t)b * p.cap;
  uint8_t* amb = p.ambiguous ? p.ambiguous + (int64_t)b * p.max_det : nullptr;
  if (amb) std::fill(amb, amb + p.max_det, (uint8_t)0);
  if (G <= 0 || n <= 0) return res;
  const float* sim = p.sim + (int64_t)b * p.sim_stride_b;
  const int32_t* gt_type = p.gt_type + (int64_t)b * p.max_gt;
  const int32_t* det_type = p.det_type + (int64_t)b * p.max_det;
  const size_t sz = (size_t)G * n;
  const float thr = p.matching_radius;
  std::vector<float> same(sz), diff;
  std::vector<double> cd(sz);
  for (int r = 0; r < G; ++r)
    for (int c = 0; c < n; ++c) {
      const float s = sim[(int64_t)r * p.sim_stride_g + c];
      const bool other = gt_type[r] != det_type[c];
      same[(size_t)r * n + c] = (other || s < thr) ? 0.f : s;
    }
  std::vector<int64_t> rows, cols, rows2, cols2;
  for (size_t k = 0; k < sz; ++k) cd[k] = (double)same[k];
  if (!lsap(G, n, cd.data(), true, rows, cols)) { res.ok = false; return res; }
  std::vector<int64_t> kr, kc;      // the kept matches
  if (p.method == 4) {
    for (size_t k = 0; k < rows.size(); ++k)
      if (same[(size_t)rows[k] * n + cols[k]] != 0.f) { kr.push_back(rows[k]); kc.push_back(cols[k]); }
  } else {
    diff.resize(sz);
    for (int r = 0; r < G; ++r)
      for (int c = 0; c < n; ++c) {
        const float s = sim[(int64_t)r * p.sim_stride_g + c];
        const bool other = gt_type[r] != det_type[c];
        diff[(size_t)r * n + c] = (!other || s < thr) ? 0.f : s;
      }
    for (size_t k = 0; k < sz; ++k) cd[k] = (double)diff[k];
    if (!lsap(G, n, cd.data(), true, rows2, cols2)) { res.ok = false; return res; }
    // same type first, any other type as a fill-in (:811-830); position k of the two solutions belongs together
    // (the reference tests `cost_diff[sol_diff] + cost_same[sol_same]` AFTER the fill-in has been written into the column
    // array of sol_same, so the same-type term is read at the filled-in column)
    for (size_t k = 0; k < rows.size(); ++k) {
      int64_t c = cols[k];
      if (!(same[(size_t)rows[k] * n + c] != 0.f) && k < cols2.size()) c = cols2[k];
      const float vs = same[(size_t)rows[k] * n + c];
      const float vd = k < rows2.size() ? diff[(size_t)rows2[k] * n + cols2[k]] : 0.f;
      if (vd + vs != 0.f) { kr.push_back(rows[k]); kc.push_back(c); }
    }
  }
  int w = 0;
  for (size_t k = 0; k < kr.size() && w < p.cap; ++k, ++w) { out_r[w] = (int32_t)kr[k]; out_c[w] = (int32_t)kc[k]; }
  if (p.use_neighbours) {
    // further candidates within the inclusion radius of a matched joint; candidates claimed by more than one joint are
    // ambiguous and leave the loss.  Method 4 works on the masked + thresholded matrix, method 6 on the raw similarity.
    std::vector<float> cost(sz);
    if (p.method == 4) cost = same;
    else
      for (int r = 0; r < G; ++r)
        for (int c = 0; c < n; ++c) cost[(size_t)r * n + c] = sim[(int64_t)r * p.sim_stride_g + c];
    for (float& x : cost)
      if (x < p.inclusion_radius) x = 0.f;
    for (int64_t c : kc)
      for (int r = 0; r < G; ++r) cost[(size_t)r * n + c] = 0.f;
    for (int c = 0; c < n; ++c) {
      int cnt = 0;
      for (int r = 0; r < G; ++r) cnt += cost[(size_t)r * n + c] != 0.f;
      if (cnt > 1) {
        amb[c] = 1;
        for (int r = 0; r < G; ++r) cost[(size_t)r * n + c] = 0.f;
      }
    }
    std::vector<char> has_match((size_t)G, 0);
    for (int64_t r : kr) has_match[(size_t)r] = 1;
    for (int r = 0; r < G && w < p.cap; ++r) {
      if (!has_match[(size_t)r]) continue;      // joints without a match of their own take no neighbours
      for (int c = 0; c < n && w < p.cap; ++c)
        if (cost[(size_t)r * n + c] != 0.f) { out_r[w] = r; out_c[w] = c; ++w; }
    }
  }
  res.n = w;
  if (p.node_person) {        // per-node labels of this image
    const int64_t off = p.node_offsets[b];
    const int32_t* person = p.gt_person + (int64_t)b * p.max_gt;
    for (int k = 0; k < w; ++k) {
      const int64_t node = off + out_c[k];
      p.node_person[node] = person[out_r[k]];
      p.node_class[node] = gt_type[out_r[k]];
      p.node_label[node] = 1.f;
    }
    if (p.use_neighbours && p.node_ambiguous)
      for (int c = 0; c < n; ++c) p.node_ambiguous[off + c] = amb[c];
  }
  return res;
}

void label_args_one(const pgmp_label_args_params& p, int b) {
  const int P = p.max_persons, J = p.num_joints;
  const int64_t off = p.node_offsets[b];
  const int n = (int)(p.node_offsets[b + 1] - off);
  const float* gt = p.gt + (int64_t)b * P * J * 3;
  const float* fac = p.factors + (int64_t)b * P * J;
  int32_t* det_type = p.det_type + (int64_t)b * p.max_det;
  std::vector<float> dx((size_t)n), dy((size_t)n);
  for (int c = 0; c < n; ++c) {
    dx[(size_t)c] = (float)p.det[(off + c) * 3 + 0];
    dy[(size_t)c] = (float)p.det[(off + c) * 3 + 1];
    det_type[c] = (int32_t)p.det[(off + c) * 3 + 2];
  }
  int g = 0;
  for (int pi = 0; pi < P; ++pi)
    for (int j = 0; j < J; ++j) {
      const float* q = gt + ((int64_t)pi * J + j) * 3;
      if (q[2] == 0.f) continue;
      if (g >= p.max_gt) continue;
      // gt[..., :2].round().float().clamp(0, clamp_max): round half to even, like torch.round
      const float px = std::fmin(std::fmax(std::nearbyint(q[0]), 0.f), p.clamp_max);
      const float py = std::fmin(std::fmax(std::nearbyint(q[1]), 0.f), p.clamp_max);
      const float f = fac[(int64_t)pi * J + j];
      float* row = p.arg + ((int64_t)b * p.max_gt + g) * p.max_det;
      for (int c = 0; c < n; ++c) {
        const float ax = px - dx[(size_t)c], ay = py - dy[(size_t)c];
        const float d2 = ax * ax + ay * ay;            // pow(2).sum(dim): two separately rounded squares, one addition
        float a = -d2 / f;
        if (a < p.min_arg) a = p.min_arg;
        row[c] = a;
      }
      p.gt_type[(int64_t)b * p.max_gt + g] = j;
      p.gt_person[(int64_t)b * p.max_gt + g] = pi;
      ++g;
    }
  p.num_gt[b] = g;
}

}  // namespace
}  // namespace pgmp

extern "C" int pgmp_match_labels(const pgmp_match_params* p) {
  using namespace pgmp;
  if (!p || p->batch < 0 || !p->sim || !p->num_gt || !p->num_det || !p->gt_type || !p->det_type || !p->match_row ||
      !p->match_col || !p->num_match || p->cap < 1)
    return set_error(PGMP_ERR_INVALID, "pgmp_match_labels: null pointer / bad sizes");
  if (p->method != 4 && p->method != 6) return set_error(PGMP_ERR_INVALID, "EDGE_LABEL_METHOD %d (4 and 6 are implemented)", p->method);
  if (p->use_neighbours && !p->ambiguous) return set_error(PGMP_ERR_INVALID, "USE_NEIGHBOURS without the ambiguity output");
  std::vector<char> ok((size_t)p->batch, 1);
  auto work = [&](int first, int step) {
    for (int b = first; b < p->batch; b += step) {
      const ImageOut r = match_one(*p, b);
      p->num_match[b] = r.n;
      ok[(size_t)b] = r.ok;
    }
  };
  int nt = p->num_threads > 0 ? p->num_threads : 1;
  if (nt > p->batch) nt = p->batch;
  if (nt <= 1) {
    work(0, 1);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, t, nt);
    for (auto& th : pool) th.join();
  }
  for (int b = 0; b < p->batch; ++b)
    if (!ok[(size_t)b]) return set_error(PGMP_ERR_INVALID, "linear sum assignment infeasible for image %d", b);
  return PGMP_OK;
}

extern "C" int pgmp_label_similarity_args(const pgmp_label_args_params* p) {
  using namespace pgmp;
  if (!p || p->batch < 0 || !p->det || !p->node_offsets || !p->gt || !p->factors || !p->arg || !p->num_gt || !p->gt_type ||
      !p->gt_person || !p->det_type || p->max_gt < 1 || p->max_det < 1)
    return set_error(PGMP_ERR_INVALID, "pgmp_label_similarity_args: null pointer / bad sizes");
  for (int b = 0; b < p->batch; ++b)
    if (p->node_offsets[b + 1] - p->node_offsets[b] > p->max_det)
      return set_error(PGMP_ERR_INVALID, "image %d has more candidates than max_det", b);
  auto work = [&](int first, int step) {
    for (int b = first; b < p->batch; b += step) label_args_one(*p, b);
  };
  int nt = p->num_threads > 0 ? p->num_threads : 1;
  if (nt > p->batch) nt = p->batch;
  if (nt <= 1) {
    work(0, 1);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, t, nt);
    for (auto& th : pool) th.join();
  }
  return PGMP_OK;
}

// the assignment alone (tests: against scipy.optimize.linear_sum_assignment)
extern "C" int pgmp_linear_sum_assignment(const double* cost, int64_t nr, int64_t nc, int maximize, int64_t* rows, int64_t* cols) {
  using namespace pgmp;
  if (!cost || !rows || !cols || nr < 0 || nc < 0) return set_error(PGMP_ERR_INVALID, "pgmp_linear_sum_assignment: bad arguments");
  std::vector<int64_t> a, b;
  if (!lsap(nr, nc, cost, maximize != 0, a, b)) return set_error(PGMP_ERR_INVALID, "cost matrix is infeasible");
  std::copy(a.begin(), a.end(), rows);
  std::copy(b.begin(), b.end(), cols);
  return PGMP_OK;
}
