// control.cu — host-side tracking control in C++ (no device code): the per-frame association and the per-shot link graph of
// TrackingByDetection.  Replaces the Python of
//   _match / _associate                      pyannote/video/tracking.py:129-182   (overlap matrix, Munkres on max - overlap)
//   the shot's nx.DiGraph + connected components  :209-244,340-347
//   _fix                                      :261-296   (positions seen at the same time are averaged)
//   _fill_gaps                                :298-329   (tracks whose end / start boxes overlap within track_max_gap are joined)
// with the same arithmetic (doubles, the same comparison and tie rules as pyannote_video_b200/tracking.py, whose tracks equal
// the reference's own on the golden scenarios).  Matrices are tiny (trackers x detections of one frame), so the
// Kuhn–Munkres potentials method runs on the host; the tracker bank itself stays on the GPU (csrc/tracker.cu).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

namespace {

struct DR {   // dlib drectangle semantics: continuous, empty if r < l or b < t (geometry.DRect)
  double l, t, r, b;
};
inline double dr_area(const DR& a) {
  if (a.t > a.b || a.l > a.r) return 0.0;
  return (a.r - a.l) * (a.b - a.t);
}
inline double match_overlap(const DR& a, const DR& b, double ratio) {
  const DR i{std::max(a.l, b.l), std::max(a.t, b.t), std::min(a.r, b.r), std::min(a.b, b.b)};
  double ov = dr_area(i);
  if (ov < ratio * dr_area(a) || ov < ratio * dr_area(b)) ov = 0.0;
  return ov;
}

// minimum-cost perfect matching of the rows of a square cost matrix (potentials method; the same sequence of operations
// as pyannote_video_b200/hungarian.py, hence the same assignment when costs tie)
void hungarian(const std::vector<double>& cost, int n, std::vector<int>& row_of_col) {
  const double INF = INFINITY;
  std::vector<double> u(n + 1, 0.0), v(n + 1, 0.0), minv(n + 1);
  std::vector<int> p(n + 1, 0), way(n + 1, 0);
  std::vector<char> used(n + 1);
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    std::fill(minv.begin(), minv.end(), INF);
    std::fill(used.begin(), used.end(), 0);
    while (true) {
      used[j0] = 1;
      const int i0 = p[j0];
      double delta = INF;
      int j1 = 0;
      for (int j = 1; j <= n; ++j) {
        if (!used[j]) {
          const double cur = cost[(size_t)(i0 - 1) * n + (j - 1)] - u[i0] - v[j];
          if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
          if (minv[j] < delta) { delta = minv[j]; j1 = j; }
        }
      }
      for (int j = 0; j <= n; ++j) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
      if (p[j0] == 0) break;
    }
    while (true) {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
      if (j0 == 0) break;
    }
  }
  row_of_col.assign(n, -1);
  for (int j = 1; j <= n; ++j) row_of_col[j - 1] = p[j] - 1;
}

struct Node {
  double t;
  double box[4];
  int status;   // 0 forward, 1 detection, 2 backward  (= the reference's sort order forward < detection < backward)
  bool operator<(const Node& o) const {
    if (t != o.t) return t < o.t;
    for (int k = 0; k < 4; ++k)
      if (box[k] != o.box[k]) return box[k] < o.box[k];
    return status < o.status;
  }
};

struct Graph {
  std::map<Node, int> index;
  std::vector<Node> nodes;
  std::vector<int> parent;
  int add(const Node& n) {
    auto it = index.find(n);
    if (it != index.end()) return it->second;
    const int i = (int)nodes.size();
    index.emplace(n, i);
    nodes.push_back(n);
    parent.push_back(i);
    return i;
  }
  int find(int i) {
    int root = i;
    while (parent[root] != root) root = parent[root];
    while (parent[i] != root) { const int nx = parent[i]; parent[i] = root; i = nx; }
    return root;
  }
  void link(const Node& a, const Node& b) {
    const int ra = find(add(a)), rb = find(add(b));
    if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
  }
};

struct Row {
  double t;
  long long box[4];
  int nf, nd, nb;   // how many forward / detection / backward positions were merged (status string = '+'.join in that order)
  int error;
};
typedef std::vector<Row> Track;

// Python's round(): half to even
inline long long py_round(double v) {
  const double r = std::nearbyint(v);   // default rounding mode: to nearest, ties to even
  return (long long)r;
}

Track fix_track(std::vector<Node> comp, double ratio) {
  // sorted(track): by (t, box tuple, status STRING) — 'backward' < 'detection' < 'forward' alphabetically
  auto str_rank = [](int s) { return s == 2 ? 0 : (s == 1 ? 1 : 2); };
  std::sort(comp.begin(), comp.end(), [&](const Node& a, const Node& b) {
    if (a.t != b.t) return a.t < b.t;
    for (int k = 0; k < 4; ++k)
      if (a.box[k] != b.box[k]) return a.box[k] < b.box[k];
    return str_rank(a.status) < str_rank(b.status);
  });
  Track out;
  size_t i = 0;
  while (i < comp.size()) {
    size_t j = i;
    while (j < comp.size() && comp[j].t == comp[i].t) ++j;
    Row r;
    r.t = comp[i].t;
    r.nf = r.nd = r.nb = 0;
    r.error = 0;
    for (size_t a = i; a < j && !r.error; ++a)
      for (size_t b = a + 1; b < j; ++b) {
        const DR ra{comp[a].box[0], comp[a].box[1], comp[a].box[2], comp[a].box[3]};
        const DR rb{comp[b].box[0], comp[b].box[1], comp[b].box[2], comp[b].box[3]};
        if (match_overlap(ra, rb, ratio) == 0.0) { r.error = 1; break; }
      }
    const double n = (double)(j - i);
    for (int k = 0; k < 4; ++k) {
      // np.mean over the group: pairwise summation in numpy for n < 8 is a plain left-to-right sum
      double s = 0.0;
      for (size_t a = i; a < j; ++a) s += comp[a].box[k];
      r.box[k] = py_round(s / n);
    }
    for (size_t a = i; a < j; ++a) {
      if (comp[a].status == 0) ++r.nf;
      else if (comp[a].status == 1) ++r.nd;
      else ++r.nb;
    }
    out.push_back(r);
    i = j;
  }
  return out;
}

void min_max_t(const Track& tr, double& lo, double& hi) {
  lo = INFINITY;
  hi = -INFINITY;
  for (const Row& r : tr) { lo = std::min(lo, r.t); hi = std::max(hi, r.t); }
}

bool track_less(const Track& a, const Track& b) {
  double al, ah, bl, bh;
  min_max_t(a, al, ah);
  min_max_t(b, bl, bh);
  if (al != bl) return al < bl;
  return ah < bh;
}

std::vector<Track> fill_gaps(std::vector<Track> tracks, double ratio, double max_gap) {
  std::stable_sort(tracks.begin(), tracks.end(), track_less);
  const int n = (int)tracks.size();
  std::vector<int> parent(n);
  for (int i = 0; i < n; ++i) parent[i] = i;
  auto find = [&](int i) {
    while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; }
    return i;
  };
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const double ti = tracks[i].back().t, tj = tracks[j].front().t;
      if (tj < ti || tj - ti > max_gap) continue;
      const Row& a = tracks[i].back();
      const Row& b = tracks[j].front();
      const DR ra{(double)a.box[0], (double)a.box[1], (double)a.box[2], (double)a.box[3]};
      const DR rb{(double)b.box[0], (double)b.box[1], (double)b.box[2], (double)b.box[3]};
      if (match_overlap(ra, rb, ratio) != 0.0) {
        const int ri = find(i), rj = find(j);
        if (ri != rj) parent[std::max(ri, rj)] = std::min(ri, rj);
      }
    }
  std::map<int, std::vector<int>> groups;
  for (int i = 0; i < n; ++i) groups[find(i)].push_back(i);
  std::vector<Track> out;
  for (auto& g : groups) {   // ascending root = sorted(groups); members ascending = sorted(groups[r])
    Track t;
    for (int k : g.second) t.insert(t.end(), tracks[k].begin(), tracks[k].end());
    out.push_back(t);
  }
  return out;
}

struct Shot {
  Graph g;
  std::vector<Track> result;
};

}  // namespace

extern "C" int pv_ctl_associate(const double* positions, int n_trackers, const double* detections, int n_detections,
                                double min_overlap_ratio, int* match) {
  PV_REQUIRE(match || n_detections == 0, "pv_ctl_associate: null output");
  for (int d = 0; d < n_detections; ++d) match[d] = -1;
  if (n_trackers < 1 || n_detections < 1) return PV_OK;
  PV_REQUIRE(positions && detections, "pv_ctl_associate: null argument");
  const int n = std::max(n_trackers, n_detections);
  std::vector<double> ov((size_t)n * n, 0.0);
  double mx = 0.0;
  for (int t = 0; t < n_trackers; ++t)
    for (int d = 0; d < n_detections; ++d) {
      const DR a{positions[4 * t], positions[4 * t + 1], positions[4 * t + 2], positions[4 * t + 3]};
      const DR b{detections[4 * d], detections[4 * d + 1], detections[4 * d + 2], detections[4 * d + 3]};
      const double o = match_overlap(a, b, min_overlap_ratio);
      ov[(size_t)t * n + d] = o;
      mx = std::max(mx, o);
    }
  std::vector<double> cost((size_t)n * n);
  for (size_t i = 0; i < cost.size(); ++i) cost[i] = mx - ov[i];
  std::vector<int> row_of_col;
  hungarian(cost, n, row_of_col);
  for (int d = 0; d < n_detections; ++d) {
    const int t = row_of_col[d];
    if (t >= 0 && t < n_trackers && ov[(size_t)t * n + d] > 0.0) match[d] = t;
  }
  return PV_OK;
}

extern "C" int pv_ctl_shot_create(void** out_handle) {
  PV_REQUIRE(out_handle, "pv_ctl_shot_create: null argument");
  *out_handle = new Shot();
  return PV_OK;
}

extern "C" int pv_ctl_shot_destroy(void* h) {
  delete static_cast<Shot*>(h);
  return PV_OK;
}

extern "C" int pv_ctl_shot_add(void* h, double t, const double* box, int status) {
  PV_REQUIRE(h && box && status >= 0 && status <= 2, "pv_ctl_shot_add: bad argument");
  Node n{t, {box[0], box[1], box[2], box[3]}, status};
  static_cast<Shot*>(h)->g.add(n);
  return PV_OK;
}

extern "C" int pv_ctl_shot_link(void* h, double ta, const double* box_a, int sa, double tb, const double* box_b, int sb) {
  PV_REQUIRE(h && box_a && box_b && sa >= 0 && sa <= 2 && sb >= 0 && sb <= 2, "pv_ctl_shot_link: bad argument");
  const Node a{ta, {box_a[0], box_a[1], box_a[2], box_a[3]}, sa};
  const Node b{tb, {box_b[0], box_b[1], box_b[2], box_b[3]}, sb};
  static_cast<Shot*>(h)->g.link(a, b);
  return PV_OK;
}

extern "C" int pv_ctl_shot_finish(void* h, double min_overlap_ratio, double max_gap, int* n_tracks, int* n_rows) {
  PV_REQUIRE(h && n_tracks && n_rows, "pv_ctl_shot_finish: null argument");
  Shot* s = static_cast<Shot*>(h);
  Graph& g = s->g;
  std::map<int, std::vector<Node>> comps;   // root (= smallest = first inserted index) -> nodes in insertion order
  for (int i = 0; i < (int)g.nodes.size(); ++i) comps[g.find(i)].push_back(g.nodes[i]);
  std::vector<Track> tracks;
  for (auto& c : comps) tracks.push_back(fix_track(c.second, min_overlap_ratio));
  tracks = fill_gaps(tracks, min_overlap_ratio, max_gap);
  std::stable_sort(tracks.begin(), tracks.end(), track_less);
  s->result = tracks;
  *n_tracks = (int)tracks.size();
  int rows = 0;
  for (const Track& t : tracks) rows += (int)t.size();
  *n_rows = rows;
  return PV_OK;
}

extern "C" int pv_ctl_shot_tracks(void* h, int* track_len, double* row_t, long long* row_box, int* row_counts) {
  PV_REQUIRE(h && track_len && row_t && row_box && row_counts, "pv_ctl_shot_tracks: null argument");
  const Shot* s = static_cast<const Shot*>(h);
  int k = 0, r = 0;
  for (const Track& t : s->result) {
    track_len[k++] = (int)t.size();
    for (const Row& row : t) {
      row_t[r] = row.t;
      for (int q = 0; q < 4; ++q) row_box[4 * r + q] = row.box[q];
      row_counts[4 * r + 0] = row.nf;
      row_counts[4 * r + 1] = row.nd;
      row_counts[4 * r + 2] = row.nb;
      row_counts[4 * r + 3] = row.error;
      ++r;
    }
  }
  return PV_OK;
}
