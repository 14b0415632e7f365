// gndt_twodmap_adapter.h — header-only bridge from the libgndt.so result tables to the
// reference's own containers, so that its host-side planner (TwoDmap::computeCost,
// AstarPlanar in GlobalPlan.h, the show* marker builders) keeps running unmodified.
//
// Include AFTER the reference's "map2D.h" (it uses daysun::TwoDmap / Cell / Slope / OcNode
// and the reference's countMorton).  Not part of the C ABI; plain C++98-compatible code.
//
// What it rebuilds, and where the reference builds the same thing:
//   morton_list                 xy keys in first-seen order              src/receiver.cpp:70
//   map_cell / Cell::map_slope  one Cell per occupied column, one Slope  include/map2D.h:598-599,
//                               per surface voxel                        632-642 / 646-659
//   map_xy                      OcNode{z, morton, xyz_centroid,          src/receiver.cpp:62-69,85-90;
//                               covariance_matrix, N} in first-seen      include/map2D.h:623-625
//                               order inside each column
#ifndef GNDT_TWODMAP_ADAPTER_H
#define GNDT_TWODMAP_ADAPTER_H

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <deque>
#include <list>
#include <map>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/gndt.h"
#include "../include/gndt_lookup.h"

namespace gndt_adapter {

// "A"+countMorton(|sx|,|sy|) — the reference's own key builder (map2D.h:971-972)
inline std::string morton_key(int sx, int sy) {
  const char q = sx > 0 ? (sy > 0 ? 'A' : 'B') : (sy > 0 ? 'C' : 'D');
  return std::string(1, q) + countMorton(std::abs(sx), std::abs(sy));
}

struct ByFirst {
  const gndt_voxel *v;
  explicit ByFirst(const gndt_voxel *vv) : v(vv) {}
  bool operator()(unsigned a, unsigned b) const { return v[a].first_index < v[b].first_index; }
};
struct ColByFirst {
  const gndt_column *c;
  explicit ColByFirst(const gndt_column *cc) : c(cc) {}
  bool operator()(unsigned a, unsigned b) const { return c[a].first_index < c[b].first_index; }
};

// Fill `m` (freshly constructed, parameters already set with setLen/setZLen/setInterval)
// from the tables of one build.  `with_map_xy` also recreates the OcNode multimap (needed
// by Slope::countUp in the "true" demand and by showInital); the slope demand's planner
// only reads map_cell.
inline void fill_twodmap(daysun::TwoDmap &m, const float origin[3], const gndt_voxel *vox, size_t n_vox,
                         const gndt_slope *slopes, size_t n_slopes, const gndt_column *cols, size_t n_cols,
                         bool with_map_xy) {
  (void)n_vox;
  m.setCloudFirst(octomath::Vector3(origin[0], origin[1], origin[2]));
  std::vector<unsigned> order(n_cols);
  for (size_t c = 0; c < n_cols; ++c) order[c] = (unsigned)c;
  std::sort(order.begin(), order.end(), ColByFirst(cols));
  for (size_t k = 0; k < n_cols; ++k) {
    const gndt_column &col = cols[order[k]];
    const std::string key = morton_key(col.sx, col.sy);
    m.morton_list.push_back(key);
    daysun::Cell *cell = new daysun::Cell(key);
    m.map_cell.insert(std::map<std::string, daysun::Cell *>::value_type(key, cell));
    for (unsigned s = col.slope_begin; s < col.slope_begin + col.slope_count && s < n_slopes; ++s) {
      const gndt_slope &g = slopes[s];
      daysun::Slope *slope = new daysun::Slope();
      slope->morton_xy = key;
      slope->morton_z = g.sz;
      slope->normal << g.normal[0], g.normal[1], g.normal[2];
      slope->rough = g.rough;
      slope->mean(0) = g.mean[0]; slope->mean(1) = g.mean[1]; slope->mean(2) = g.mean[2];
      slope->h = slope->g = slope->f = FLT_MAX;
      // slope demand: never assigned on the initial build (map2D.h:636) and a Slope never
      // carries UP there; true demand: what the lazy Slope::countUp (map2D.h:275) would store
      slope->up = (g.flags & GNDT_F_UP) != 0;
      slope->down = (g.flags & GNDT_F_DOWN) != 0;
      slope->father = NULL;
      cell->map_slope.insert(std::make_pair(g.sz, slope));
    }
    if (with_map_xy) {
      std::vector<unsigned> vs(col.voxel_count);
      for (unsigned i = 0; i < col.voxel_count; ++i) vs[i] = col.voxel_begin + i;
      std::sort(vs.begin(), vs.end(), ByFirst(vox));  // multimap equal-range order = first seen
      for (size_t i = 0; i < vs.size(); ++i) {
        const gndt_voxel &v = vox[vs[i]];
        daysun::OcNode *node = new daysun::OcNode();
        node->morton = key;
        node->z = v.sz;
        if (v.flags & GNDT_F_FITTED) {
          node->N = (int)v.count;
          node->xyz_centroid << v.mean[0], v.mean[1], v.mean[2];
          node->covariance_matrix(0, 0) = v.scatter[0]; node->covariance_matrix(0, 1) = v.scatter[1];
          node->covariance_matrix(0, 2) = v.scatter[2]; node->covariance_matrix(1, 1) = v.scatter[3];
          node->covariance_matrix(1, 2) = v.scatter[4]; node->covariance_matrix(2, 2) = v.scatter[5];
          node->covariance_matrix(1, 0) = v.scatter[1]; node->covariance_matrix(2, 0) = v.scatter[2];
          node->covariance_matrix(2, 1) = v.scatter[4];
          node->_isSlope = (v.flags & GNDT_F_SLOPE) != 0;
        }
        m.map_xy.insert(std::multimap<std::string, daysun::OcNode *>::value_type(key, node));
      }
    }
  }
}

// Integer-keyed view of map_cell for the planner's inner loops:
// `find` replaces  map_cell.find(quadrant + countMorton(x, y))  (map2D.h:269-272, 396-398;
// GlobalPlan.h:58-61) and `neighbor` replaces  countLRFB + map_cell.find  (map2D.h:197-263,
// 539-546) by binary searches over the column table: no string is built or compared.
// Build it once after fill_twodmap(); it points into `m` and into the caller's column table.
struct CellIndex {
  const gndt_column *cols;
  size_t n_cols;
  std::vector<daysun::Cell *> cells;  // cells[i] = the Cell of cols[i]

  CellIndex(daysun::TwoDmap &m, const gndt_column *c, size_t n) : cols(c), n_cols(n), cells(n) {
    for (size_t i = 0; i < n; ++i) {
      std::map<std::string, daysun::Cell *>::iterator it = m.map_cell.find(morton_key(c[i].sx, c[i].sy));
      cells[i] = it == m.map_cell.end() ? NULL : it->second;
    }
  }
  daysun::Cell *find(int sx, int sy) const {
    const int64_t i = gndtl_find_column(cols, n_cols, sx, sy);
    return i < 0 ? NULL : cells[(size_t)i];
  }
  // dir: 0 left, 1 right, 2 forward, 3 back (leftMtn / rightMtn / forMtn / backMtn of countLRFB)
  daysun::Cell *neighbor(int sx, int sy, int dir) const {
    const int64_t i = gndtl_neighbor_column(cols, n_cols, sx, sy, dir);
    return i < 0 ? NULL : cells[(size_t)i];
  }
  daysun::Slope *find_slope(int sx, int sy, int sz) const {
    daysun::Cell *c = find(sx, sy);
    if (!c) return NULL;
    std::map<int, daysun::Slope *, CmpByKeyUD>::iterator it = c->map_slope.find(sz);
    return it == c->map_slope.end() ? NULL : it->second;
  }
};

// The traversability graph of gndt_build_edges (CSR over the slope table) bound to the host
// Slope objects: AccessibleNeighborsFast returns the list TwoDmap::AccessibleNeighbors
// (map2D.h:530-548) returns — same Slopes, same order — without building a key string,
// searching map_cell or evaluating countAngle: a slice of an integer array.
struct SlopeGraph {
  std::vector<daysun::Slope *> slope;                 // slope-table index -> host object
  std::map<const daysun::Slope *, unsigned> index;    // host object -> slope-table index
  std::vector<unsigned char> last_in_cell;            // 1: no Slope above it in its Cell
  const uint32_t *offsets, *targets;

  SlopeGraph(daysun::TwoDmap &m, const gndt_column *cols, size_t n_cols, size_t n_slopes, const uint32_t *off, const uint32_t *tgt)
      : slope(n_slopes, (daysun::Slope *)NULL), last_in_cell(n_slopes, 0), offsets(off), targets(tgt) {
    for (size_t c = 0; c < n_cols; ++c) {
      std::map<std::string, daysun::Cell *>::iterator it = m.map_cell.find(morton_key(cols[c].sx, cols[c].sy));
      if (it == m.map_cell.end()) continue;
      unsigned k = cols[c].slope_begin;  // Cell::map_slope iterates in ascending z = slope-table order
      for (std::map<int, daysun::Slope *, CmpByKeyUD>::iterator sit = it->second->map_slope.begin(); sit != it->second->map_slope.end(); ++sit, ++k)
        if (k < n_slopes) { slope[k] = sit->second; index[sit->second] = k; }
      if (cols[c].slope_count) last_in_cell[cols[c].slope_begin + cols[c].slope_count - 1] = 1;
    }
  }
  std::list<daysun::Slope *> AccessibleNeighborsFast(daysun::Slope *s) const {
    std::list<daysun::Slope *> out;
    std::map<const daysun::Slope *, unsigned>::const_iterator it = index.find(s);
    if (it == index.end()) return out;
    for (uint32_t e = offsets[it->second]; e < offsets[it->second + 1]; ++e) out.push_back(slope[targets[e]]);
    return out;
  }

  // TwoDmap::CollisionCheck (map2D.h:351-411) on slope-table indices; same control flow, the
  // reference's expressions kept verbatim (including `(a < b) + 2*r`, a bool plus a float, :385).
  bool collides(unsigned i, int n, RobotSphere &robot, std::vector<unsigned> &mark, unsigned &stamp) const {
    daysun::Slope *s = slope[i];
    const float r = robot.getRobotR();
    if (s->up == true) return true;
    std::vector<unsigned> all(1, i), now(1, i), add;
    ++stamp;
    mark[i] = stamp;  // isContainedQ(*, allSlope): identity of (morton_xy, morton_z) = identity of the Slope
    while (n > 0) {
      for (size_t a = 0; a < now.size(); ++a)
        for (uint32_t e = offsets[now[a]]; e < offsets[now[a] + 1]; ++e) {
          const unsigned t = targets[e];
          if (mark[t] != stamp) { mark[t] = stamp; add.push_back(t); all.push_back(t); }
        }
      --n;
      now.swap(add);
      add.clear();
    }
    for (size_t a = 0; a < all.size(); ++a) {
      daysun::Slope *o = slope[all[a]];
      if ((o->mean(2) < s->mean(2)) && o->up == true) return true;
      if ((o->mean(2) > s->mean(2)) && (((o->mean(2) < s->mean(2)) + 2 * r)) && (o->mean(2) - s->mean(2) > robot.getReachableHeight())) return true;
    }
    if (!last_in_cell[i]) {  // the next Slope above in the same Cell (:397-407)
      daysun::Slope *up = slope[i + 1];
      return (up->mean(2) < s->mean(2) + 2 * r) && (up->mean(2) - s->mean(2) > robot.getReachableHeight());
    }
    return false;
  }

  // TwoDmap::computeCost, demand "slope" (map2D.h:1285-1350): the goal-seeded FIFO relaxation
  // of Slope::h, with the graph for the neighbour lists and flags for the list-membership scans
  // (isContainedQ is O(list) per call in the reference).  Same visiting order, same float
  // expressions, hence the same h on every Slope.  Returns the number of traversable Slopes
  // ("traversability slopes", :1387), or -1 when the goal is not on a Slope (:1302).
  long computeCostFast(daysun::TwoDmap &m, octomath::Vector3 goal, RobotSphere &robot) const {
    std::string key;
    int z;
    m.transMortonXYZ(goal, key, z);
    std::map<std::string, daysun::Cell *>::iterator it = m.map_cell.find(key);
    if (it == m.map_cell.end()) return 0;  // the reference falls through with an empty queue
    std::map<int, daysun::Slope *, CmpByKeyUD>::iterator ss = it->second->map_slope.find(z);
    if (ss == it->second->map_slope.end()) return -1;
    const unsigned g = index.find(ss->second)->second;
    slope[g]->h = 0;
    const int n = (ceil(2 * robot.getRobotR() / m.getGridLen()) - 1) / 2;
    std::vector<unsigned char> in_q(slope.size(), 0), in_closed(slope.size(), 0), in_trav(slope.size(), 0);
    std::vector<unsigned> mark(slope.size(), 0);
    unsigned stamp = 0;
    std::deque<unsigned> Q(1, g);
    in_q[g] = 1;
    long n_trav = 0;
    while (!Q.empty()) {
      const unsigned cur = Q.front();
      daysun::Slope *c = slope[cur];
      if (!collides(cur, n, robot, mark, stamp)) {
        for (uint32_t e = offsets[cur]; e < offsets[cur + 1]; ++e) {
          const unsigned t = targets[e];
          daysun::Slope *nb = slope[t];
          if (nb->up == true) {
            nb->h = FLT_MAX;
            in_closed[t] = 1;
          } else if (nb->h > c->h + m.TravelCost(c->mean, nb->mean, goal(2))) {
            nb->h = c->h + m.TravelCost(c->mean, nb->mean, goal(2));
            if (!in_q[t] && !in_closed[t] && !in_trav[t]) { Q.push_back(t); in_q[t] = 1; }
          }
        }
        in_trav[cur] = 1;
        ++n_trav;
      } else {
        c->h = FLT_MAX;
        in_closed[cur] = 1;
      }
      Q.pop_front();
      in_q[cur] = 0;
    }
    return n_trav;
  }
};

}  // namespace gndt_adapter
#endif
