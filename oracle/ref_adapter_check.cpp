// ref_adapter_check.cpp — TEST INFRASTRUCTURE (oracle/_ref build).  Fills a fresh
// daysun::TwoDmap through adapter/gndt_twodmap_adapter.h and compares every container with
// the map the reference itself built in the last gndt_ref_build call (global `map2D` of
// src/receiver.cpp).  Part of the same translation unit as ref_driver.cpp (the reference
// headers have no include guards for their function definitions).
#include <time.h>
#include "../adapter/gndt_twodmap_adapter.h"

extern "C" int gndt_ref_adapter_check(const float *origin, const gndt_params *P, const gndt_voxel *vox, size_t nv,
                                      const gndt_slope *sl, size_t ns, const gndt_column *cols, size_t nc,
                                      int compare_floats) {
  daysun::TwoDmap m(P->grid_len, P->z_len);
  m.setInterval(P->slope_interval);
  gndt_adapter::fill_twodmap(m, origin, vox, nv, sl, ns, cols, nc, true);
  int bad = 0;
  // morton_list: same keys, same (first-seen) order
  if (m.morton_list.size() != map2D.morton_list.size()) return -1;
  if (!std::equal(m.morton_list.begin(), m.morton_list.end(), map2D.morton_list.begin())) bad += 1;
  // map_cell / map_slope
  if (m.map_cell.size() != map2D.map_cell.size()) return -2;
  std::map<std::string, Cell *>::iterator a = m.map_cell.begin(), b = map2D.map_cell.begin();
  for (; a != m.map_cell.end(); ++a, ++b) {
    if (a->first != b->first || a->second->getMorton() != b->second->getMorton()) { bad += 1000; continue; }
    if (a->second->map_slope.size() != b->second->map_slope.size()) { bad += 1000; continue; }
    std::map<int, Slope *, CmpByKeyUD>::iterator sa = a->second->map_slope.begin(), sb = b->second->map_slope.begin();
    for (; sa != a->second->map_slope.end(); ++sa, ++sb) {
      Slope *x = sa->second, *y = sb->second;
      if (sa->first != sb->first || x->morton_xy != y->morton_xy || x->morton_z != y->morton_z || x->down != y->down ||
          x->up != y->up || x->h != y->h || x->father != y->father)
        bad += 1000;
      if (compare_floats)
        for (int k = 0; k < 3; ++k)
          if (x->mean(k) != y->mean(k) || x->normal(k) != y->normal(k) || x->rough != y->rough) { bad += 1; break; }
    }
  }
  // map_xy: same keys and, inside each column, the same node order
  if (m.map_xy.size() != map2D.map_xy.size()) return -3;
  std::multimap<std::string, OcNode *>::iterator na = m.map_xy.begin(), nb = map2D.map_xy.begin();
  for (; na != m.map_xy.end(); ++na, ++nb) {
    if (na->first != nb->first || na->second->z != nb->second->z || na->second->N != nb->second->N) bad += 1000;
    if (compare_floats && !(na->second->xyz_centroid == nb->second->xyz_centroid)) bad += 1;
  }
  // the planner's entry points work on the adapter-built map: transMortonXYZ + lookup
  std::string key; int z;
  m.transMortonXYZ(octomath::Vector3(origin[0] + 1.234f, origin[1] - 0.77f, origin[2]), key, z);
  std::string key2; int z2;
  map2D.transMortonXYZ(octomath::Vector3(origin[0] + 1.234f, origin[1] - 0.77f, origin[2]), key2, z2);
  if (key != key2 || z != z2) bad += 1000;
  return bad;
}

// ---- integer-keyed lookups (adapter CellIndex / include/gndt_lookup.h) against the reference's
// own string path: map_cell.find(key) for every cell and a ring of empty cells around it, and
// the four neighbour keys produced by the reference's private TwoDmap::countLRFB (reached
// through an explicit-instantiation accessor: access control is not checked there).
namespace {
typedef void (daysun::TwoDmap::*LrfbFn)(std::string, int, int, std::string &, std::string &, std::string &, std::string &);
template <typename Tag, LrfbFn M> struct PrivateAccess { friend LrfbFn gndt_get(Tag) { return M; } };
struct LrfbTag { friend LrfbFn gndt_get(LrfbTag); };
template struct PrivateAccess<LrfbTag, &daysun::TwoDmap::countLRFB>;
double now_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
}  // namespace

// Returns the number of mismatches; rates[0] = string lookups/s, rates[1] = integer lookups/s
// over the same sequence of neighbour queries (4 per cell).
extern "C" int gndt_ref_lookup_check(const float *origin, const gndt_params *P, const gndt_voxel *vox, size_t nv,
                                     const gndt_slope *sl, size_t ns, const gndt_column *cols, size_t nc, double *rates) {
  daysun::TwoDmap m(P->grid_len, P->z_len);
  m.setInterval(P->slope_interval);
  gndt_adapter::fill_twodmap(m, origin, vox, nv, sl, ns, cols, nc, false);
  gndt_adapter::CellIndex index(m, cols, nc);
  const LrfbFn lrfb = gndt_get(LrfbTag());
  int bad = 0;
  for (size_t i = 0; i < nc; ++i) {
    const int sx = cols[i].sx, sy = cols[i].sy;
    std::map<std::string, Cell *>::iterator it = m.map_cell.find(gndt_adapter::morton_key(sx, sy));
    if (it == m.map_cell.end() || index.find(sx, sy) != it->second) ++bad;
    for (int dx = -2; dx <= 2; ++dx)
      for (int dy = -2; dy <= 2; ++dy) {  // present and absent cells around it (skipping index 0)
        const int tx = sx + dx, ty = sy + dy;
        if (tx == 0 || ty == 0) { if (index.find(tx, ty) != NULL) ++bad; continue; }
        std::map<std::string, Cell *>::iterator jt = m.map_cell.find(gndt_adapter::morton_key(tx, ty));
        if (index.find(tx, ty) != (jt == m.map_cell.end() ? (Cell *)NULL : jt->second)) ++bad;
      }
    const std::string q(1, sx > 0 ? (sy > 0 ? 'A' : 'B') : (sy > 0 ? 'C' : 'D'));
    std::string k[4];
    (m.*lrfb)(q, std::abs(sx), std::abs(sy), k[0], k[1], k[2], k[3]);
    for (int d = 0; d < 4; ++d) {
      std::map<std::string, Cell *>::iterator jt = m.map_cell.find(k[d]);
      if (index.neighbor(sx, sy, d) != (jt == m.map_cell.end() ? (Cell *)NULL : jt->second)) ++bad;
    }
    // slopes of the cell: map_slope.find(z) for every layer present and one absent
    for (unsigned s = cols[i].slope_begin; s < cols[i].slope_begin + cols[i].slope_count; ++s) {
      if (gndtl_find_slope(cols, nc, sl, sx, sy, sl[s].sz) != (int64_t)s) ++bad;
      if (index.find_slope(sx, sy, sl[s].sz) == NULL) ++bad;
    }
    if (gndtl_find_slope(cols, nc, sl, sx, sy, 32760) != -1) ++bad;
  }
  if (rates) {  // the planner's inner loop: 4 neighbour cells per expansion
    size_t hits = 0;
    double t0 = now_s();
    for (size_t i = 0; i < nc; ++i) {
      const int sx = cols[i].sx, sy = cols[i].sy;
      const std::string q(1, sx > 0 ? (sy > 0 ? 'A' : 'B') : (sy > 0 ? 'C' : 'D'));
      std::string k[4];
      (m.*lrfb)(q, std::abs(sx), std::abs(sy), k[0], k[1], k[2], k[3]);
      for (int d = 0; d < 4; ++d) hits += m.map_cell.find(k[d]) != m.map_cell.end();
    }
    double t1 = now_s();
    size_t hits2 = 0;
    for (size_t i = 0; i < nc; ++i)
      for (int d = 0; d < 4; ++d) hits2 += index.neighbor(cols[i].sx, cols[i].sy, d) != NULL;
    double t2 = now_s();
    if (hits != hits2) ++bad;
    rates[0] = 4.0 * nc / (t1 - t0);
    rates[1] = 4.0 * nc / (t2 - t1);
  }
  return bad;
}

// ---- traversability graph (gndt_build_edges / oracle CSR) against the reference's own
// AccessibleNeighbors for EVERY Slope (same Slopes, same order), and TwoDmap::computeCost against
// the adapter's computeCostFast on two identical maps (same h on every Slope).
//   out[0] reference AccessibleNeighbors calls/s   out[1] graph calls/s
//   out[2] reference computeCost seconds           out[3] computeCostFast seconds
//   out[4] Slopes with finite h                    out[5] traversable Slopes (computeCostFast's return)
extern "C" int gndt_ref_graph_check(const float *origin, const gndt_params *P, const gndt_voxel *vox, size_t nv,
                                    const gndt_slope *sl, size_t ns, const gndt_column *cols, size_t nc,
                                    const uint32_t *off, const uint32_t *tgt, const float *goal, double *out) {
  daysun::TwoDmap m1(P->grid_len, P->z_len), m2(P->grid_len, P->z_len);
  m1.setInterval(P->slope_interval);
  m2.setInterval(P->slope_interval);
  gndt_adapter::fill_twodmap(m1, origin, vox, nv, sl, ns, cols, nc, false);
  gndt_adapter::fill_twodmap(m2, origin, vox, nv, sl, ns, cols, nc, false);
  gndt_adapter::SlopeGraph g1(m1, cols, nc, ns, off, tgt), g2(m2, cols, nc, ns, off, tgt);
  int bad = 0;
  for (size_t i = 0; i < ns; ++i) {
    if (!g1.slope[i]) { ++bad; continue; }
    std::list<Slope *> a = m1.AccessibleNeighbors(g1.slope[i], robot, 2.5f), b = g1.AccessibleNeighborsFast(g1.slope[i]);
    if (a.size() != b.size() || !std::equal(a.begin(), a.end(), b.begin())) ++bad;
  }
  if (out) {
    size_t n1 = 0, n2 = 0;
    double t0 = now_s();
    for (size_t i = 0; i < ns; ++i) n1 += m1.AccessibleNeighbors(g1.slope[i], robot, 2.5f).size();
    double t1 = now_s();
    for (size_t i = 0; i < ns; ++i) n2 += g1.AccessibleNeighborsFast(g1.slope[i]).size();
    double t2 = now_s();
    if (n1 != n2) ++bad;
    out[0] = ns / (t1 - t0);
    out[1] = ns / (t2 - t1);
  }
  if (goal) {
    ros::Publisher pub, pub2;
    const octomath::Vector3 gpos(goal[0], goal[1], goal[2]);
    double t0 = now_s();
    m1.computeCost(gpos, robot, pub, pub2, "slope");  // the reference's own cost map
    double t1 = now_s();
    const long n_trav = g2.computeCostFast(m2, gpos, robot);
    double t2 = now_s();
    size_t finite = 0;
    for (size_t i = 0; i < ns; ++i) {
      if (g1.slope[i]->h != g2.slope[i]->h) ++bad;
      finite += g1.slope[i]->h < FLT_MAX;
    }
    if (out) { out[2] = t1 - t0; out[3] = t2 - t1; out[4] = (double)finite; out[5] = (double)n_trav; }
  }
  return bad;
}
