// ref_adapter_check.cpp — TEST INFRASTRUCTURE (oracle/_ref build).  Fills a fresh
// daysun::TwoDmap through adapter/gndt_twodmap_adapter.h and compares every container with
// the map the reference itself built in the last gndt_ref_build call (global `map2D` of
// src/receiver.cpp).  Part of the same translation unit as ref_driver.cpp (the reference
// headers have no include guards for their function definitions).
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
