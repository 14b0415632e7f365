// ref_driver.cpp — ROS-free driver around the reference's OWN sources (oracle/_ref build).
//
// TEST INFRASTRUCTURE ONLY.  This translation unit #includes /root/reference/src/receiver.cpp
// (and through it include/GlobalPlan.h, map2D.h, robot.h, Stopwatch.h, Vector3.h) from
// where they lie, unmodified, against the inert shims in oracle/shims/.  `main` of the
// reference is renamed away; its chatterCallback (src/receiver.cpp:137-176) is invoked
// directly with a shim PointCloud2, so setCloudFirst, the uniformDivision loop,
// create2DMap, isSlope, countRoughNormal, transMortonXYZ and countMorton all run as
// written by the reference.  Only PCL's centroid/scatter and Eigen's eigen-solver are
// restated (shims; third-party code that is not in the reference tree).
//
// Output goes to oracle/_ref/libgndt_ref.so (git-ignored, travels to the GPU box).
// No reference source is copied into this repository.
#include <algorithm>
#include <vector>
#include <sstream>
#include <iostream>

#define main gndt_ref_receiver_main_unused
#include "receiver.cpp"  // found via -I/root/reference/src ; pulls GlobalPlan.h/map2D.h
#undef main

#include "../include/gndt.h"

namespace {

struct RefResult {
  uint64_t n_input, n_binned, n_dropped, n_outside_tile;
  uint64_t n_columns, n_voxels, n_fitted, n_slopes;
  gndt_voxel *voxels;
  gndt_column *columns;
  uint32_t *morton_list;
  double division_s, calculate_s, edges_s;
};

void key_to_signed(const std::string &key, int &sx, int &sy) {
  int a, b;
  mortonToXY(a, b, strToInt(key.substr(1)));  // the reference's own inverse (Stopwatch.h:171)
  char q = key[0];
  sx = (q == 'A' || q == 'B') ? a : -a;
  sy = (q == 'A' || q == 'C') ? b : -b;
}

long contiguous(int s) { return s > 0 ? s - 1 : s; }

double parse_after(const std::string &log, const std::string &tag) {
  size_t p = log.find(tag);
  if (p == std::string::npos) return -1.0;
  return atof(log.c_str() + p + tag.size());
}

void clear_map() {
  // the reference never frees anything (one build per process); release what the last
  // build allocated so the oracle can be called repeatedly from one test process
  for (auto &kv : map2D.map_xy) delete kv.second;
  for (auto &kv : map2D.map_cell) {
    for (auto &s : kv.second->map_slope) delete s.second;
    delete kv.second;
  }
  map2D.map_xy.clear();
  map2D.map_cell.clear();
  map2D.morton_list.clear();
  map2D.changeMorton_list.clear();
}

}  // namespace

extern "C" {

// Same signature/semantics as gndt_oracle_build (mode is ignored: the reference is
// binary32).  Only origin_is_first_point == 1 and untiled builds exist in the reference.
int gndt_ref_build(const float *xyz, size_t n, size_t stride_floats, const gndt_params *P,
                   int /*mode*/, RefResult **out) {
  if (!xyz || !P || !out || n < 2 || stride_floats < 3) return GNDT_ERR_INVALID_ARG;
  if (!P->origin_is_first_point || P->tile_lo < P->tile_hi || P->normalize_cov || P->min_points != MINPOINTSIZE)
    return GNDT_ERR_INVALID_ARG;
  clear_map();

  std::vector<pcl::PointXYZ> pts(n);
  for (size_t i = 0; i < n; ++i)
    pts[i] = pcl::PointXYZ(xyz[i * stride_floats], xyz[i * stride_floats + 1], xyz[i * stride_floats + 2]);
  boost::shared_ptr<sensor_msgs::PointCloud2> msg(new sensor_msgs::PointCloud2);
  msg->data = pts.data();
  msg->n = n;

  // what main() does with the ROS params (src/receiver.cpp:256-269)
  demand = (P->demand == GNDT_DEMAND_TRUE) ? "true" : "slope";
  map2D.setLen(P->grid_len);
  map2D.setZLen(P->z_len);
  map2D.setInterval(P->slope_interval);
  // park pos/goal far outside any map so computeCost / A* (host planner, out of scope)
  // find no goal cell and return immediately
  robot.setPos("30000,30000,30000");
  robot.setGoal("-30000,-30000,-30000");

  std::ostringstream log;
  std::streambuf *old = std::cout.rdbuf(log.rdbuf());
  chatterCallback(msg);  // src/receiver.cpp:137-176, unmodified
  std::cout.rdbuf(old);

  RefResult *R = (RefResult *)calloc(1, sizeof(RefResult));
  R->n_input = n;
  R->division_s = parse_after(log.str(), "division time: ");
  R->calculate_s = parse_after(log.str(), "calculate time: ");

  // ---- flatten the reference's containers into the canonical tables -----------------
  struct Rec { long cx, cy, cz; gndt_voxel v; std::string key; };
  std::vector<Rec> recs;
  recs.reserve(map2D.map_xy.size());
  uint32_t rank = 0;
  std::map<std::string, uint32_t> col_rank;
  uint32_t crank = 0;
  double t_edges = 0;
  for (list<string>::iterator it = map2D.morton_list.begin(); it != map2D.morton_list.end(); ++it) {
    col_rank[*it] = crank++;
    auto range = map2D.map_xy.equal_range(*it);
    Cell *cell = map2D.map_cell.count(*it) ? map2D.map_cell[*it] : NULL;
    for (auto nit = range.first; nit != range.second; ++nit) {
      daysun::OcNode *nd = nit->second;
      Rec r;
      memset(&r.v, 0, sizeof(r.v));
      int sx, sy;
      key_to_signed(nd->morton, sx, sy);
      r.key = nd->morton;
      r.v.sx = sx; r.v.sy = sy; r.v.sz = nd->z;
      r.cx = contiguous(sx); r.cy = contiguous(sy); r.cz = contiguous(nd->z);
      r.v.count = (uint32_t)(nd->N + (int)nd->test_cloud.points.size());
      r.v.first_index = rank++;  // NOT a cloud index: traversal rank (first-seen order)
      R->n_binned += r.v.count;
      if (nd->N >= MINPOINTSIZE) {
        r.v.flags |= GNDT_F_FITTED;
        for (int k = 0; k < 3; ++k) r.v.mean[k] = nd->xyz_centroid(k);
        const Eigen::Matrix3f &C = nd->covariance_matrix;
        r.v.scatter[0] = C(0, 0); r.v.scatter[1] = C(0, 1); r.v.scatter[2] = C(0, 2);
        r.v.scatter[3] = C(1, 1); r.v.scatter[4] = C(1, 2); r.v.scatter[5] = C(2, 2);
        Eigen::EigenSolver<Eigen::Matrix3f> es(C);
        double e[3] = {es.evd[0], es.evd[1], es.evd[2]};
        std::sort(e, e + 3);
        for (int k = 0; k < 3; ++k) r.v.evals[k] = (float)e[k];
        float rough; Eigen::Vector3f nrm;
        nd->countRoughNormal(rough, nrm);  // map2D.h:110-133
        r.v.rough = rough;
        for (int k = 0; k < 3; ++k) r.v.normal[k] = nrm(k);
        R->n_fitted++;
      }
      Slope *s = NULL;
      if (cell) {
        auto sit = cell->map_slope.find(nd->z);
        if (sit != cell->map_slope.end()) s = sit->second;
      }
      if (s) {
        r.v.flags |= GNDT_F_SLOPE;
        if (s->down) r.v.flags |= GNDT_F_DOWN;
        for (int k = 0; k < 3; ++k) { r.v.mean[k] = s->mean(k); r.v.normal[k] = s->normal(k); }
        r.v.rough = s->rough;
        R->n_slopes++;
      } else if ((r.v.flags & GNDT_F_FITTED) && P->demand == GNDT_DEMAND_SLOPE) {
        r.v.flags |= GNDT_F_UP;  // isSlope returned false <=> up (map2D.h:101-104)
      }
      recs.push_back(r);
    }
  }
  // local traversability through the reference's own AccessibleNeighbors (map2D.h:529-548)
  double t0 = stopwatch();
  for (auto &r : recs) {
    if (!(r.v.flags & GNDT_F_SLOPE)) continue;
    Slope *s = map2D.map_cell[r.key]->map_slope[r.v.sz];
    float comand = (P->demand == GNDT_DEMAND_TRUE) ? 4.f : 2.5f;
    list<Slope *> nb = map2D.AccessibleNeighbors(s, robot, comand);
    for (Slope *t : nb) {
      int tx, ty;
      key_to_signed(t->morton_xy, tx, ty);
      if (tx == r.v.sx) r.v.flags |= (contiguous(ty) < contiguous(r.v.sy)) ? GNDT_F_REACH_L : GNDT_F_REACH_R;
      else r.v.flags |= (contiguous(tx) > contiguous(r.v.sx)) ? GNDT_F_REACH_F : GNDT_F_REACH_B;
    }
  }
  if (P->demand == GNDT_DEMAND_TRUE)  // Slope::up as left behind by the lazy countUp calls
    for (auto &r : recs)
      if (r.v.flags & GNDT_F_SLOPE) {
        Slope *s = map2D.map_cell[r.key]->map_slope[r.v.sz];
        if (s->countUp(map2D.map_xy, map2D.getInterval())) r.v.flags |= GNDT_F_UP;
      }
  t_edges = stopwatch() - t0;
  R->edges_s = t_edges;

  std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) {
    if (a.cx != b.cx) return a.cx < b.cx;
    if (a.cy != b.cy) return a.cy < b.cy;
    return a.cz < b.cz;
  });
  R->n_voxels = recs.size();
  R->n_columns = map2D.morton_list.size();
  R->voxels = (gndt_voxel *)calloc(recs.size() + 1, sizeof(gndt_voxel));
  R->columns = (gndt_column *)calloc(R->n_columns + 1, sizeof(gndt_column));
  R->morton_list = (uint32_t *)calloc(R->n_columns + 1, sizeof(uint32_t));
  size_t nc = 0, ns = 0;
  for (size_t i = 0; i < recs.size(); ++i) {
    R->voxels[i] = recs[i].v;
    if (i == 0 || recs[i - 1].key != recs[i].key) {
      R->voxels[i].flags |= GNDT_F_COLUMN_HEAD;
      gndt_column &c = R->columns[nc];
      c.sx = recs[i].v.sx; c.sy = recs[i].v.sy; c.voxel_begin = (uint32_t)i;
      c.first_index = col_rank[recs[i].key];  // rank in morton_list
      c.slope_begin = (uint32_t)ns;
      R->morton_list[c.first_index] = (uint32_t)nc;
      nc++;
    }
    R->columns[nc - 1].voxel_count++;
    R->voxels[i].column = (uint32_t)(nc - 1);
    R->voxels[i].slope = 0xFFFFFFFFu;
    if (recs[i].v.flags & GNDT_F_SLOPE) { R->voxels[i].slope = (uint32_t)ns++; R->columns[nc - 1].slope_count++; }
  }
  *out = R;
  return GNDT_OK;
}

void gndt_ref_free(RefResult *R) {
  if (!R) return;
  free(R->voxels); free(R->columns); free(R->morton_list);
  free(R);
}

// the reference's key helpers, exported for known-answer tests
int gndt_ref_count_morton(int a, int b, char *buf) {
  std::string s = countMorton(a, b);  // Stopwatch.h:116-147
  strcpy(buf, s.c_str());
  return (int)s.size();
}
void gndt_ref_morton_to_xy(int morton, int *a, int *b) { mortonToXY(*a, *b, morton); }
int gndt_ref_trans_morton_xyz(const float origin[3], float grid_len, float z_len,
                              const float pos[3], char *key, int *sz) {
  daysun::TwoDmap m(grid_len, z_len);
  m.setCloudFirst(Vector3(origin[0], origin[1], origin[2]));
  std::string k;
  m.transMortonXYZ(Vector3(pos[0], pos[1], pos[2]), k, *sz);  // map2D.h:950-976
  strcpy(key, k.c_str());
  return (int)k.size();
}

}  // extern "C"
#include "ref_adapter_check.cpp"
