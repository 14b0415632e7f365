/*
 * gndt_oracle.c — CPU ORACLE for the grid-NDT map-construction path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the algorithm of
 * daysun/grid_ndt's map-construction path, used solely as the checker for the CUDA
 * implementation (tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
 * --impl reference legs).  Nothing under grid_ndt_b200/ may import, link or call it.
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference tree).  It is a restatement with integer keys and flat arrays, not a copy:
 * the reference keys everything by decimal-Morton strings in std::multimap.
 *
 * Parity pinning: the reference ships no tests/golden vectors.  This
 * restatement is pinned (tests/test_oracle_*.py) against
 *   (i)  the reference's own headers compiled here against third-party shims
 *        (oracle/_ref/libgndt_ref.so, built by oracle/Makefile from the sources where
 *        they lie under /root/reference) — control flow, container ordering, key
 *        strings and all integer outputs bit-exact;
 *   (ii) committed golden vectors produced by that library (tests/golden/).
 * The floating-point kernels of PCL (centroid/scatter) and Eigen (EigenSolver) are
 * third-party code that is NOT in the reference tree and has no pinned version
 * (package.xml:45-46): their published algorithms are restated below and for them the
 * status is "parity unpinned" (see DESIGN.md §Oracle).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math: every float
 * operation below must be a true IEEE binary32 operation).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/gndt.h"

#include <float.h>
#if defined(FLT_EVAL_METHOD) && FLT_EVAL_METHOD != 0
#error "oracle needs true binary32 evaluation (FLT_EVAL_METHOD == 0)"
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct oracle_result {
  uint64_t n_input, n_binned, n_dropped, n_outside_tile;
  uint64_t n_columns, n_voxels, n_fitted, n_slopes;
  gndt_voxel *voxels;   /* canonical order: ascending (cx,cy,cz)                       */
  gndt_column *columns; /* same order as the voxel table                               */
  uint32_t *morton_list; /* column ids (index into columns[]) in first-seen order      */
  double division_s;    /* stage 1 wall time ("division time", receiver.cpp:157)        */
  double calculate_s;   /* stage 2 wall time ("calculate time", receiver.cpp:162)       */
  double edges_s;
} oracle_result;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------
 * Key arithmetic
 * ---------------------------------------------------------------------------------- */

/* One axis of TwoDmap::transMortonXYZ (include/map2D.h:963-970): all binary32.
 * n = (int)ceil(float(abs(p - p0)/len)), 0 -> 1; sign: p > p0 ? + : -.
 * Returns 0 when the result is outside the supported +-GNDT_MAX_INDEX range
 * (Stopwatch.h:102-110 wraps there) or not finite. */
static int axis_index(float p, float p0, float len, int32_t *s) {
  float d = p - p0;
  float a = fabsf(d);
  float q = a / len;
  float c = ceilf(q);
  if (!(c <= (float)GNDT_MAX_INDEX)) return 0; /* also catches NaN */
  int32_t n = (int32_t)c;
  if (n == 0) n = 1;
  *s = (p > p0) ? n : -n;
  return 1;
}

int gndt_oracle_trans_morton_xyz(const float origin[3], float grid_len, float z_len,
                                 const float pos[3], int32_t *sx, int32_t *sy, int32_t *sz) {
  /* map2D.h:950-973.  Quadrant letters (:952-962) are the sign pair of (sx,sy). */
  if (!isfinite(pos[0]) || !isfinite(pos[1]) || !isfinite(pos[2])) return -1;
  if (!axis_index(pos[0], origin[0], grid_len, sx)) return -1;
  if (!axis_index(pos[1], origin[1], grid_len, sy)) return -1;
  if (!axis_index(pos[2], origin[2], z_len, sz)) return -1;
  return 0;
}

/* countMorton (include/Stopwatch.h:116-147), restated on integers but keeping its
 * observable overflow behaviour: binToDec (:102-110) accumulates in `unsigned int`
 * (high bits fall off) and returns `(int)`.  Returned as int32. */
int32_t gndt_oracle_count_morton(int a, int b) {
  /* decToBinStr2 (:39-47) yields "" for 0 and loops `for(a=n; a; a/=2)`; negative inputs
   * are outside the domain here (indices are >= 1). */
  int la = 0, lb = 0;
  for (int t = a; t; t /= 2) la++;
  for (int t = b; t; t /= 2) lb++;
  int L = la > lb ? la : lb; /* zero-padded to equal length (:119-130) */
  unsigned int acc = 0;      /* binToDec's accumulator                  */
  for (int k = L - 1; k >= 0; --k) { /* result string is MSB first: line[k'],column[k'] */
    acc <<= 1;
    acc |= (unsigned)((a >> k) & 1); /* `line`   bit -> odd position  */
    acc <<= 1;
    acc |= (unsigned)((b >> k) & 1); /* `column` bit -> even position */
  }
  return (int32_t)acc;
}

/* mortonToXY (include/Stopwatch.h:171-189) */
void gndt_oracle_morton_to_xy(int morton, int *a, int *b) {
  int len = 0;
  for (int t = morton; t; t /= 2) len++;
  if (len % 2) len++;
  unsigned ua = 0, ub = 0;
  for (int k = len - 1; k >= 0; k -= 2) {
    ua = (ua << 1) | (unsigned)((morton >> k) & 1);
    ub = (ub << 1) | (unsigned)((morton >> (k - 1)) & 1);
  }
  *a = (int)ua;
  *b = (int)ub;
}

/* "A"+countMorton(...) (map2D.h:971-972) */
int gndt_oracle_morton_string(int32_t sx, int32_t sy, char *buf) {
  char q = sx > 0 ? (sy > 0 ? 'A' : 'B') : (sy > 0 ? 'C' : 'D');
  return sprintf(buf, "%c%d", q, (int)gndt_oracle_count_morton(abs(sx), abs(sy)));
}

/* ------------------------------------------------------------------------------------
 * Binning store (uniformDivision, src/receiver.cpp:41-93) on integer keys
 * ---------------------------------------------------------------------------------- */

typedef struct onode {  /* one OcNode (map2D.h:38-57) */
  int32_t sx, sy, sz;
  uint32_t count;       /* test_cloud.points.size() before create2DMap               */
  uint32_t first, head, tail; /* point chain (insertion order)                       */
  int32_t next_in_col;  /* multimap equal-range order = insertion order              */
  int32_t col;
  int N;                /* OcNode::N                                                  */
  float cen[3];         /* OcNode::xyz_centroid (zero-initialised, :55)               */
  float cov[6];
  double evals[3];
  double nrm[3];
  float rough;
  uint32_t flags;
} onode;

typedef struct ocol {   /* one xy key of map_xy / one Cell */
  int32_t sx, sy;
  int32_t first_node, last_node;
  uint32_t n_nodes;
} ocol;

typedef struct store {
  onode *nodes; size_t n_nodes, cap_nodes;
  ocol *cols;   size_t n_cols, cap_cols;
  int32_t *slots; size_t n_slots; /* open addressing on (sx,sy) -> column id */
  uint32_t *next_pt;
} store;

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
static uint64_t xykey(int32_t sx, int32_t sy) {
  return ((uint64_t)(uint32_t)sx << 32) | (uint32_t)sy;
}

static void store_rehash(store *st, size_t n_slots) {
  int32_t *s = (int32_t *)malloc(n_slots * sizeof(int32_t));
  for (size_t i = 0; i < n_slots; ++i) s[i] = -1;
  for (size_t c = 0; c < st->n_cols; ++c) {
    size_t h = mix64(xykey(st->cols[c].sx, st->cols[c].sy)) & (n_slots - 1);
    while (s[h] >= 0) h = (h + 1) & (n_slots - 1);
    s[h] = (int32_t)c;
  }
  free(st->slots);
  st->slots = s;
  st->n_slots = n_slots;
}

static int32_t store_find_col(const store *st, int32_t sx, int32_t sy) {
  size_t h = mix64(xykey(sx, sy)) & (st->n_slots - 1);
  while (st->slots[h] >= 0) {
    const ocol *c = &st->cols[st->slots[h]];
    if (c->sx == sx && c->sy == sy) return st->slots[h];
    h = (h + 1) & (st->n_slots - 1);
  }
  return -1;
}

static int32_t store_new_node(store *st, int32_t col, int32_t sz, uint32_t pt) {
  if (st->n_nodes == st->cap_nodes) {
    st->cap_nodes = st->cap_nodes ? st->cap_nodes * 2 : 1024;
    st->nodes = (onode *)realloc(st->nodes, st->cap_nodes * sizeof(onode));
  }
  onode *nd = &st->nodes[st->n_nodes];
  memset(nd, 0, sizeof(*nd)); /* OcNode ctor: zero covariance, zero centroid, N = 0 (:52-57) */
  nd->sx = st->cols[col].sx; nd->sy = st->cols[col].sy; nd->sz = sz;
  nd->count = 1; nd->first = nd->head = nd->tail = pt;
  nd->next_in_col = -1; nd->col = col;
  return (int32_t)st->n_nodes++;
}

/* uniformDivision(temp,false) (receiver.cpp:41-93) for point index `pt`. */
static void store_insert(store *st, int32_t sx, int32_t sy, int32_t sz, uint32_t pt) {
  st->next_pt[pt] = UINT32_MAX;
  int32_t c = store_find_col(st, sx, sy);
  if (c < 0) { /* map_xy.count(morton_xy)==0 (:60): new node, morton_list.push_back (:70) */
    if (st->n_cols == st->cap_cols) {
      st->cap_cols = st->cap_cols ? st->cap_cols * 2 : 1024;
      st->cols = (ocol *)realloc(st->cols, st->cap_cols * sizeof(ocol));
    }
    if ((st->n_cols + 1) * 2 > st->n_slots) store_rehash(st, st->n_slots * 2);
    c = (int32_t)st->n_cols++;
    st->cols[c].sx = sx; st->cols[c].sy = sy; st->cols[c].n_nodes = 1;
    size_t h = mix64(xykey(sx, sy)) & (st->n_slots - 1);
    while (st->slots[h] >= 0) h = (h + 1) & (st->n_slots - 1);
    st->slots[h] = c;
    int32_t nd = store_new_node(st, c, sz, pt);
    st->cols[c].first_node = st->cols[c].last_node = nd;
    return;
  }
  /* walk the equal range in insertion order looking for the same z (:72-83) */
  for (int32_t nd = st->cols[c].first_node; nd >= 0; nd = st->nodes[nd].next_in_col) {
    if (st->nodes[nd].sz == sz) {
      st->next_pt[st->nodes[nd].tail] = pt;
      st->nodes[nd].tail = pt;
      st->nodes[nd].count++;
      return;
    }
  }
  /* not found: new node appended to the column's range (:84-91) */
  int32_t nd = store_new_node(st, c, sz, pt);
  st->nodes[st->cols[c].last_node].next_in_col = nd;
  st->cols[c].last_node = nd;
  st->cols[c].n_nodes++;
}

/* ------------------------------------------------------------------------------------
 * Third-party numerics restated
 * ---------------------------------------------------------------------------------- */

/* pcl::compute3DCentroid (dense branch) + pcl::computeCovarianceMatrix(cloud, centroid,
 * Matrix3f) as called at map2D.h:621-622.  [3P: PCL common/impl/centroid.hpp, 1.7/1.8
 * era, version not pinned by the reference.]  binary32, sequential in point order;
 * the covariance is the UN-normalised scatter. */
static void fit_faithful32(const float *xyz, size_t sf, const store *st, onode *nd) {
  float cx = 0.f, cy = 0.f, cz = 0.f;
  for (uint32_t p = nd->head; p != UINT32_MAX; p = st->next_pt[p]) {
    cx = cx + xyz[p * sf + 0];
    cy = cy + xyz[p * sf + 1];
    cz = cz + xyz[p * sf + 2];
  }
  float n = (float)nd->count;
  cx = cx / n; cy = cy / n; cz = cz / n;
  float xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
  for (uint32_t p = nd->head; p != UINT32_MAX; p = st->next_pt[p]) {
    float px = xyz[p * sf + 0] - cx;
    float py = xyz[p * sf + 1] - cy;
    float pz = xyz[p * sf + 2] - cz;
    yy = yy + py * py;
    yz = yz + py * pz;
    zz = zz + pz * pz;
    float ax = px * px, ay = py * px, az = pz * px; /* pt *= pt.x() */
    xx = xx + ax;
    xy = xy + ay;
    xz = xz + az;
  }
  nd->cen[0] = cx; nd->cen[1] = cy; nd->cen[2] = cz;
  nd->cov[0] = xx; nd->cov[1] = xy; nd->cov[2] = xz; nd->cov[3] = yy; nd->cov[4] = yz; nd->cov[5] = zz;
}

/* The same quantities in binary64 two-pass (exact mean), rounded to binary32 once. */
static void fit_truth64(const float *xyz, size_t sf, const store *st, onode *nd) {
  double c[3] = {0, 0, 0};
  for (uint32_t p = nd->head; p != UINT32_MAX; p = st->next_pt[p])
    for (int k = 0; k < 3; ++k) c[k] += (double)xyz[p * sf + k];
  for (int k = 0; k < 3; ++k) c[k] /= (double)nd->count;
  /* one correction pass makes the mean exact to ~1 ulp(double) even for huge voxels */
  double r[3] = {0, 0, 0};
  for (uint32_t p = nd->head; p != UINT32_MAX; p = st->next_pt[p])
    for (int k = 0; k < 3; ++k) r[k] += (double)xyz[p * sf + k] - c[k];
  for (int k = 0; k < 3; ++k) c[k] += r[k] / (double)nd->count;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (uint32_t p = nd->head; p != UINT32_MAX; p = st->next_pt[p]) {
    double dx = (double)xyz[p * sf + 0] - c[0];
    double dy = (double)xyz[p * sf + 1] - c[1];
    double dz = (double)xyz[p * sf + 2] - c[2];
    s[0] += dx * dx; s[1] += dx * dy; s[2] += dx * dz;
    s[3] += dy * dy; s[4] += dy * dz; s[5] += dz * dz;
  }
  for (int k = 0; k < 3; ++k) nd->cen[k] = (float)c[k];
  for (int k = 0; k < 6; ++k) nd->cov[k] = (float)s[k];
}

/* Symmetric 3x3 eigen-decomposition by cyclic Jacobi in binary64.
 * Stands in for Eigen::EigenSolver<Matrix3f> at map2D.h:111-113 [3P: Eigen, version not
 * pinned; its fp32 Hessenberg+QR iteration cannot be reproduced bit-for-bit without the
 * library].  For a symmetric PSD input both return the same real eigenpairs up to fp32
 * noise; eigenvectors are unit length, sign arbitrary.  Zero rows/columns stay exactly
 * zero (rotations with a zero pivot are skipped), so rank-deficient axis-aligned
 * patches yield an exact 0 eigenvalue like the reference's `roughness == 0` test expects
 * (map2D.h:131). */
static void jacobi3(const float cov[6], double w[3], double V[3][3]) {
  double A[3][3] = {{cov[0], cov[1], cov[2]}, {cov[1], cov[3], cov[4]}, {cov[2], cov[4], cov[5]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        double app = A[p][p], aqq = A[q][q], apq = A[p][q];
        A[p][p] = app - t * apq;
        A[q][q] = aqq + t * apq;
        A[p][q] = A[q][p] = 0.0;
        int r = 3 - p - q;
        double arp = A[r][p], arq = A[r][q];
        A[r][p] = A[p][r] = c * arp - s * arq;
        A[r][q] = A[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
}

/* OcNode::countRoughNormal (map2D.h:110-133): smallest eigenvalue (strict `<` chain of
 * :114-130 picks the first minimum in the order the chain visits them) -> rough, its
 * eigenvector -> normal; rough == 0 -> 0.01 (:131-132). */
static void count_rough_normal(onode *nd) {
  double w[3], V[3][3];
  jacobi3(nd->cov, w, V);
  /* the reference compares the solver's (unordered) diagonal as binary32 values */
  float e0 = (float)w[0], e1 = (float)w[1], e2 = (float)w[2];
  int k;
  if (e0 < e1) k = (e0 < e2) ? 0 : 2; else k = (e1 < e2) ? 1 : 2;
  float rough = (float)w[k];
  if (rough == 0) rough = 0.01f;
  nd->rough = rough;
  for (int i = 0; i < 3; ++i) nd->nrm[i] = V[i][k];
  /* ascending eigenvalues for the record */
  double s[3] = {w[0], w[1], w[2]};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (s[j] > s[j + 1]) { double t = s[j]; s[j] = s[j + 1]; s[j + 1] = t; }
  for (int i = 0; i < 3; ++i) nd->evals[i] = s[i];
}

/* OcNode::isSlope (map2D.h:66-108).  Reads the CURRENT centroids of the column's nodes:
 * real for nodes already fitted, the constructor's zeros otherwise (order dependence,
 * SURVEY Q8). */
static int is_slope(const store *st, const onode *nd, int min_points, float interval, int *up,
                    int *down) {
  if (nd->N < min_points) return 0; /* :67 */
  int zadd = nd->sz + 1, zminus = nd->sz - 1; /* :69-75 */
  if (nd->sz == -1) zadd = 1; else if (nd->sz == 1) zminus = -1;
  for (int32_t u = st->cols[nd->col].first_node; u >= 0; u = st->nodes[u].next_in_col) {
    const onode *o = &st->nodes[u];
    float dz = o->cen[2] - nd->cen[2];
    if (o->sz == zminus && fabsf(dz) > interval) *down = 1; /* :85-89  */
    if (o->sz == zadd && fabsf(dz) > interval) *up = 1;     /* :94-98  */
  }
  return !*up; /* :101-107 */
}

/* Slope::countUp (map2D.h:147-177): evaluated lazily, i.e. on FINAL centroids. */
static int count_up(const store *st, const onode *nd, float interval) {
  int zadd = nd->sz + 1;
  if (nd->sz == -1) zadd = 1;
  for (int32_t u = st->cols[nd->col].first_node; u >= 0; u = st->nodes[u].next_in_col) {
    const onode *o = &st->nodes[u];
    float dz = o->cen[2] - nd->cen[2]; /* Slope::mean(2) == centroid z */
    if (o->sz == zadd && fabsf(dz) > interval) return 1;
  }
  return 0;
}

/* TwoDmap::countAngle (map2D.h:477-482), C++11 overloads: dot in binary32 (Eigen
 * Vector3f::dot), pow(float,int) -> double, res stored to float, acos(float) -> float,
 * *180 float, /M_PI double, stored to float; folded to <= 90.  NaN when res > 1. */
static float count_angle(const float n1[3], const float n2[3]) {
  float dot = n1[0] * n2[0];
  dot = dot + n1[1] * n2[1];
  dot = dot + n1[2] * n2[2];
  double l1 = sqrt((double)n1[0] * n1[0] + (double)n1[1] * n1[1] + (double)n1[2] * n1[2]);
  double l2 = sqrt((double)n2[0] * n2[0] + (double)n2[1] * n2[1] + (double)n2[2] * n2[2]);
  float res = (float)((double)dot / (l1 * l2));
  float a = acosf(res) * 180;
  float an = (float)((double)a / M_PI);
  if (an > 90) an = 180 - an;
  return an;
}

static int32_t signed_next(int32_t s) { return s == -1 ? 1 : s + 1; }  /* countLRFB index-1 */
static int32_t signed_prev(int32_t s) { return s == 1 ? -1 : s - 1; }  /* quadrant crossing */

/* countLRFB (map2D.h:197-263) + countReachable (comand 2.5 / 4 branches, :266-296):
 * direction bit set iff the neighbour Cell holds >= 1 Slope passing all four tests. */
static uint32_t reach_bits(const store *st, const onode *c, const gndt_params *P) {
  uint32_t bits = 0;
  const int32_t nbx[4] = {c->sx, c->sx, signed_next(c->sx), signed_prev(c->sx)};
  const int32_t nby[4] = {signed_prev(c->sy), signed_next(c->sy), c->sy, c->sy};
  const uint32_t bit[4] = {GNDT_F_REACH_L, GNDT_F_REACH_R, GNDT_F_REACH_F, GNDT_F_REACH_B};
  float cn[3] = {(float)c->nrm[0], (float)c->nrm[1], (float)c->nrm[2]};
  for (int d = 0; d < 4; ++d) {
    if (abs(nbx[d]) > GNDT_MAX_INDEX || abs(nby[d]) > GNDT_MAX_INDEX) continue;
    int32_t col = store_find_col(st, nbx[d], nby[d]);
    if (col < 0) continue; /* map_cell.find == end (:270) */
    for (int32_t u = st->cols[col].first_node; u >= 0; u = st->nodes[u].next_in_col) {
      const onode *s = &st->nodes[u];
      if (!(s->flags & GNDT_F_SLOPE)) continue;
      if (s->flags & GNDT_F_UP) continue;          /* sit->second->up != true (:276,284) */
      if (!(s->rough <= P->rough_max)) continue;   /* :277,285 */
      float sn[3] = {(float)s->nrm[0], (float)s->nrm[1], (float)s->nrm[2]};
      if (!(count_angle(sn, cn) <= P->angle_max_deg)) continue; /* :278,286 */
      float dz = s->cen[2] - c->cen[2];
      if (!(fabsf(dz) <= P->reach_height)) continue;            /* :279,287 */
      bits |= bit[d];
      break;
    }
  }
  return bits;
}

/* ------------------------------------------------------------------------------------
 * Driver: chatterCallback's two hot loops (src/receiver.cpp:145-162)
 * ---------------------------------------------------------------------------------- */

static int cmp_node_canonical(const void *a, const void *b) {
  const onode *x = *(const onode *const *)a, *y = *(const onode *const *)b;
  int32_t ax = x->sx > 0 ? x->sx - 1 : x->sx, bx = y->sx > 0 ? y->sx - 1 : y->sx;
  if (ax != bx) return ax < bx ? -1 : 1;
  int32_t ay = x->sy > 0 ? x->sy - 1 : x->sy, by = y->sy > 0 ? y->sy - 1 : y->sy;
  if (ay != by) return ay < by ? -1 : 1;
  int32_t az = x->sz > 0 ? x->sz - 1 : x->sz, bz = y->sz > 0 ? y->sz - 1 : y->sz;
  return az < bz ? -1 : (az > bz);
}

/* mode: 0 = faithful32 (the reference's arithmetic), 1 = truth64.
 * xyz: n records of stride_floats floats (x,y,z first). */
int gndt_oracle_build(const float *xyz, size_t n, size_t stride_floats, const gndt_params *P,
                      int mode, oracle_result **out) {
  if (!xyz || !P || !out || n == 0 || stride_floats < 3) return GNDT_ERR_INVALID_ARG;
  oracle_result *R = (oracle_result *)calloc(1, sizeof(*R));
  store st;
  memset(&st, 0, sizeof(st));
  st.n_slots = 1024;
  st.slots = (int32_t *)malloc(st.n_slots * sizeof(int32_t));
  for (size_t i = 0; i < st.n_slots; ++i) st.slots[i] = -1;
  st.next_pt = (uint32_t *)malloc(n * sizeof(uint32_t));

  /* setCloudFirst(points[0]) (receiver.cpp:145); loop from i = 1 (:150) */
  float origin[3];
  size_t start = 0;
  if (P->origin_is_first_point) {
    origin[0] = xyz[0]; origin[1] = xyz[1]; origin[2] = xyz[2];
    start = 1;
  } else {
    origin[0] = P->origin[0]; origin[1] = P->origin[1]; origin[2] = P->origin[2];
  }
  const int tiled = P->tile_lo < P->tile_hi;
  R->n_input = n;

  double t0 = now_s();
  for (size_t i = start; i < n; ++i) {
    int32_t sx, sy, sz;
    if (gndt_oracle_trans_morton_xyz(origin, P->grid_len, P->z_len, &xyz[i * stride_floats], &sx,
                                     &sy, &sz)) {
      R->n_dropped++;
      continue;
    }
    if (tiled) {
      int32_t cx = sx > 0 ? sx - 1 : sx;
      if (cx < P->tile_lo || cx >= P->tile_hi) { R->n_outside_tile++; continue; }
    }
    store_insert(&st, sx, sy, sz, (uint32_t)i);
    R->n_binned++;
  }
  double t1 = now_s();

  /* create2DMap (map2D.h:592-668): columns in morton_list (first-seen) order, nodes of a
   * column in insertion order */
  for (size_t c = 0; c < st.n_cols; ++c) {
    for (int32_t u = st.cols[c].first_node; u >= 0; u = st.nodes[u].next_in_col) {
      onode *nd = &st.nodes[u];
      if ((int)nd->count < P->min_points) continue; /* :611 */
      if (mode == 0) fit_faithful32(xyz, stride_floats, &st, nd);
      else fit_truth64(xyz, stride_floats, &st, nd);
      if (P->normalize_cov) {
        float nn = (float)nd->count;
        for (int k = 0; k < 6; ++k) nd->cov[k] = nd->cov[k] / nn;
      }
      nd->N += (int)nd->count; /* :625 */
      nd->flags |= GNDT_F_FITTED;
      count_rough_normal(nd); /* recorded for every fitted voxel; Slopes use it (:642,659) */
      int up = 0, down = 0;
      if (P->demand == GNDT_DEMAND_SLOPE) { /* :630-643 */
        if (is_slope(&st, nd, P->min_points, P->slope_interval, &up, &down)) nd->flags |= GNDT_F_SLOPE;
        if (up) nd->flags |= GNDT_F_UP;
        if (down) nd->flags |= GNDT_F_DOWN;
      } else if (P->demand == GNDT_DEMAND_TRUE) { /* :644-660: every fitted voxel, down=false */
        nd->flags |= GNDT_F_SLOPE;
      }
    }
  }
  if (P->demand == GNDT_DEMAND_TRUE) /* lazy Slope::countUp on final centroids (:275) */
    for (size_t u = 0; u < st.n_nodes; ++u)
      if ((st.nodes[u].flags & GNDT_F_SLOPE) && count_up(&st, &st.nodes[u], P->slope_interval))
        st.nodes[u].flags |= GNDT_F_UP;
  double t2 = now_s();

  /* local traversability predicates (map2D.h:197-296).  In the slope demand Slope::up
   * is never assigned (map2D.h:636) and stays false, and only !up voxels are Slopes. */
  for (size_t u = 0; u < st.n_nodes; ++u)
    if (st.nodes[u].flags & GNDT_F_SLOPE) st.nodes[u].flags |= reach_bits(&st, &st.nodes[u], P);
  double t3 = now_s();

  /* canonical tables */
  R->n_columns = st.n_cols;
  R->n_voxels = st.n_nodes;
  onode **order = (onode **)malloc((st.n_nodes + 1) * sizeof(onode *));
  for (size_t u = 0; u < st.n_nodes; ++u) order[u] = &st.nodes[u];
  qsort(order, st.n_nodes, sizeof(onode *), cmp_node_canonical);
  R->voxels = (gndt_voxel *)calloc(st.n_nodes + 1, sizeof(gndt_voxel));
  R->columns = (gndt_column *)calloc(st.n_cols + 1, sizeof(gndt_column));
  R->morton_list = (uint32_t *)calloc(st.n_cols + 1, sizeof(uint32_t));
  uint32_t *col_to_out = (uint32_t *)malloc((st.n_cols + 1) * sizeof(uint32_t));
  size_t nc = 0;
  for (size_t i = 0; i < st.n_nodes; ++i) {
    const onode *nd = order[i];
    gndt_voxel *v = &R->voxels[i];
    v->sx = nd->sx; v->sy = nd->sy; v->sz = nd->sz;
    v->count = nd->count; v->first_index = nd->first;
    for (int k = 0; k < 3; ++k) v->mean[k] = nd->cen[k];
    for (int k = 0; k < 6; ++k) v->scatter[k] = nd->cov[k];
    for (int k = 0; k < 3; ++k) { v->evals[k] = (float)nd->evals[k]; v->normal[k] = (float)nd->nrm[k]; }
    v->rough = nd->rough;
    v->flags = nd->flags;
    if (i == 0 || order[i - 1]->col != nd->col) {
      v->flags |= GNDT_F_COLUMN_HEAD;
      gndt_column *c = &R->columns[nc];
      c->sx = nd->sx; c->sy = nd->sy; c->voxel_begin = (uint32_t)i;
      c->slope_begin = (uint32_t)R->n_slopes;
      c->first_index = st.nodes[st.cols[nd->col].first_node].first;
      col_to_out[nd->col] = (uint32_t)nc++;
    }
    R->columns[nc - 1].voxel_count++;
    v->column = (uint32_t)(nc - 1);
    v->slope = 0xFFFFFFFFu;
    if (nd->flags & GNDT_F_FITTED) R->n_fitted++;
    if (nd->flags & GNDT_F_SLOPE) { v->slope = (uint32_t)R->n_slopes++; R->columns[nc - 1].slope_count++; }
  }
  for (size_t c = 0; c < st.n_cols; ++c) R->morton_list[c] = col_to_out[c];
  R->division_s = t1 - t0;
  R->calculate_s = t2 - t1;
  R->edges_s = t3 - t2;

  free(col_to_out); free(order);
  free(st.nodes); free(st.cols); free(st.slots); free(st.next_pt);
  *out = R;
  return GNDT_OK;
}

/* ------------------------------------------------------------------------------------
 * The traversability graph: for every Slope the list TwoDmap::AccessibleNeighbors returns
 * (map2D.h:530-548): countLRFB's four cells in the order left, right, forward, back, and in
 * each cell the Slopes in std::map<int,Slope*> order (ascending z) that pass countReachable's
 * tests (comand 2.5 / 4 and the checkList variant, :284-287,306-315).  Works from the
 * canonical tables of a build; slope index = rank of the SLOPE voxel in table order.
 * offsets has n_slopes + 1 entries.  Pinned against the reference's own AccessibleNeighbors
 * inside oracle/_ref (ref_adapter_check.cpp, gndt_ref_graph_check).
 * ---------------------------------------------------------------------------------- */
#include "../include/gndt_lookup.h"
int gndt_oracle_edges(const gndt_voxel *vox, size_t nv, const gndt_column *cols, size_t nc, const gndt_params *P,
                      uint32_t **offsets_out, uint32_t **targets_out, size_t *n_slopes_out, size_t *n_targets_out) {
  if (!vox || !cols || !P || !offsets_out || !targets_out) return GNDT_ERR_INVALID_ARG;
  size_t ns = 0;
  for (size_t v = 0; v < nv; ++v) ns += (vox[v].flags & GNDT_F_SLOPE) ? 1 : 0;
  uint32_t *off = (uint32_t *)calloc(ns + 1, sizeof(uint32_t));
  size_t cap = 4 * ns + 16, nt = 0;
  uint32_t *tgt = (uint32_t *)malloc(cap * sizeof(uint32_t));
  size_t i = 0;
  for (size_t v = 0; v < nv; ++v) {
    if (!(vox[v].flags & GNDT_F_SLOPE)) continue;
    const gndt_voxel *c = &vox[v];
    off[i++] = (uint32_t)nt;
    for (int d = 0; d < 4; ++d) {
      const int64_t col = gndtl_neighbor_column(cols, nc, c->sx, c->sy, d);
      if (col < 0) continue;
      for (uint32_t u = cols[col].voxel_begin; u < cols[col].voxel_begin + cols[col].voxel_count; ++u) {
        const gndt_voxel *s = &vox[u];
        if (!(s->flags & GNDT_F_SLOPE)) continue;
        if (s->flags & GNDT_F_UP) continue;
        if (!(s->rough <= P->rough_max)) continue;
        if (!(count_angle(s->normal, c->normal) <= P->angle_max_deg)) continue;
        float dz = s->mean[2] - c->mean[2];
        if (!(fabsf(dz) <= P->reach_height)) continue;
        if (nt == cap) { cap *= 2; tgt = (uint32_t *)realloc(tgt, cap * sizeof(uint32_t)); }
        tgt[nt++] = s->slope;
      }
    }
  }
  off[ns] = (uint32_t)nt;
  *offsets_out = off; *targets_out = tgt;
  if (n_slopes_out) *n_slopes_out = ns;
  if (n_targets_out) *n_targets_out = nt;
  return GNDT_OK;
}
void gndt_oracle_free_edges(uint32_t *offsets, uint32_t *targets) { free(offsets); free(targets); }

void gndt_oracle_free(oracle_result *R) {
  if (!R) return;
  free(R->voxels); free(R->columns); free(R->morton_list);
  free(R);
}

/* ------------------------------------------------------------------------------------
 * The reference's only deterministic in-tree fixture: the "bridge_ground" cloud of
 * src/test/genePcd.cpp:29-200, stated as a table of sweeps.  Each sweep is the pair of
 * nested `for (float u = u0; u <cmp> u1; u += 0.025)` loops of that file: the counter is
 * a float, the step and bounds are doubles (so u = (float)((double)u + 0.025)).
 * Fills at most cap points (x,y,z,0 as 4 floats); the cloud is 600x600 zero-initialised
 * points (genePcd.cpp:30-33) so the tail stays (0,0,0).  Returns the points assigned.
 * ---------------------------------------------------------------------------------- */
typedef struct sweep {
  char outer, inner;          /* axis swept by the outer / inner loop: 'x','y','z'      */
  double o0, o1; int o_incl;  /* outer: from o0 while (o_incl ? <= : <) o1              */
  double i0, i1; int i_incl;
  char rule;                  /* how the third coordinate is set                        */
  double k;                   /* constant used by the rule                              */
} sweep;

size_t gndt_oracle_bridge_ground(float *xyzw, size_t cap) {
  static const sweep S[] = {
      /* ramps (genePcd.cpp:36-52): z = 0.5x+0.5 and z = -0.5x+8.5 */
      {'x', 'y', 1, 5.025, 1, 1, 5, 1, 'u', 0},
      {'x', 'y', 11 - 0.025, 15, 1, 1, 5, 1, 'd', 0},
      /* deck z = 3 (:54-61) */
      {'x', 'y', 5, 11, 0, 1, 5, 0, 'c', 3},
      /* ground z = 1 in three bands (:72-95) and four corner pads (:97-128) */
      {'x', 'y', 0, 5, 1, 0, 6, 1, 'c', 1},
      {'x', 'y', 6, 10, 1, 0, 6, 1, 'c', 1},
      {'x', 'y', 11, 16, 1, 0, 6, 1, 'c', 1},
      {'x', 'y', 5, 6, 1, 0, 1, 1, 'c', 1},
      {'x', 'y', 5, 6, 1, 5, 6, 1, 'c', 1},
      {'x', 'y', 10, 11, 1, 0, 1, 1, 'c', 1},
      {'x', 'y', 10, 11, 1, 5, 6, 1, 'c', 1},
      /* pier walls at y = 1 / y = 5 (:131-162) */
      {'x', 'z', 5, 6, 1, 1, 3, 1, 'y', 1},
      {'x', 'z', 10, 11, 1, 1, 3, 1, 'y', 1},
      {'x', 'z', 5, 6, 1, 1, 3, 1, 'y', 5},
      {'x', 'z', 10, 11, 1, 1, 3, 1, 'y', 5},
      /* pier walls at x = 5, 6, 10, 11 (:165-196) */
      {'y', 'z', 1, 5, 1, 1, 3, 1, 'x', 5},
      {'y', 'z', 1, 5, 1, 1, 3, 1, 'x', 6},
      {'y', 'z', 1, 5, 1, 1, 3, 1, 'x', 10},
      {'y', 'z', 1, 5, 1, 1, 3, 1, 'x', 11},
  };
  size_t i = 0;
  memset(xyzw, 0, cap * 4 * sizeof(float));
  for (size_t s = 0; s < sizeof(S) / sizeof(S[0]); ++s) {
    const sweep *w = &S[s];
    for (float o = (float)w->o0; w->o_incl ? ((double)o <= w->o1) : ((double)o < w->o1);
         o = (float)((double)o + 0.025)) {
      for (float in = (float)w->i0; w->i_incl ? ((double)in <= w->i1) : ((double)in < w->i1);
           in = (float)((double)in + 0.025)) {
        if (i >= cap) return i;
        float x = 0, y = 0, z = 0;
        if (w->outer == 'x') x = o; else if (w->outer == 'y') y = o;
        if (w->inner == 'y') y = in; else if (w->inner == 'z') z = in;
        switch (w->rule) {
          case 'u': z = (float)((double)x * 0.5 + 0.5); break;
          case 'd': z = (float)((double)x * (-0.5) + 8.5); break;
          case 'c': z = (float)w->k; break;
          case 'y': y = (float)w->k; break;
          case 'x': x = (float)w->k; break;
        }
        xyzw[i * 4 + 0] = x; xyzw[i * 4 + 1] = y; xyzw[i * 4 + 2] = z;
        i++;
      }
    }
  }
  return i;
}
