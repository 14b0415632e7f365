/* gndt_lookup.h — integer-keyed lookups in the result tables (host side, header-only C).
 *
 * The reference's planner finds a cell by building the decimal-Morton string of its indices and
 * searching a std::map<string,Cell*> (countReachable include/map2D.h:266-296, CollisionCheck
 * :351-411, computeCost :1285-1397, findRoute include/GlobalPlan.h:56-61,79-83): ~1.5 us of
 * string building plus O(log C) string compares per lookup.  The column table of a build is
 * already sorted by (x, y) cell index, so the same lookup is a binary search over two integers
 * (SURVEY.md section 8(f) rank 2).  These functions work on the caller's copies of the tables;
 * libgndt.so exports them as gndt_find_column / gndt_find_slope / gndt_neighbor_column and
 * adapter/gndt_twodmap_adapter.h uses them to hand the planner Cell pointers without strings.
 *
 * Order of the column table: ascending CONTIGUOUS index c = s > 0 ? s - 1 : s of the signed
 * non-zero indices (sx, then sy) — the order in which gndt_build writes it. */
#ifndef GNDT_LOOKUP_H
#define GNDT_LOOKUP_H
#include "gndt.h"

#ifdef __cplusplus
extern "C" {
#endif

static inline int32_t gndtl_contiguous(int32_t s) { return s > 0 ? s - 1 : s; }
static inline int32_t gndtl_signed(int32_t c) { return c >= 0 ? c + 1 : c; }

/* Index of column (sx, sy) in cols[0..n), or -1 (also for sx == 0 / sy == 0, which no cell has). */
static inline int64_t gndtl_find_column(const gndt_column *cols, size_t n, int32_t sx, int32_t sy) {
  if (!cols || sx == 0 || sy == 0) return -1;
  const int32_t cx = gndtl_contiguous(sx), cy = gndtl_contiguous(sy);
  size_t lo = 0, hi = n;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    const int32_t mx = gndtl_contiguous(cols[mid].sx), my = gndtl_contiguous(cols[mid].sy);
    if (mx < cx || (mx == cx && my < cy)) lo = mid + 1; else hi = mid;
  }
  return (lo < n && cols[lo].sx == sx && cols[lo].sy == sy) ? (int64_t)lo : -1;
}

/* The 4-neighbourhood of countLRFB (include/map2D.h:197-263): one step in signed non-zero index
 * space, crossing the quadrant boundary between -1 and +1.  dir: 0 left (y-1), 1 right (y+1),
 * 2 forward (x+1), 3 back (x-1) — the directions of GNDT_F_REACH_L/R/F/B.  Returns the column
 * index or -1 when that cell is empty. */
static inline int64_t gndtl_neighbor_column(const gndt_column *cols, size_t n, int32_t sx, int32_t sy, int dir) {
  if (sx == 0 || sy == 0 || dir < 0 || dir > 3) return -1;
  int32_t cx = gndtl_contiguous(sx), cy = gndtl_contiguous(sy);
  if (dir == 0) cy -= 1; else if (dir == 1) cy += 1; else if (dir == 2) cx += 1; else cx -= 1;
  return gndtl_find_column(cols, n, gndtl_signed(cx), gndtl_signed(cy));
}

/* Cell::map_slope.find(sz) of column (sx, sy): index into the slope table, or -1.  The slopes
 * of a column are contiguous and ascending in sz. */
static inline int64_t gndtl_find_slope(const gndt_column *cols, size_t n_cols, const gndt_slope *slopes, int32_t sx,
                                       int32_t sy, int32_t sz) {
  const int64_t c = gndtl_find_column(cols, n_cols, sx, sy);
  if (c < 0 || !slopes) return -1;
  size_t lo = cols[c].slope_begin, hi = (size_t)cols[c].slope_begin + cols[c].slope_count;
  const size_t end = hi;
  while (lo < hi) {
    const size_t mid = lo + (hi - lo) / 2;
    if (slopes[mid].sz < sz) lo = mid + 1; else hi = mid;
  }
  return (lo < end && slopes[lo].sz == sz) ? (int64_t)lo : -1;
}

#ifdef __cplusplus
}
#endif
#endif
