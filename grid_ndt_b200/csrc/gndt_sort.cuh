// gndt_sort.cuh — binning of the cloud by (cx,cy,cz): bounds pass, key-layout plan and a
// single-sweep LSD radix partition that moves the 16-byte points themselves.
//
// Replaces the reference's uniformDivision loop (src/receiver.cpp:41-93,150-154): there,
// every point walks a std::multimap<string,OcNode*> keyed by a decimal Morton string and
// is appended to its voxel's point list.  Here the cloud is stably partitioned so that
// each voxel is one contiguous run (in cloud order) and each x-y column is a contiguous
// run of voxels ordered by z.  Stability gives `first_index` (the first-seen order that
// the reference's isSlope depends on) for free.
//
// Traffic per pass: read 16 B + write 16 B per point (one sweep, decoupled look-back).
// A tile (3072 points, 48 KB) is brought into shared memory by ONE TMA bulk copy
// (cp.async.bulk, completion on an mbarrier) and never passes through registers as a
// whole: threads read the one or two coordinates their digit needs, rank, write a 4-byte
// permutation entry, and the copy-out gathers 16-byte points through it.
#pragma once
#include "gndt_device.cuh"

namespace gndt {

#ifndef GNDT_SORT_THREADS
#define GNDT_SORT_THREADS 384
#endif
#ifndef GNDT_SORT_ITEMS
#define GNDT_SORT_ITEMS 8
#endif
#ifndef GNDT_SORT_MINBLOCKS
#define GNDT_SORT_MINBLOCKS 3
#endif
#ifdef GNDT_SORT_WHIST16
typedef unsigned short whist_t;
#else
typedef u32 whist_t;
#endif
constexpr int kSortThreads = GNDT_SORT_THREADS;
constexpr int kSortItems = GNDT_SORT_ITEMS;
constexpr int kSortTile = kSortThreads * kSortItems;  // 3072 points = 48 KB staged
constexpr int kSortWarps = kSortThreads / 32;
static_assert(kSortThreads >= kRadixBins && kSortTile <= 65536, "tile shape");
#ifndef GNDT_SORT_GROUP
#define GNDT_SORT_GROUP 16
#endif
constexpr int kSortGroup = GNDT_SORT_GROUP;  // tiles per look-back group
constexpr int kGroupSumBits = 26;            // group word: arrivals << 26 | sum of the tiles' digit counts
static_assert(kSortGroup >= 1 && kSortGroup < 64 && (long long)kSortGroup * kSortTile < (1ll << kGroupSumBits), "group word");

struct __align__(128) SortSmem {
  float4 in[kSortTile];            // the tile, in arrival order (TMA destination)
  u32 slot[kSortTile];             // digit << 16 | source position, in digit order; pass 0 first
                                   // uses this space for the histograms of the later digits
  whist_t whist[kSortWarps][kRadixBins];  // per-warp digit counts (<= 32 * kSortItems each)
  u32 tile_off[kRadixBins];
  u32 gbase[kRadixBins];
  u32 warp_sums[16];
  unsigned long long mbar;
  u32 tile_id;
  u32 n_valid_tile;
};
static_assert((kMaxPasses - 1) * kRadixBins <= kSortTile, "later-digit histograms alias slot[]");

// Load point i of a strided cloud (first 12 bytes of each record are x,y,z).
__device__ __forceinline__ float4 load_point(const float *in, size_t stride_f, size_t i, bool vec) {
  if (vec) return ld_stream(reinterpret_cast<const float4 *>(in) + i);
  const float *p = in + i * stride_f;
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
}

// ---------------------------------------------------------------------------------------
// K1a: bounds + histogram of the (bounds-independent) first digit + validity counters.
// The index map of transMortonXYZ (map2D.h:950-973) is monotone in each coordinate, so the
// index bounds are the indices of the coordinate bounds: this pass keeps min/max of the raw
// floats of valid points and divides only for the z index (the first digit).  A point is
// valid iff |p - p0| <= P.max_abs[axis], the largest offset whose index is <= GNDT_MAX_INDEX
// (found on the host with the same IEEE operations).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bounds_kernel(Ctl *ctl, u32 *hist0, const float *in,
                                                     size_t stride_f, size_t n_in, size_t start,
                                                     DevParams P) {
  __shared__ u32 sh[kRadixBins];
  __shared__ float red[6][8];
  __shared__ u32 cnt[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec = (stride_f == 4) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  float o[3];
  if (P.origin_first) { o[0] = __ldg(in); o[1] = __ldg(in + 1); o[2] = __ldg(in + 2); }
  else { o[0] = P.origin[0]; o[1] = P.origin[1]; o[2] = P.origin[2]; }
  if (blockIdx.x == 0 && tid == 0) { ctl->origin[0] = o[0]; ctl->origin[1] = o[1]; ctl->origin[2] = o[2]; }
  sh[tid] = 0;
  if (tid < 3) cnt[tid] = 0;
  __syncthreads();
  const bool tiled = P.tile_lo < P.tile_hi;
  const float inf = __int_as_float(0x7f800000);
  float lo_x = inf, lo_y = inf, lo_z = inf, hi_x = -inf, hi_y = -inf, hi_z = -inf;
  u32 n_ok = 0, n_drop = 0, n_out = 0;
  // 4 points per thread per trip: the loads are issued together, all lanes stay converged
  // (no early `continue`) so the warp-aggregated histogram update sees full warps
  constexpr int kUnroll = 4;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t w0 = start + (size_t)blockIdx.x * blockDim.x + (tid & ~31); w0 < n_in; w0 += kUnroll * step) {  // warp-uniform trip count
    const size_t i0 = w0 + lane;
    float4 p[kUnroll];
#pragma unroll
    for (int u2 = 0; u2 < kUnroll; ++u2) {
      const size_t i = i0 + u2 * step;
      p[u2] = (i < n_in) ? load_point(in, stride_f, i, vec) : make_float4(inf, inf, inf, 0.f);
    }
#pragma unroll
    for (int u2 = 0; u2 < kUnroll; ++u2) {
      const size_t i = i0 + u2 * step;
      const bool live = i < n_in;
      bool ok = live && fabsf(__fsub_rn(p[u2].x, o[0])) <= P.max_abs[0] && fabsf(__fsub_rn(p[u2].y, o[1])) <= P.max_abs[0] &&
                fabsf(__fsub_rn(p[u2].z, o[2])) <= P.max_abs[1];
      if (live && !ok) n_drop++;  // NaN / Inf / out of the supported index range
      if (ok && tiled) {
        int cx;
        axis_idx(p[u2].x, o[0], false, P, cx);
        if (cx < P.tile_lo || cx >= P.tile_hi) { n_out++; ok = false; }
      }
      u32 d = kInvalidDigit;
      if (ok) {
        n_ok++;
        lo_x = fminf(lo_x, p[u2].x); lo_y = fminf(lo_y, p[u2].y); lo_z = fminf(lo_z, p[u2].z);
        hi_x = fmaxf(hi_x, p[u2].x); hi_y = fmaxf(hi_y, p[u2].y); hi_z = fmaxf(hi_z, p[u2].z);
        int cz;
        axis_idx(p[u2].z, o[2], true, P, cz);
        d = first_digit(cz);
      }
      const u32 peers = __match_any_sync(0xffffffffu, d);  // flat scenes: most lanes share a z bin
      if (ok && lane == __ffs(peers) - 1) atomicAdd(&sh[d], (u32)__popc(peers));
    }
  }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    lo_x = fminf(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o2)); lo_y = fminf(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o2));
    lo_z = fminf(lo_z, __shfl_xor_sync(0xffffffffu, lo_z, o2)); hi_x = fmaxf(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o2));
    hi_y = fmaxf(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o2)); hi_z = fmaxf(hi_z, __shfl_xor_sync(0xffffffffu, hi_z, o2));
    n_ok += __shfl_xor_sync(0xffffffffu, n_ok, o2); n_drop += __shfl_xor_sync(0xffffffffu, n_drop, o2);
    n_out += __shfl_xor_sync(0xffffffffu, n_out, o2);
  }
  if (lane == 0) {
    red[0][warp] = hi_x; red[1][warp] = hi_y; red[2][warp] = hi_z; red[3][warp] = lo_x; red[4][warp] = lo_y; red[5][warp] = lo_z;
    atomicAdd(&cnt[0], n_ok); atomicAdd(&cnt[1], n_drop); atomicAdd(&cnt[2], n_out);
  }
  __syncthreads();
  if (tid < 6 && cnt[0]) {
    float m = red[tid][0];
    for (int w = 1; w < 8; ++w) m = (tid < 3) ? fmaxf(m, red[tid][w]) : fminf(m, red[tid][w]);
    const int ax = tid % 3;
    int c;
    axis_idx(m, o[ax], ax == 2, P, c);  // index of the extreme coordinate
    atomicMax(&ctl->max_cx_b + tid, (u32)((tid < 3 ? c : -c) + kIdxBias));
  }
  if (tid == 0) {
    if (cnt[0]) atomicAdd(&ctl->n_valid, (u64)cnt[0]);
    if (cnt[1]) atomicAdd(&ctl->n_dropped, (u64)cnt[1]);
    if (cnt[2]) atomicAdd(&ctl->n_outside, (u64)cnt[2]);
  }
  if (sh[tid]) atomicAdd(&hist0[tid], sh[tid]);
}

__device__ __forceinline__ int bit_width(u32 v) { return 32 - __clz(v); }

// ---------------------------------------------------------------------------------------
// K1b: derive the key layout from the bounds: field widths, z bias, digit schedule.
// Pass 0 is always the low 8 bits of the z field; the remaining bits are split evenly into
// ceil(R/8) digits of at most 8 bits.
// ---------------------------------------------------------------------------------------
__global__ void plan_kernel(Ctl *ctl) {
  if (threadIdx.x != 0) return;
  if (ctl->n_valid == 0) {
    ctl->cx_min = ctl->cy_min = 0; ctl->cz_bias = -128; ctl->bx = ctl->by = 1; ctl->bz = 8;
    ctl->n_passes = 1; ctl->shift[0] = 0; ctl->bits[0] = 8; ctl->cx_max = 0;
    return;
  }
  int cx_max = (int)ctl->max_cx_b - kIdxBias, cy_max = (int)ctl->max_cy_b - kIdxBias, cz_max = (int)ctl->max_cz_b - kIdxBias;
  int cx_min = kIdxBias - (int)ctl->max_ncx_b, cy_min = kIdxBias - (int)ctl->max_ncy_b, cz_min = kIdxBias - (int)ctl->max_ncz_b;
  int t = cz_min + 128;                      // floor division by 256
  int fl = (t >= 0) ? (t >> 8) : -((255 - t) >> 8);
  int cz_bias = fl * 256 - 128;
  int bx = max(1, bit_width((u32)(cx_max - cx_min)));
  int by = max(1, bit_width((u32)(cy_max - cy_min)));
  int bz = max(8, bit_width((u32)(cz_max - cz_bias)));
  ctl->cx_min = cx_min; ctl->cy_min = cy_min; ctl->cz_bias = cz_bias; ctl->cx_max = cx_max;
  ctl->bx = bx; ctl->by = by; ctl->bz = bz;
  int R = bx + by + bz - 8;
  int extra = (R + 7) / 8;
  int base = R / extra, rem = R % extra;
  ctl->shift[0] = 0; ctl->bits[0] = 8;
  int s = 8;
  for (int p = 1; p <= extra; ++p) {
    int b = base + (p <= rem ? 1 : 0);
    ctl->shift[p] = s; ctl->bits[p] = b;
    s += b;
  }
  ctl->n_passes = 1 + extra;
}

// ---------------------------------------------------------------------------------------
// K2: one radix partition pass.  FIRST=true reads the caller's cloud (any stride), drops
// invalid / out-of-strip points, tags each survivor with its cloud index in .w and builds
// the digit histograms of ALL later passes; FIRST=false moves already-tagged points between
// the two work buffers and evaluates only the key fields its digit covers (usually one IEEE
// divide per point instead of three).
//
//  1. the tile is fetched into shared memory by one TMA bulk copy (strided / unaligned
//     caller clouds fall back to per-thread loads)
//  2. stable in-tile rank: __match_any_sync groups equal digits inside a warp, per-warp
//     digit counters in shared memory, then a scan across the warps per digit
//  3. tile digit counts are published for the decoupled look-back while the permutation
//     (digit, source position) is written in digit order
//  4. copy-out: consecutive threads gather their point through the permutation and write
//     consecutive addresses inside a digit run
// ---------------------------------------------------------------------------------------
template <bool FIRST, bool FAST>
__global__ void __launch_bounds__(kSortThreads, GNDT_SORT_MINBLOCKS)
sort_pass_kernel(Ctl *ctl, int pass, const float *in_raw, size_t stride_f, size_t n_in, size_t start,
                 const float4 *src, float4 *dst, u32 *lb, u64 *glb, u32 *hist_all, DevParams P) {
  extern __shared__ __align__(128) unsigned char smem_sort[];
  SortSmem &S = *reinterpret_cast<SortSmem *>(smem_sort);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const int n_passes = ctl->n_passes;
  if (!FIRST && pass >= n_passes) return;
  if (tid == 0) {
    S.tile_id = atomicAdd(&ctl->ticket[pass], 1u);
    mbar_init(&S.mbar, 1);
  }
  const int bits = ctl->bits[pass];
  const int n_bins = 1 << bits;  // digits of this pass: 256 or fewer
  for (int i = tid; i < (kSortWarps << bits); i += kSortThreads) S.whist[i >> bits][i & (n_bins - 1)] = 0;
  u32 *later_hist = S.slot;  // [kMaxPasses-1][256], pass 0 only
  if (FIRST)
    for (int i = tid; i < (n_passes - 1) * kRadixBins; i += kSortThreads) later_hist[i] = 0;
  __syncthreads();
  const int tile = (int)S.tile_id;
  const size_t M = FIRST ? (n_in - start) : (size_t)ctl->n_valid;
  const size_t base = (size_t)tile * kSortTile;
  if (base >= M) return;
  const int cnt = (int)min((size_t)kSortTile, M - base);

  // ---- 1. tile -> shared memory
  const bool vec = !FIRST || ((stride_f == 4) && ((reinterpret_cast<uintptr_t>(in_raw) & 15) == 0));
  if (vec) {
    if (tid == 0) {
      const float4 *g = FIRST ? reinterpret_cast<const float4 *>(in_raw) + start + base : src + base;
      tma_load_1d(S.in, g, (u32)cnt * 16u, &S.mbar);
    }
  } else {
    for (int i = tid; i < cnt; i += kSortThreads) S.in[i] = load_point(in_raw, stride_f, start + base + i, false);
  }

  const KeyLayout L = load_layout(ctl);
  const int shift = ctl->shift[pass];
  const u32 mask = (1u << bits) - 1u;
  // digit = OR over the fields of ((field >> rs) << ls), all 32-bit: a field at key offset
  // `off` contributes its bits [shift-off, ...) when off <= shift, else lands at off-shift
  const int off_y = L.bz, off_x = L.bz + L.by;
  const int rs_z = shift, rs_y = max(shift - off_y, 0), ls_y = max(off_y - shift, 0);
  const int rs_x = max(shift - off_x, 0), ls_x = max(off_x - shift, 0);
  const float o[3] = {ctl->origin[0], ctl->origin[1], ctl->origin[2]};
  const bool tiled = P.tile_lo < P.tile_hi;
  // key fields this pass's digit overlaps: z [0,bz), y [bz,bz+by), x [bz+by, ...)
  const int need = FIRST ? 7
                         : ((shift + bits > L.bz + L.by ? 1 : 0) | ((shift < L.bz + L.by && shift + bits > L.bz) ? 2 : 0) |
                            (shift < L.bz ? 4 : 0));
  int q_shift[kMaxPasses];
  u32 q_mask[kMaxPasses];
#pragma unroll
  for (int q = 0; q < kMaxPasses; ++q) {
    q_shift[q] = FIRST ? ctl->shift[q] : 0;
    q_mask[q] = FIRST ? ((1u << ctl->bits[q]) - 1u) : 0u;
  }

  if (vec) {
    if (!mbar_wait(&S.mbar, 0)) atomicOr(&ctl->err, kErrWatchdog);
  } else {
    __syncthreads();
  }

  // ---- digits
  u32 dg[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int i = warp * (32 * kSortItems) + k * 32 + lane;
    dg[k] = kInvalidDigit;
    if (i < cnt) {
      const float4 e = S.in[i];
      int cx, cy, cz;
      if (FIRST) {
        bool ok = point_indices_t<FAST>(e.x, e.y, e.z, o, P, cx, cy, cz);
        if (ok && tiled && (cx < P.tile_lo || cx >= P.tile_hi)) ok = false;
        if (ok) {
          dg[k] = first_digit(cz);
          const u64 key = compact_key(cx, cy, cz, L);
#pragma unroll
          for (int q = 1; q < kMaxPasses; ++q)
            if (q < n_passes) atomicAdd(&later_hist[(q - 1) * kRadixBins + ((u32)(key >> q_shift[q]) & q_mask[q])], 1u);
        }
      } else {
        point_indices_masked_t<FAST>(e.x, e.y, e.z, o, P, need, cx, cy, cz);
        u32 d = 0;
        if (need & 1) d |= ((u32)(cx - L.cx_min) >> rs_x) << ls_x;
        if (need & 2) d |= ((u32)(cy - L.cy_min) >> rs_y) << ls_y;
        if (need & 4) d |= (u32)(cz - L.cz_bias) >> rs_z;
        dg[k] = d & mask;
      }
    }
  }

  // ---- 2. stable rank inside the warp's 8x32 block of points
  u32 rank[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const u32 peers = __match_any_sync(0xffffffffu, dg[k]);
    const int leader = __ffs(peers) - 1;
    const u32 r = __popc(peers & ((1u << lane) - 1u));
    u32 old = 0;
    if (lane == leader && dg[k] != kInvalidDigit) {
      old = S.whist[warp][dg[k]];
      S.whist[warp][dg[k]] = (whist_t)(old + __popc(peers));
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[k] = old + r;
    __syncwarp();
  }
  __syncthreads();

  // pass 0: the histograms of the later digits leave shared memory before slot[] reuses it
  if (FIRST) {
    for (int i = tid; i < (n_passes - 1) * kRadixBins; i += kSortThreads) {
      const u32 c = later_hist[i];
      if (c) atomicAdd(&hist_all[kRadixBins + i], c);
    }
  }

  // ---- 3. per digit (thread d < 256): scan over the warps, tile count, publication
  // A digit takes part only if it exists in this pass (tid < n_bins) and some point of the
  // cloud has it (global count != 0): empty z levels / narrow digits cost no look-back.
  u32 tile_count = 0;
  u32 *my_word = lb + (size_t)tile * kRadixBins + (tid & (kRadixBins - 1));
  const u32 global_count = (tid < n_bins) ? hist_all[(size_t)pass * kRadixBins + tid] : 0u;
  const bool live_digit = global_count != 0;
  if (live_digit) {
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const u32 t = S.whist[w][tid];
      S.whist[w][tid] = (whist_t)tile_count;
      tile_count += t;
    }
    st_relaxed(my_word, (tile == 0 ? kFlagIncl : kFlagAgg) | tile_count);
    atomicAdd(reinterpret_cast<u32 *>(glb + (size_t)(tile / kSortGroup) * kRadixBins + tid), (1u << kGroupSumBits) | tile_count);
  }
  // exclusive scans over the digits (other threads contribute zeros)
  const u32 toff = block_exclusive_scan(tile_count, S.warp_sums);
  const u32 digit_base = block_exclusive_scan(global_count, S.warp_sums);
  if (tid < kRadixBins) S.tile_off[tid] = toff;
  __syncthreads();  // also orders the later_hist reads above before the slot[] writes below

  // ---- permutation in digit order (stable)
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    if (dg[k] != kInvalidDigit) {
      const u32 pos = S.tile_off[dg[k]] + S.whist[warp][dg[k]] + rank[k];
      S.slot[pos] = (dg[k] << 16) | (u32)(warp * (32 * kSortItems) + k * 32 + lane);
    }
  }

  // ---- decoupled look-back for digit `tid`, two levels
  // Measured (clock64 per phase): with one word per (tile, digit) the walk went back ~70 tiles
  // (every in-flight predecessor that has published its count but not yet its prefix) and was
  // 22-37 % of a CTA's life.  Tiles are therefore also summed per GROUP of kSortGroup
  // consecutive tiles: each tile adds its count to the group's word with one atomic whose top
  // bits count arrivals, so a complete group costs one load instead of kSortGroup.  The walk
  // is: own group's earlier tiles (tile words), then whole groups backwards until a group
  // that already knows its inclusive prefix.
  if (tid < kRadixBins) {
    u32 prefix = 0;
    if (tile > 0 && live_digit) {
      constexpr int kLookBatch = 8;
      const int grp = tile / kSortGroup;
      const int lo = grp * kSortGroup;
      bool done = false;
      for (int j = tile - 1; j >= lo && !done; j -= kLookBatch) {
        u32 w[kLookBatch];
#pragma unroll
        for (int q = 0; q < kLookBatch; ++q)
          w[q] = (j - q >= lo) ? ld_relaxed(my_word - (size_t)(tile - (j - q)) * kRadixBins) : 0u;
#pragma unroll
        for (int q = 0; q < kLookBatch; ++q) {
          if (done || j - q < lo) break;
          u32 spins = 0;
          while ((w[q] & kFlagMask) == 0 && ++spins < kSpinLimit)
            w[q] = ld_relaxed(my_word - (size_t)(tile - (j - q)) * kRadixBins);
          if ((w[q] & kFlagMask) == 0) { atomicOr(&ctl->err, kErrWatchdog); done = true; break; }
          prefix += w[q] & kValMask;
          if (w[q] & kFlagIncl) done = true;
        }
      }
      const u64 *gw = glb + tid;
      for (int g = grp - 1; g >= 0 && !done; g -= kLookBatch) {
        u64 w[kLookBatch];
#pragma unroll
        for (int q = 0; q < kLookBatch; ++q) w[q] = (g - q >= 0) ? ld_relaxed64(gw + (size_t)(g - q) * kRadixBins) : 0ull;
#pragma unroll
        for (int q = 0; q < kLookBatch; ++q) {
          if (done || g - q < 0) break;
          u32 spins = 0;  // ready: inclusive prefix known (high word) or all kSortGroup tiles have arrived
          while (!((u32)(w[q] >> 32) & kFlagIncl) && ((u32)w[q] >> kGroupSumBits) != (u32)kSortGroup && ++spins < kSpinLimit)
            w[q] = ld_relaxed64(gw + (size_t)(g - q) * kRadixBins);
          const u32 hi = (u32)(w[q] >> 32), low = (u32)w[q];
          if (hi & kFlagIncl) { prefix += hi & kValMask; done = true; }
          else if ((low >> kGroupSumBits) == (u32)kSortGroup) prefix += low & ((1u << kGroupSumBits) - 1u);
          else { atomicOr(&ctl->err, kErrWatchdog); done = true; }
        }
      }
      st_relaxed(my_word, kFlagIncl | (prefix + tile_count));
      if (tile - lo == kSortGroup - 1)  // last tile of its group: the group's inclusive prefix
        st_relaxed(reinterpret_cast<u32 *>(glb + (size_t)grp * kRadixBins + tid) + 1, kFlagIncl | (prefix + tile_count));
    }
    S.gbase[tid] = digit_base + prefix - toff;
    if (tid == kRadixBins - 1) S.n_valid_tile = toff + tile_count;
  }
  __syncthreads();

  // ---- 4. copy out through the permutation
  const int n_valid_tile = (int)S.n_valid_tile;
  for (int i = tid; i < n_valid_tile; i += kSortThreads) {
    const u32 s = S.slot[i];
    const u32 from = s & 0xFFFFu;
    float4 e = S.in[from];
    if (FIRST) e.w = __uint_as_float(P.idx_offset + (u32)(start + base + from));
    st_stream(dst + S.gbase[s >> 16] + i, e);
  }
}

}  // namespace gndt
