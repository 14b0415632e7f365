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
// Traffic per pass: read 16 B + write 16 B per point (one sweep, decoupled look-back), the
// digit histogram of pass p+1 is accumulated while pass p moves the data.
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
#define GNDT_SORT_MINBLOCKS 2
#endif
constexpr int kSortThreads = GNDT_SORT_THREADS;
constexpr int kSortItems = GNDT_SORT_ITEMS;
constexpr int kSortTile = kSortThreads * kSortItems;  // 3072 points = 48 KB staged
constexpr int kSortWarps = kSortThreads / 32;

struct SortSmem {
  float4 stage[kSortTile];
  unsigned short sdig[kSortTile];
  u32 whist[kSortWarps][kRadixBins];
  u32 tile_off[kRadixBins];
  u32 gbase[kRadixBins];
  u32 later_hist[kMaxPasses - 1][kRadixBins];  // pass 0 only: digit histograms of passes 1..5
  u32 warp_sums[16];
  u32 tile_id;
  u32 n_valid_tile;
};

// Load point i of a strided cloud (first 12 bytes of each record are x,y,z).
__device__ __forceinline__ float4 load_point(const float *in, size_t stride_f, size_t i, bool vec) {
  if (vec) return ld_stream(reinterpret_cast<const float4 *>(in) + i);
  const float *p = in + i * stride_f;
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
}

// ---------------------------------------------------------------------------------------
// K1a: bounds of the contiguous indices + histogram of the (bounds-independent) first
// digit + validity counters.  transMortonXYZ arithmetic (map2D.h:950-973) happens here
// for the first time; pass 0 repeats it bit-identically.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bounds_kernel(Ctl *ctl, u32 *hist0, const float *in,
                                                     size_t stride_f, size_t n_in, size_t start,
                                                     DevParams P) {
  __shared__ u32 sh[kRadixBins];
  __shared__ int red[6][8];
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
  int mx = -kIdxBias, my = -kIdxBias, mz = -kIdxBias, nx = -kIdxBias, ny = -kIdxBias, nz = -kIdxBias;
  u32 n_ok = 0, n_drop = 0, n_out = 0;
  for (size_t i = start + (size_t)blockIdx.x * blockDim.x + tid; i < n_in; i += (size_t)gridDim.x * blockDim.x) {
    float4 p = load_point(in, stride_f, i, vec);
    int cx, cy, cz;
    if (!point_indices(p.x, p.y, p.z, o, P.grid_len, P.z_len, cx, cy, cz)) { n_drop++; continue; }
    if (tiled && (cx < P.tile_lo || cx >= P.tile_hi)) { n_out++; continue; }
    n_ok++;
    mx = max(mx, cx); my = max(my, cy); mz = max(mz, cz);
    nx = max(nx, -cx); ny = max(ny, -cy); nz = max(nz, -cz);
    {  // warp-aggregated: flat scenes put most of a warp into one z bin
      const u32 d = first_digit(cz);
      const u32 peers = __match_any_sync(__activemask(), d);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[d], (u32)__popc(peers));
    }
  }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o2)); my = max(my, __shfl_xor_sync(0xffffffffu, my, o2));
    mz = max(mz, __shfl_xor_sync(0xffffffffu, mz, o2)); nx = max(nx, __shfl_xor_sync(0xffffffffu, nx, o2));
    ny = max(ny, __shfl_xor_sync(0xffffffffu, ny, o2)); nz = max(nz, __shfl_xor_sync(0xffffffffu, nz, o2));
    n_ok += __shfl_xor_sync(0xffffffffu, n_ok, o2); n_drop += __shfl_xor_sync(0xffffffffu, n_drop, o2);
    n_out += __shfl_xor_sync(0xffffffffu, n_out, o2);
  }
  if (lane == 0) {
    red[0][warp] = mx; red[1][warp] = my; red[2][warp] = mz; red[3][warp] = nx; red[4][warp] = ny; red[5][warp] = nz;
    atomicAdd(&cnt[0], n_ok); atomicAdd(&cnt[1], n_drop); atomicAdd(&cnt[2], n_out);
  }
  __syncthreads();
  if (tid < 6) {
    int m = red[tid][0];
    for (int w = 1; w < 8; ++w) m = max(m, red[tid][w]);
    if (cnt[0]) atomicMax(&ctl->max_cx_b + tid, (u32)(m + kIdxBias));
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
// the digit histograms of ALL later passes (the indices are in registers anyway);
// FIRST=false moves already-tagged points between the two work buffers and evaluates only
// the key fields its digit covers (usually one IEEE divide per point instead of three).
//
//  1. all loads of the tile are issued before any dependent work (8 x 512 B in flight per warp)
//  2. stable in-tile rank: __match_any_sync groups equal digits inside a warp, per-warp
//     digit counters in shared memory, then a scan across the 12 warps per digit
//  3. tile digit counts are published for the decoupled look-back while the points are
//     reordered into shared memory (so the global writes are runs of equal digits)
//  4. coalesced copy-out to the digit's global run
// ---------------------------------------------------------------------------------------
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads, GNDT_SORT_MINBLOCKS)
sort_pass_kernel(Ctl *ctl, int pass, const float *in_raw, size_t stride_f, size_t n_in, size_t start,
                 const float4 *src, float4 *dst, u32 *lb, u32 *hist_all, DevParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortSmem &S = *reinterpret_cast<SortSmem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const int n_passes = ctl->n_passes;
  if (!FIRST && pass >= n_passes) return;
  if (tid == 0) S.tile_id = atomicAdd(&ctl->ticket[pass], 1u);
  for (int i = tid; i < kSortWarps * kRadixBins; i += kSortThreads) (&S.whist[0][0])[i] = 0;
  if (FIRST)
    for (int i = tid; i < (kMaxPasses - 1) * kRadixBins; i += kSortThreads) (&S.later_hist[0][0])[i] = 0;
  __syncthreads();
  const int tile = (int)S.tile_id;
  const size_t M = FIRST ? (n_in - start) : (size_t)ctl->n_valid;
  const size_t base = (size_t)tile * kSortTile;
  if (base >= M) return;
  const int cnt = (int)min((size_t)kSortTile, M - base);

  const KeyLayout L = load_layout(ctl);
  const int shift = ctl->shift[pass], bits = ctl->bits[pass];
  const u32 mask = (1u << bits) - 1u;
  const float o[3] = {ctl->origin[0], ctl->origin[1], ctl->origin[2]};
  const bool tiled = P.tile_lo < P.tile_hi;
  const bool vec = FIRST && (stride_f == 4) && ((reinterpret_cast<uintptr_t>(in_raw) & 15) == 0);
  // key fields this pass's digit overlaps: z [0,bz), y [bz,bz+by), x [bz+by, ...)
  const int need = FIRST ? 7
                         : ((shift + bits > L.bz + L.by ? 1 : 0) | ((shift < L.bz + L.by && shift + bits > L.bz) ? 2 : 0) |
                            (shift < L.bz ? 4 : 0));

  // ---- 1. loads first
  float4 e[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int i = warp * (32 * kSortItems) + k * 32 + lane;
    if (i < cnt) {
      if (FIRST) e[k] = load_point(in_raw, stride_f, start + base + i, vec);
      else e[k] = ld_stream(src + base + i);
    }
  }
  int q_shift[kMaxPasses];
  u32 q_mask[kMaxPasses];
#pragma unroll
  for (int q = 0; q < kMaxPasses; ++q) {
    q_shift[q] = FIRST ? ctl->shift[q] : 0;
    q_mask[q] = FIRST ? ((1u << ctl->bits[q]) - 1u) : 0u;
  }
  u32 dg[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int i = warp * (32 * kSortItems) + k * 32 + lane;
    dg[k] = kInvalidDigit;
    if (i < cnt) {
      int cx, cy, cz;
      if (FIRST) {
        e[k].w = __uint_as_float((u32)(start + base + i));
        bool ok = point_indices(e[k].x, e[k].y, e[k].z, o, P.grid_len, P.z_len, cx, cy, cz);
        if (ok && tiled && (cx < P.tile_lo || cx >= P.tile_hi)) ok = false;
        if (ok) {
          dg[k] = first_digit(cz);
          const u64 key = compact_key(cx, cy, cz, L);
#pragma unroll
          for (int q = 1; q < kMaxPasses; ++q)
            if (q < n_passes) atomicAdd(&S.later_hist[q - 1][(u32)(key >> q_shift[q]) & q_mask[q]], 1u);
        }
      } else {
        point_indices_masked(e[k].x, e[k].y, e[k].z, o, P.grid_len, P.z_len, need, cx, cy, cz);
        u64 key = 0;
        if (need & 1) key |= (u64)(u32)(cx - L.cx_min) << (L.by + L.bz);
        if (need & 2) key |= (u64)(u32)(cy - L.cy_min) << L.bz;
        if (need & 4) key |= (u64)(u32)(cz - L.cz_bias);
        dg[k] = (u32)(key >> shift) & mask;
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
      S.whist[warp][dg[k]] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[k] = old + r;
    __syncwarp();
  }
  __syncthreads();

  // ---- 3. per digit (thread d < 256): scan over the warps, tile count, publication
  u32 tile_count = 0;
  u32 *my_word = lb + (size_t)tile * kRadixBins + (tid & (kRadixBins - 1));
  if (tid < kRadixBins) {
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const u32 t = S.whist[w][tid];
      S.whist[w][tid] = tile_count;
      tile_count += t;
    }
    st_relaxed(my_word, (tile == 0 ? kFlagIncl : kFlagAgg) | tile_count);
  }
  // exclusive scans over the 256 digits (threads >= 256 contribute zeros)
  const u32 toff = block_exclusive_scan(tile_count, S.warp_sums);
  const u32 digit_base = block_exclusive_scan(tid < kRadixBins ? hist_all[(size_t)pass * kRadixBins + tid] : 0u, S.warp_sums);
  if (tid < kRadixBins) S.tile_off[tid] = toff;
  __syncthreads();

  // ---- reorder into shared memory by digit (stable)
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    if (dg[k] != kInvalidDigit) {
      const u32 pos = S.tile_off[dg[k]] + S.whist[warp][dg[k]] + rank[k];
      S.stage[pos] = e[k];
      S.sdig[pos] = (unsigned short)dg[k];
    }
  }

  // ---- decoupled look-back for digit `tid`
  if (tid < kRadixBins) {
    u32 prefix = 0;
    if (tile > 0) {
      // walk back over the predecessors' words, kLookBatch independent loads per round trip
      constexpr int kLookBatch = 8;
      bool done = false;
      for (int j = tile - 1; j >= 0 && !done; j -= kLookBatch) {
        u32 w[kLookBatch];
#pragma unroll
        for (int q = 0; q < kLookBatch; ++q)
          w[q] = (j - q >= 0) ? ld_relaxed(my_word - (size_t)(tile - (j - q)) * kRadixBins) : kFlagIncl;
#pragma unroll
        for (int q = 0; q < kLookBatch; ++q) {
          if (done) break;
          u32 spins = 0;
          while ((w[q] & kFlagMask) == 0 && ++spins < kSpinLimit)
            w[q] = ld_relaxed(my_word - (size_t)(tile - (j - q)) * kRadixBins);
          if ((w[q] & kFlagMask) == 0) { atomicOr(&ctl->err, kErrWatchdog); done = true; break; }
          prefix += w[q] & kValMask;
          if (w[q] & kFlagIncl) done = true;
        }
      }
      st_relaxed(my_word, kFlagIncl | (prefix + tile_count));
    }
    S.gbase[tid] = digit_base + prefix - toff;
    if (tid == kRadixBins - 1) S.n_valid_tile = toff + tile_count;
  }
  __syncthreads();

  // ---- 4. copy out: consecutive threads write consecutive addresses inside a digit run
  const int n_valid_tile = (int)S.n_valid_tile;
  for (int i = tid; i < n_valid_tile; i += kSortThreads) {
    const u32 d = S.sdig[i];
    st_stream(dst + S.gbase[d] + i, S.stage[i]);
  }
  if (FIRST)
    for (int i = tid; i < (n_passes - 1) * kRadixBins; i += kSortThreads) {
      const u32 c = (&S.later_hist[0][0])[i];
      if (c) atomicAdd(&hist_all[kRadixBins + i], c);
    }
}

}  // namespace gndt
