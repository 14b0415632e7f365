// gndt_reduce.cuh — per-voxel NDT fit on the partitioned cloud: mean, 3x3 scatter, smallest
// eigenpair, one 96-byte record per voxel.
//
// Replaces the fitting half of TwoDmap::create2DMap (reference include/map2D.h:606-627:
// pcl::compute3DCentroid + pcl::computeCovarianceMatrix per OcNode with >= MINPOINTSIZE
// points) and OcNode::countRoughNormal (map2D.h:110-133, Eigen::EigenSolver).
//
// Numerics: the reference sums sequentially in binary32.  Here every voxel is fitted by a
// two-pass (mean, then centred products) binary64 accumulation over its contiguous run
// staged in shared memory; runs that straddle tiles are merged with Chan's pairwise
// update.  Results are rounded to binary32 once.  Two-pass + Chan keeps the reference's
// exact zeros (all points sharing a coordinate -> that scatter row is exactly 0), which
// its `roughness == 0 -> 0.01` rule (map2D.h:131-132) depends on.
#pragma once
#include "gndt_device.cuh"

namespace gndt {

constexpr int kRedThreads = 256;
constexpr int kRedItems = 8;
constexpr int kRedTile = kRedThreads * kRedItems;  // 2048 points
constexpr int kLongRun = 64;                        // runs longer than this are reduced by the whole CTA
constexpr int kMaxLong = kRedTile / kLongRun;

struct Moments {
  double n;
  double m[3];  // mean
  double s[6];  // centred scatter xx,xy,xz,yy,yz,zz
};

constexpr u32 kCarryHasLead = 1u, kCarryLeadContinues = 2u, kCarryHasTail = 4u;
struct TileCarry {
  u32 flags;
  u32 tail_slot;
  u32 tail_first;
  u32 pad;
  u64 tail_key;
  u64 pad2;
  Moments lead;  // leading run when it continues a voxel begun in an earlier tile
  Moments tail;  // last run when its voxel continues into the next tile
};

// Chan et al. pairwise combination of two (n, mean, centred scatter) triples.
__device__ __forceinline__ void merge_moments(Moments &a, const Moments &b) {
  if (b.n == 0.0) return;
  if (a.n == 0.0) { a = b; return; }
  const double n = a.n + b.n;
  const double d0 = b.m[0] - a.m[0], d1 = b.m[1] - a.m[1], d2 = b.m[2] - a.m[2];
  const double w = a.n * b.n / n, f = b.n / n;
  a.s[0] += b.s[0] + d0 * d0 * w; a.s[1] += b.s[1] + d0 * d1 * w; a.s[2] += b.s[2] + d0 * d2 * w;
  a.s[3] += b.s[3] + d1 * d1 * w; a.s[4] += b.s[4] + d1 * d2 * w; a.s[5] += b.s[5] + d2 * d2 * w;
  a.m[0] += d0 * f; a.m[1] += d1 * f; a.m[2] += d2 * f;
  a.n = n;
}

// One plane rotation that diagonalises the 2x2 block (app apq; apq aqq): the closed form
// of a single Jacobi step.  Returns the rotated diagonal in place and (c, s).
__device__ __forceinline__ void sym2(double &app, double &aqq, double apq, double &c, double &s) {
  const double theta = (aqq - app) / (2.0 * apq);
  const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  c = 1.0 / sqrt(t * t + 1.0);
  s = t * c;
  app -= t * apq;
  aqq += t * apq;
}

// Closed-form eigen-decomposition of a symmetric 3x3 (a = xx,xy,xz,yy,yz,zz).
// w[k] / column k of V are eigenpairs; when the matrix decouples exactly (an axis whose
// off-diagonals are both 0.0) the pairs stay in axis order so that exact ties (e.g. two
// zero eigenvalues) are broken the same way the reference's strict-< chain breaks them on
// a solver that leaves a diagonal matrix untouched.  Otherwise: trigonometric solution of
// the characteristic cubic, eigenvalues ascending, eigenvectors by the largest cross
// product of rows of (A - wI).
__device__ void eig3_sym(const double a[6], double w[3], double V[3][3]) {
  const double a00 = a[0], a01 = a[1], a02 = a[2], a11 = a[3], a12 = a[4], a22 = a[5];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  w[0] = a00; w[1] = a11; w[2] = a22;
  const int nz = (a01 != 0.0) + (a02 != 0.0) + (a12 != 0.0);
  if (nz == 0) return;
  if (nz == 1) {
    double c, s;
    if (a01 != 0.0) {
      sym2(w[0], w[1], a01, c, s);
      V[0][0] = c; V[1][0] = -s; V[0][1] = s; V[1][1] = c;
    } else if (a02 != 0.0) {
      sym2(w[0], w[2], a02, c, s);
      V[0][0] = c; V[2][0] = -s; V[0][2] = s; V[2][2] = c;
    } else {
      sym2(w[1], w[2], a12, c, s);
      V[1][1] = c; V[2][1] = -s; V[1][2] = s; V[2][2] = c;
    }
    return;
  }
  // general case, scaled to unit max-norm
  double sc = fmax(fmax(fabs(a00), fabs(a11)), fabs(a22));
  sc = fmax(sc, fmax(fmax(fabs(a01), fabs(a02)), fabs(a12)));
  const double inv = 1.0 / sc;
  const double b00 = a00 * inv, b01 = a01 * inv, b02 = a02 * inv, b11 = a11 * inv, b12 = a12 * inv, b22 = a22 * inv;
  const double q = (b00 + b11 + b22) / 3.0;
  const double c00 = b00 - q, c11 = b11 - q, c22 = b22 - q;
  const double p2 = c00 * c00 + c11 * c11 + c22 * c22 + 2.0 * (b01 * b01 + b02 * b02 + b12 * b12);
  const double p = sqrt(p2 / 6.0);
  const double ip = 1.0 / p;
  const double d00 = c00 * ip, d11 = c11 * ip, d22 = c22 * ip, d01 = b01 * ip, d02 = b02 * ip, d12 = b12 * ip;
  double r = 0.5 * (d00 * (d11 * d22 - d12 * d12) - d01 * (d01 * d22 - d12 * d02) + d02 * (d01 * d12 - d11 * d02));
  r = fmin(1.0, fmax(-1.0, r));
  const double phi = acos(r) / 3.0;
  const double e_max = q + 2.0 * p * cos(phi);
  const double e_min = q + 2.0 * p * cos(phi + 2.0943951023931954923);
  const double e_mid = 3.0 * q - e_max - e_min;
  const double ev[3] = {e_min, e_mid, e_max};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double l = ev[k];
    const double r0[3] = {b00 - l, b01, b02}, r1[3] = {b01, b11 - l, b12}, r2[3] = {b02, b12, b22 - l};
    double x0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
    double x1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
    double x2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double n0 = x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2];
    const double n1 = x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2];
    const double n2 = x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2];
    double bx = x0[0], by = x0[1], bz = x0[2], bn = n0;
    if (n1 > bn) { bx = x1[0]; by = x1[1]; bz = x1[2]; bn = n1; }
    if (n2 > bn) { bx = x2[0]; by = x2[1]; bz = x2[2]; bn = n2; }
    if (bn > 0.0) {
      const double in = rsqrt(bn);
      V[0][k] = bx * in; V[1][k] = by * in; V[2][k] = bz * in;
    } else {  // triple eigenvalue: any basis
      V[0][k] = (k == 0); V[1][k] = (k == 1); V[2][k] = (k == 2);
    }
    w[k] = l * sc;
  }
}

// Build and store one voxel record.
__device__ void finalize_voxel(gndt_voxel *table, u32 slot, u64 vkey, u32 first, const Moments &mo,
                               const DevParams &P, u32 *err) {
  if (slot >= P.max_voxels) { atomicOr(err, kErrCapacity); return; }
  const int cx = (int)(u32)(vkey >> 32) - kIdxBias, cy = (int)((u32)(vkey >> 16) & 0xFFFFu) - kIdxBias,
            cz = (int)((u32)vkey & 0xFFFFu) - kIdxBias;
  float f[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) f[i] = 0.f;
  u32 *u = reinterpret_cast<u32 *>(f);
  u[0] = (u32)signed_index(cx); u[1] = (u32)signed_index(cy); u[2] = (u32)signed_index(cz);
  const u32 count = (u32)mo.n;
  u[3] = count; u[4] = first;
  if ((int)count >= P.min_points) {
    f[5] = (float)mo.m[0]; f[6] = (float)mo.m[1]; f[7] = (float)mo.m[2];
    double a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      float s = (float)mo.s[k];
      if (P.normalize_cov) s = __fdiv_rn(s, (float)count);
      f[8 + k] = s;
      a[k] = (double)s;  // the reference's solver sees the binary32 matrix (map2D.h:111)
    }
    double w[3], V[3][3];
    eig3_sym(a, w, V);
    // OcNode::countRoughNormal's strict-< chain on the solver's diagonal (map2D.h:114-130)
    const float e0 = (float)w[0], e1 = (float)w[1], e2 = (float)w[2];
    int k;
    if (e0 < e1) k = (e0 < e2) ? 0 : 2; else k = (e1 < e2) ? 1 : 2;
    float rough = (float)w[k];
    if (rough == 0.f) rough = 0.01f;  // map2D.h:131-132
    double s0 = w[0], s1 = w[1], s2 = w[2], t;
    if (s0 > s1) { t = s0; s0 = s1; s1 = t; }
    if (s1 > s2) { t = s1; s1 = s2; s2 = t; }
    if (s0 > s1) { t = s0; s0 = s1; s1 = t; }
    f[14] = (float)s0; f[15] = (float)s1; f[16] = (float)s2;
    f[17] = (float)V[0][k]; f[18] = (float)V[1][k]; f[19] = (float)V[2][k];
    f[20] = rough;
    u[21] = GNDT_F_FITTED;
  }
  float4 *dst = reinterpret_cast<float4 *>(table + slot);
#pragma unroll
  for (int i = 0; i < 6; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}

struct RedSmem {
  float4 pts[kRedTile];
  u64 key[kRedTile];
  unsigned short run_start[kRedTile + 2];
  unsigned short long_list[kMaxLong];
  u32 seg_cnt[kRedItems][8];
  double red[8][6];
  double bc[3];
  u64 prev_key, next_key;
  u32 n_long, tile_id, n_runs, vox_base;
};

__device__ __forceinline__ u64 point_key(const float4 &p, const float o[3], const DevParams &P) {
  int cx, cy, cz;
  point_indices(p.x, p.y, p.z, o, P.grid_len, P.z_len, cx, cy, cz);
  return voxel_key(cx, cy, cz);
}

// K3: one CTA per tile of 2048 sorted points.
__global__ void __launch_bounds__(kRedThreads)
reduce_kernel(Ctl *ctl, const float4 *buf_a, const float4 *buf_b, gndt_voxel *table, TileCarry *carry,
              u32 *tile_state, DevParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RedSmem &S = *reinterpret_cast<RedSmem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { S.tile_id = atomicAdd(&ctl->ticket[6], 1u); S.n_long = 0; }
  __syncthreads();
  const int tile = (int)S.tile_id;
  const size_t M = (size_t)ctl->n_valid;
  const size_t base = (size_t)tile * kRedTile;
  if (base >= M) return;
  const int cnt = (int)min((size_t)kRedTile, M - base);
  const float4 *src = ((ctl->n_passes - 1) & 1) ? buf_b : buf_a;  // pass p writes A when p is even
  const float o[3] = {ctl->origin[0], ctl->origin[1], ctl->origin[2]};

  // ---- stage the tile, one voxel key per point
#pragma unroll
  for (int k = 0; k < kRedItems; ++k) {
    const int i = k * kRedThreads + tid;
    if (i < cnt) {
      const float4 p = ld_stream(src + base + i);
      S.pts[i] = p;
      S.key[i] = point_key(p, o, P);
    }
  }
  if (tid == 0) S.prev_key = (base > 0) ? point_key(src[base - 1], o, P) : ~0ull;
  if (tid == 32) S.next_key = (base + cnt < M) ? point_key(src[base + cnt], o, P) : ~0ull;
  __syncthreads();

  // ---- run heads (position 0 always starts a run of this tile)
  u32 ball[kRedItems];
#pragma unroll
  for (int k = 0; k < kRedItems; ++k) {
    const int i = k * kRedThreads + tid;
    const bool head = (i < cnt) && (i == 0 || S.key[i] != S.key[i - 1]);
    ball[k] = __ballot_sync(0xffffffffu, head);
    if (lane == 0) S.seg_cnt[k][warp] = __popc(ball[k]);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the 64 (k,warp) counts, two per lane
    u32 c0 = (&S.seg_cnt[0][0])[2 * lane], c1 = (&S.seg_cnt[0][0])[2 * lane + 1];
    u32 inc = c0 + c1;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
      u32 t = __shfl_up_sync(0xffffffffu, inc, o2);
      if (lane >= o2) inc += t;
    }
    const u32 exc = inc - (c0 + c1);
    (&S.seg_cnt[0][0])[2 * lane] = exc;
    (&S.seg_cnt[0][0])[2 * lane + 1] = exc + c0;
    if (lane == 31) S.n_runs = inc;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kRedItems; ++k) {
    const int i = k * kRedThreads + tid;
    if (ball[k] & (1u << lane)) S.run_start[S.seg_cnt[k][warp] + __popc(ball[k] & ((1u << lane) - 1u))] = (unsigned short)i;
  }
  const int n_runs = (int)S.n_runs;
  const bool first_is_head = S.prev_key != S.key[0];
  const bool last_complete = S.next_key != S.key[cnt - 1];
  if (tid == 0) {
    S.run_start[n_runs] = (unsigned short)cnt;
    const u32 heads = (u32)n_runs - (first_is_head ? 0u : 1u);
    const u32 vb = lookback_u32(tile_state + tile, tile, 1, heads, &ctl->err);
    S.vox_base = vb;
    if (base + cnt == M) ctl->n_voxels = vb + heads;
  }
  __syncthreads();
  const u32 vox_base = S.vox_base;
  TileCarry *my_carry = carry + tile;

  auto emit = [&](int j, const Moments &mo) {
    const int s = S.run_start[j];
    const bool cont = (j == 0) && !first_is_head;
    const bool open_end = (j == n_runs - 1) && !last_complete;
    const u32 slot = vox_base + (u32)j - (first_is_head ? 0u : 1u);
    if (cont) {
      my_carry->lead = mo;
      atomicOr(&my_carry->flags, kCarryHasLead | (open_end ? kCarryLeadContinues : 0u));
    } else if (open_end) {
      my_carry->tail = mo;
      my_carry->tail_slot = slot;
      my_carry->tail_first = __float_as_uint(S.pts[s].w);
      my_carry->tail_key = S.key[s];
      atomicOr(&my_carry->flags, kCarryHasTail);
    } else {
      finalize_voxel(table, slot, S.key[s], __float_as_uint(S.pts[s].w), mo, P, &ctl->err);
    }
  };

  // ---- short runs: one thread per run, two passes over shared memory
  for (int j = tid; j < n_runs; j += kRedThreads) {
    const int s = S.run_start[j], e = S.run_start[j + 1];
    if (e - s > kLongRun) {
      S.long_list[atomicAdd(&S.n_long, 1u)] = (unsigned short)j;
      continue;
    }
    Moments mo;
    double sx = 0, sy = 0, sz = 0;
    for (int i = s; i < e; ++i) {
      const float4 p = S.pts[i];
      sx += (double)p.x; sy += (double)p.y; sz += (double)p.z;
    }
    mo.n = (double)(e - s);
    mo.m[0] = sx / mo.n; mo.m[1] = sy / mo.n; mo.m[2] = sz / mo.n;
    double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
    for (int i = s; i < e; ++i) {
      const float4 p = S.pts[i];
      const double dx = (double)p.x - mo.m[0], dy = (double)p.y - mo.m[1], dz = (double)p.z - mo.m[2];
      xx += dx * dx; xy += dx * dy; xz += dx * dz; yy += dy * dy; yz += dy * dz; zz += dz * dz;
    }
    mo.s[0] = xx; mo.s[1] = xy; mo.s[2] = xz; mo.s[3] = yy; mo.s[4] = yz; mo.s[5] = zz;
    emit(j, mo);
  }
  __syncthreads();

  // ---- long runs (heavy voxels): the whole CTA reduces each one
  const int n_long = (int)S.n_long;
  for (int l = 0; l < n_long; ++l) {
    const int j = S.long_list[l];
    const int s = S.run_start[j], e = S.run_start[j + 1];
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = s + tid; i < e; i += kRedThreads) {
      const float4 p = S.pts[i];
      a0 += (double)p.x; a1 += (double)p.y; a2 += (double)p.z;
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { S.red[warp][0] = a0; S.red[warp][1] = a1; S.red[warp][2] = a2; }
    __syncthreads();
    if (tid < 3) {
      double t = 0;
      for (int w2 = 0; w2 < 8; ++w2) t += S.red[w2][tid];
      S.bc[tid] = t / (double)(e - s);
    }
    __syncthreads();
    const double m0 = S.bc[0], m1 = S.bc[1], m2 = S.bc[2];
    double q[6] = {0, 0, 0, 0, 0, 0};
    for (int i = s + tid; i < e; i += kRedThreads) {
      const float4 p = S.pts[i];
      const double dx = (double)p.x - m0, dy = (double)p.y - m1, dz = (double)p.z - m2;
      q[0] += dx * dx; q[1] += dx * dy; q[2] += dx * dz; q[3] += dy * dy; q[4] += dy * dz; q[5] += dz * dz;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = warp_sum(q[k]);
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < 6; ++k) S.red[warp][k] = q[k];
    __syncthreads();
    if (tid == 0) {
      Moments mo;
      mo.n = (double)(e - s);
      mo.m[0] = m0; mo.m[1] = m1; mo.m[2] = m2;
      for (int k = 0; k < 6; ++k) {
        double t = 0;
        for (int w2 = 0; w2 < 8; ++w2) t += S.red[w2][k];
        mo.s[k] = t;
      }
      emit(j, mo);
    }
    __syncthreads();
  }
}

// K3b: voxels whose run straddles tiles: chain the partial moments in tile order.
__global__ void fixup_kernel(Ctl *ctl, gndt_voxel *table, const TileCarry *carry, DevParams P) {
  const size_t M = (size_t)ctl->n_valid;
  const int n_tiles = (int)((M + kRedTile - 1) / kRedTile);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
    if (!(carry[t].flags & kCarryHasTail)) continue;
    Moments mo = carry[t].tail;
    for (int u = t + 1; u < n_tiles && (carry[u].flags & kCarryHasLead); ++u) {
      merge_moments(mo, carry[u].lead);
      if (!(carry[u].flags & kCarryLeadContinues)) break;
    }
    finalize_voxel(table, carry[t].tail_slot, carry[t].tail_key, carry[t].tail_first, mo, P, &ctl->err);
  }
}

}  // namespace gndt
