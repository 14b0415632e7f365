// gndt_update.cuh — streaming fusion of one scan into the resident map (gndt_update).
//
// Replaces changeCallback + TwoDmap::change2DMap (reference src/receiver.cpp:179-212,
// include/map2D.h:672-822).  That path is dead code in the reference (uniformDivision never
// fills the lPoints lists change2DMap consumes, its subscription is commented out) and its
// pooled-covariance formula mixes scatters with sample covariances and uses an integer
// M*N/(M+N) (map2D.h:777-780); the contract here is batch equivalence: after any number of
// updates the map equals one build over the concatenation of all clouds.  The resident
// state is the sorted table of raw binary64 moments; a scan is reduced to its own sorted
// moments table by the normal front end and merged with Chan's update (exact in n, mean and
// centred scatter up to binary64 rounding).
#pragma once
#include "gndt_device.cuh"
#include "gndt_reduce.cuh"

namespace gndt {

// persistent (never memset) totals over all clouds fused into the resident map
struct Totals {
  u64 n_points, n_valid, n_dropped, n_outside;
};

__global__ void totals_kernel(Totals *tot, const Ctl *ctl, u64 n_points, int reset) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x || blockIdx.x) return;
  if (reset) { tot->n_points = 0; tot->n_valid = 0; tot->n_dropped = 0; tot->n_outside = 0; }
  tot->n_points += n_points; tot->n_valid += ctl->n_valid; tot->n_dropped += ctl->n_dropped; tot->n_outside += ctl->n_outside;
}

__device__ __forceinline__ u32 lower_bound_key(const VoxMoments *t, u32 n, u64 key) {
  u32 lo = 0, hi = n;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (t[mid].key < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ u32 lower_bound_u64(const u64 *t, u32 n, u64 key) {
  u32 lo = 0, hi = n;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (t[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// U1: one thread per scan voxel: fold it into the resident voxel with the same key, or
// flag it as new.  At most one scan voxel maps to a resident voxel, so no atomics.
__global__ void update_match_kernel(const Ctl *scan_ctl, VoxMoments *res, u32 n_res, const VoxMoments *scan, u32 *is_new) {
  const u32 n_scan = scan_ctl->n_voxels;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n_scan; j += gridDim.x * blockDim.x) {
    const VoxMoments s = scan[j];
    const u32 i = lower_bound_key(res, n_res, s.key);
    if (i < n_res && res[i].key == s.key) {
      VoxMoments *r = res + i;
      Moments a, b;
      a.n = (double)r->count; b.n = (double)s.count;
#pragma unroll
      for (int k = 0; k < 3; ++k) { a.m[k] = r->m[k]; b.m[k] = s.m[k]; }
#pragma unroll
      for (int k = 0; k < 6; ++k) { a.s[k] = r->s[k]; b.s[k] = s.s[k]; }
      merge_moments(a, b);  // resident points came first in the concatenated cloud
      r->count = (u32)a.n;
      r->first = min(r->first, s.first);
#pragma unroll
      for (int k = 0; k < 3; ++k) r->m[k] = a.m[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) r->s[k] = a.s[k];
      is_new[j] = 0;
    } else {
      is_new[j] = 1;
    }
  }
}

// U2: exclusive scan of is_new (single CTA, chunked; scans are small by nature) and the
// compacted keys of the new voxels.
__global__ void __launch_bounds__(1024) update_compact_kernel(Ctl *scan_ctl, const VoxMoments *scan, const u32 *is_new,
                                                              u32 *new_pos, u64 *new_keys, u32 *n_new_out) {
  __shared__ u32 warp_sums[32];
  __shared__ u32 carry, chunk_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 n_scan = scan_ctl->n_voxels;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (u32 base = 0; base < n_scan; base += 1024) {
    const u32 j = base + tid;
    const u32 f = (j < n_scan) ? is_new[j] : 0u;
    u32 inc = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const u32 w = warp_sums[lane];
      u32 wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sums[lane] = wi - w;  // exclusive offset of each warp inside the chunk
      if (lane == 31) chunk_total = wi;
    }
    __syncthreads();
    const u32 pos = carry + warp_sums[warp] + inc - f;
    if (j < n_scan) {
      new_pos[j] = pos;
      if (f) new_keys[pos] = scan[j].key;
    }
    __syncthreads();
    if (tid == 0) carry += chunk_total;
    __syncthreads();
  }
  if (tid == 0) *n_new_out = carry;
}

// U3a: resident voxel i moves to i + (number of new keys below it).
__global__ void update_merge_resident_kernel(const VoxMoments *res, u32 n_res, const u64 *new_keys, const u32 *n_new_p,
                                             VoxMoments *out) {
  const u32 n_new = *n_new_p;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_res; i += gridDim.x * blockDim.x) {
    const VoxMoments v = res[i];
    out[i + lower_bound_u64(new_keys, n_new, v.key)] = v;
  }
}

// U3b: new voxel with rank r moves to r + (number of resident keys below it); the last
// thread-independent bit: the merged size goes into the control block for the back end.
__global__ void update_merge_new_kernel(Ctl *ctl, const VoxMoments *res, u32 n_res, const VoxMoments *scan,
                                        const u32 *is_new, const u32 *new_pos, const u32 *n_new_p, VoxMoments *out,
                                        u32 max_voxels) {
  const u32 n_scan = ctl->n_voxels_scan;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n_scan; j += gridDim.x * blockDim.x) {
    if (!is_new[j]) continue;
    const VoxMoments v = scan[j];
    const u32 dst = new_pos[j] + lower_bound_key(res, n_res, v.key);
    if (dst < max_voxels) out[dst] = v; else atomicOr(&ctl->err, kErrCapacity);
  }
}

// Prepare the control block for a back-end run over a table of `n_res + n_new` voxels.
__global__ void update_prepare_backend_kernel(Ctl *ctl, u32 n_res, const u32 *n_new_p) {
  if (threadIdx.x || blockIdx.x) return;
  ctl->n_voxels_scan = ctl->n_voxels;
  ctl->n_voxels = n_res + *n_new_p;
  ctl->ticket[7] = 0;
  ctl->n_columns = 0; ctl->n_slopes = 0; ctl->n_fitted = 0;
}

}  // namespace gndt
