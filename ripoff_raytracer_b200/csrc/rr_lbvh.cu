// rr_lbvh.cu -- GPU LBVH builder (Morton codes -> radix sort -> Karras
// hierarchy -> atomic-flag refit -> traversal packing), one hierarchy per
// segment (mesh) of a primitive array, all segments built in one batch.
//
// Replaces the reference's host SAH builder (src/readobj.hpp:96-267: recursive,
// 5 candidate planes x 3 axes, in-place partition of triangleList).  The build
// is bit-exact against oracle/rr_oracle.c (lbvh_build_segment): identical keys,
// identical sorted order (stable in the uploaded index), identical topology and
// boxes -- tests/test_gpu_parity.py::test_lbvh_build_order_is_bit_exact (and ::test_full_size_configs_are_bit_exact).
//
// Every kernel here is HBM/latency-bound integer and min/max work; nothing is
// GEMM-shaped.  Compiled with -fmad=false (the key quantisation must round
// exactly like the CPU statement).
#include <math.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "rr_internal.h"

namespace rr {

#define RR_CK(x)                        \
  do {                                  \
    cudaError_t e_ = (x);               \
    if (e_ != cudaSuccess) return e_;   \
  } while (0)

static inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// ---- primitive boxes ------------------------------------------------------
__global__ void k_tri_boxes(const rr_triangle* __restrict__ tris, uint64_t n, float* __restrict__ box) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* t = reinterpret_cast<const float4*>(tris + i);
  float4 a = __ldg(t), b = __ldg(t + 1), c = __ldg(t + 2);
  float* o = box + 6 * i;
  o[0] = fminf(fminf(a.x, b.x), c.x);
  o[1] = fminf(fminf(a.y, b.y), c.y);
  o[2] = fminf(fminf(a.z, b.z), c.z);
  o[3] = fmaxf(fmaxf(a.x, b.x), c.x);
  o[4] = fmaxf(fmaxf(a.y, b.y), c.y);
  o[5] = fmaxf(fmaxf(a.z, b.z), c.z);
}

__global__ void k_sphere_boxes(const rr_sphere* __restrict__ sph, uint64_t n, float* __restrict__ box) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* s = reinterpret_cast<const float4*>(sph + i);
  float4 c = __ldg(s);
  float r = __ldg(reinterpret_cast<const float*>(s + 1));
  float* o = box + 6 * i;
  o[0] = c.x - r; o[1] = c.y - r; o[2] = c.z - r;
  o[3] = c.x + r; o[4] = c.y + r; o[5] = c.z + r;
}

cudaError_t launch_tri_boxes(const rr_triangle* d_tris, uint64_t n, float* d_box, cudaStream_t s) {
  if (!n) return cudaSuccess;
  k_tri_boxes<<<grid_for(n, 256), 256, 0, s>>>(d_tris, n, d_box);
  return cudaGetLastError();
}
cudaError_t launch_sphere_boxes(const rr_sphere* d_sph, uint64_t n, float* d_box, cudaStream_t s) {
  if (!n) return cudaSuccess;
  k_sphere_boxes<<<grid_for(n, 256), 256, 0, s>>>(d_sph, n, d_box);
  return cudaGetLastError();
}

// ---- segments ---------------------------------------------------------------
// Largest s with first[s] <= i, or -1.
__device__ __forceinline__ int seg_search(const uint32_t* __restrict__ first, int n_segs, uint32_t i) {
  int lo = 0, hi = n_segs;  // first[lo-1] <= i < first[hi]
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(first + mid) <= i) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_seg_init(int* seg_box_ord, uint32_t n_segs) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_segs * 6) return;
  seg_box_ord[i] = (i % 6) < 3 ? f2ord(INFINITY) : f2ord(-INFINITY);
}

// seg_id[i] = segment of prim i (n_segs when uncovered); reduces the segment boxes.
__global__ void k_seg_assign(const float* __restrict__ box, uint64_t n_total, const uint32_t* __restrict__ seg_first,
                             const uint32_t* __restrict__ seg_count, int n_segs, uint32_t* __restrict__ seg_id,
                             int* __restrict__ seg_box_ord) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i < n_total;
  int s = -1;
  float b[6];
  if (valid) {
    s = seg_search(seg_first, n_segs, (uint32_t)i);
    if (s >= 0 && (uint32_t)i - seg_first[s] >= seg_count[s]) s = -1;
    seg_id[i] = s >= 0 ? (uint32_t)s : (uint32_t)n_segs;
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = box[6 * i + k];
  }
  // warp-uniform segment: shuffle-reduce, one lane issues the atomics
  unsigned full = 0xffffffffu;
  int s0 = __shfl_sync(full, s, 0);
  bool uniform = __all_sync(full, valid && s == s0 && s >= 0);
  if (uniform) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      float v = b[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        float o = __shfl_xor_sync(full, v, off);
        v = k < 3 ? fminf(v, o) : fmaxf(v, o);
      }
      b[k] = v;
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        atomicMin(seg_box_ord + 6 * s + k, f2ord(b[k]));
        atomicMax(seg_box_ord + 6 * s + 3 + k, f2ord(b[3 + k]));
      }
    }
  } else if (valid && s >= 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atomicMin(seg_box_ord + 6 * s + k, f2ord(b[k]));
      atomicMax(seg_box_ord + 6 * s + 3 + k, f2ord(b[3 + k]));
    }
  }
}

__global__ void k_seg_decode(const int* seg_box_ord, float* seg_box, uint32_t n_segs) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_segs * 6) return;
  seg_box[i] = ord2f(seg_box_ord[i]);
}

// ---- Morton keys -------------------------------------------------------------
__device__ __forceinline__ uint64_t expand21(uint32_t v) {
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ uint32_t quant21(float c, float lo, float ext) {
  if (!(ext > 0.0f)) return 0u;
  float q = (c - lo) / ext;
  float g = q * 2097152.0f;
  if (!(g > 0.0f)) return 0u;
  if (g >= 2097151.0f) return 2097151u;
  return (uint32_t)g;
}

__global__ void k_morton(const float* __restrict__ box, uint64_t n_total, const uint32_t* __restrict__ seg_id,
                         const float* __restrict__ seg_box, uint32_t n_segs, uint64_t* __restrict__ keys,
                         uint32_t* __restrict__ vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  uint32_t s = seg_id[i];
  uint64_t code = 0;
  if (s < n_segs) {
    const float* sb = seg_box + 6 * s;
    uint32_t g[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float c = (box[6 * i + a] + box[6 * i + 3 + a]) * 0.5f;
      g[a] = quant21(c, sb[a], sb[3 + a] - sb[a]);
    }
    code = (expand21(g[0]) << 2) | (expand21(g[1]) << 1) | expand21(g[2]);
  }
  keys[i] = code;
  vals[i] = (uint32_t)i;
}

// ---- stable LSD radix sort, 8 bits per pass ---------------------------------
// One pass = histogram per tile, exclusive scan over (digit, tile) -- per digit value across the tiles, then across the
// digit values inside the scatter kernel --, stable scatter.
// A tile is processed by 8 warps; warp w owns the w-th contiguous eighth of the
// tile, walks it 32 consecutive items at a time and ranks equal digits with
// __match_any_sync, so equal keys keep their input order.
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;  // per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

template <bool SEG_DIGIT>
__device__ __forceinline__ uint32_t sort_digit(uint64_t key, uint32_t val, const uint32_t* __restrict__ seg_id, int shift) {
  if (SEG_DIGIT) return (__ldg(seg_id + val) >> shift) & 0xffu;
  return (uint32_t)(key >> shift) & 0xffu;
}

template <bool SEG_DIGIT>
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                            const uint32_t* __restrict__ seg_id, uint64_t n, int shift,
                                                            uint32_t* __restrict__ hist, uint32_t n_tiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
#pragma unroll
  for (int k = 0; k < SORT_ITEMS; ++k) {
    uint64_t i = base + (uint64_t)k * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[sort_digit<SEG_DIGIT>(keys[i], vals[i], seg_id, shift)], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// Exclusive scan of the per-tile counters of ONE digit value per block (256 blocks run side by side), in place; the
// total of every digit value goes to totals[digit].  The scan ACROSS the 256 digit values is done by the scatter
// kernel itself (256 numbers, one block-wide scan), so a radix pass is three launches and none of them is serial.
__global__ void __launch_bounds__(1024) k_scan_bins(uint32_t* __restrict__ hist, uint32_t n_tiles, uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  uint32_t* data = hist + (uint64_t)blockIdx.x * n_tiles;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n_tiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_tiles ? data[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= (unsigned)off) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      const uint32_t w = warp_sums[lane];
      uint32_t ws = w;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, ws, off);
        if (lane >= (unsigned)off) ws += y;
      }
      warp_sums[lane] = ws - w;  // exclusive
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    if (i < n_tiles) data[i] = carry + warp_sums[wid] + (x - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_sums[wid] + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

template <bool SEG_DIGIT>
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                               uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                               const uint32_t* __restrict__ seg_id, uint64_t n, int shift,
                                                               const uint32_t* __restrict__ offsets, uint32_t n_tiles,
                                                               const uint32_t* __restrict__ totals) {
  constexpr int WARPS = SORT_THREADS / 32;
  constexpr int ROUNDS = SORT_TILE / WARPS / 32;
  static_assert(SORT_THREADS == 256, "one thread per digit value");
  __shared__ uint32_t wh[WARPS][256];
  __shared__ uint32_t digit_base_warp[WARPS];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < WARPS * 256; k += SORT_THREADS) (&wh[0][0])[k] = 0;
  __syncthreads();
  const uint64_t chunk = (uint64_t)blockIdx.x * SORT_TILE + (uint64_t)w * (ROUNDS * 32);
  uint64_t key[ROUNDS];
  uint32_t val[ROUNDS];
  uint32_t dig[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint64_t i = chunk + (uint64_t)r * 32 + lane;
    if (i < n) {
      key[r] = keys_in[i];
      val[r] = vals_in[i];
      dig[r] = sort_digit<SEG_DIGIT>(key[r], val[r], seg_id, shift);
      atomicAdd(&wh[w][dig[r]], 1u);
    } else {
      key[r] = 0; val[r] = 0;
      dig[r] = 0x100u + lane;  // matches nobody
    }
  }
  // first output position of every digit value: exclusive scan of the 256 totals (thread = digit value)
  uint32_t digit_base;
  {
    const uint32_t tot = totals[threadIdx.x];
    uint32_t x = tot;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= (unsigned)off) x += y;
    }
    if (lane == 31) digit_base_warp[w] = x;
    __syncthreads();  // (also orders the per-warp histograms above)
    uint32_t before = 0;
#pragma unroll
    for (int ww = 0; ww < WARPS; ++ww) before += ww < (int)w ? digit_base_warp[ww] : 0u;
    digit_base = before + x - tot;
  }
  {
    const unsigned bin = threadIdx.x;
    uint32_t running = digit_base + offsets[(uint64_t)bin * n_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < WARPS; ++ww) {
      uint32_t c = wh[ww][bin];
      wh[ww][bin] = running;
      running += c;
    }
  }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    unsigned m = __match_any_sync(0xffffffffu, dig[r]);
    bool ok = dig[r] < 0x100u;
    unsigned rank = __popc(m & lt);
    uint32_t pos = 0;
    if (ok) pos = wh[w][dig[r]] + rank;
    __syncwarp();
    if (ok && rank == 0) wh[w][dig[r]] += __popc(m);
    __syncwarp();
    if (ok) {
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

// ---- hierarchy (Karras 2012) -------------------------------------------------
__device__ __forceinline__ int key_delta(const uint64_t* __restrict__ codes, int64_t n, int64_t i, int64_t j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = codes[i], b = codes[j];
  if (a == b) return 64 + __clz((int)((uint32_t)i ^ (uint32_t)j));
  return __clzll((long long)(a ^ b));
}

__global__ void k_karras(const uint64_t* __restrict__ codes_all, uint64_t n, const uint32_t* __restrict__ seg_sfirst,
                         const uint32_t* __restrict__ seg_count, int n_segs, int32_t* __restrict__ left,
                         int32_t* __restrict__ right, int32_t* __restrict__ parent, int32_t* __restrict__ leaf_parent,
                         uint32_t* __restrict__ range_first, uint32_t* __restrict__ range_count) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  int s = seg_search(seg_sfirst, n_segs, (uint32_t)g);
  // zero-count segments share their sfirst with the next one: seg_search returns the last of them
  const uint32_t first = seg_sfirst[s];
  const int64_t N = seg_count[s];
  const int64_t i = (int64_t)g - first;
  if (i >= N - 1) return;
  const uint64_t* codes = codes_all + first;
  int d = (key_delta(codes, N, i, i + 1) - key_delta(codes, N, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = key_delta(codes, N, i, i - d);
  int64_t lmax = 2;
  while (key_delta(codes, N, i, i + lmax * d) > dmin) lmax *= 2;
  int64_t l = 0;
  for (int64_t t = lmax / 2; t >= 1; t /= 2)
    if (key_delta(codes, N, i, i + (l + t) * d) > dmin) l += t;
  int64_t j = i + l * d;
  int dnode = key_delta(codes, N, i, j);
  int64_t sp = 0;
  for (int64_t t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (key_delta(codes, N, i, i + (sp + t) * d) > dnode) sp += t;
    if (t <= 1) break;
  }
  int64_t gamma = i + sp * d + (d < 0 ? -1 : 0);
  int64_t lo = i < j ? i : j, hi = i < j ? j : i;
  int32_t me = (int32_t)(first + i);
  int32_t L, R;
  if (lo == gamma) { L = ~(int32_t)(first + gamma); leaf_parent[first + gamma] = me; }
  else { L = (int32_t)(first + gamma); parent[first + gamma] = me; }
  if (hi == gamma + 1) { R = ~(int32_t)(first + gamma + 1); leaf_parent[first + gamma + 1] = me; }
  else { R = (int32_t)(first + gamma + 1); parent[first + gamma + 1] = me; }
  left[me] = L;
  right[me] = R;
  range_first[me] = (uint32_t)(first + lo);  // the sorted slots under this node: [first + lo, first + hi]
  range_count[me] = (uint32_t)(hi - lo + 1);
}

__device__ __forceinline__ void load_ref_box(int32_t ref, const uint32_t* __restrict__ order,
                                             const float* __restrict__ prim_box, const float* bounds, float* b) {
  if (ref < 0) {
    const float* p = prim_box + 6 * (uint64_t)order[~ref];
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = __ldg(p + k);
  } else {
    const float* p = bounds + 6 * (uint64_t)ref;
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = __ldcg(p + k);  // written by another thread of this launch
  }
}

// Bottom-up refit: the second thread to reach a node computes its box.  Also
// records the depth of the deepest leaf (root = 1).
__global__ void k_refit(uint64_t n, const uint32_t* __restrict__ order, const float* __restrict__ prim_box,
                        const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                        const int32_t* __restrict__ parent, const int32_t* __restrict__ leaf_parent,
                        float* bounds, unsigned int* flags, unsigned int* max_depth) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  int32_t p = leaf_parent[g];
  unsigned depth = 1;
  bool working = true;
  while (p >= 0) {
    depth++;
    if (working) {
      unsigned old = atomicAdd(flags + p, 1u);
      if (old == 0) {
        working = false;
      } else {
        float a[6], c[6];
        load_ref_box(left[p], order, prim_box, bounds, a);
        load_ref_box(right[p], order, prim_box, bounds, c);
        float* o = bounds + 6 * (uint64_t)p;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          __stcg(o + k, fminf(a[k], c[k]));
          __stcg(o + 3 + k, fmaxf(a[3 + k], c[3 + k]));
        }
        __threadfence();
      }
    }
    p = parent[p];
  }
  if (depth > 1) atomicMax(max_depth, depth);
}

// Traversal node: a 4-wide node, 128 bytes = one cache line, collapsed from the binary LBVH top-down: a wide node
// starts from the two children of a binary node and keeps replacing the inner child with the LARGEST SURFACE AREA by
// that child's two children until it has four (or only leaves are left).  Every inner child becomes a wide node of
// the next level (it keeps its binary index; binary nodes that were absorbed are never fetched).
// A subtree of at most `leaf_max` (<= 4) primitives is not descended: the primitives of an LBVH subtree are consecutive
// sorted slots, so the subtree becomes ONE leaf reference (first slot, count) with the subtree's box, and its nodes are
// never packed or fetched.  Fewer levels and a third of the node bytes for a few more primitive tests per ray.
//   q0..q2 = min.x[4] min.y[4] min.z[4]   q3..q5 = max.x[4] max.y[4] max.z[4]   q6 = ref[4]   q7 = (#children, -, -, -)
// Unused child slots hold a NaN box: every comparison of the slab test fails, no ray enters it (an inverted
// box would not do: the test orders the two planes of a slab itself).  The boxes are inflated by
// the segment's box_delta() so that culling is conservative (the closest hit then does not depend on the order
// in which a traversal visits the nodes).  ref: inner = index in the combined node array (ref_offset added),
// leaf = -((first slot << 2 | count - 1) + 2) (so that -1 is free for "pop", rr_render.cu ref_slot / ref_count).
__global__ void k_wide_roots(const uint32_t* __restrict__ seg_root, const uint32_t* __restrict__ seg_count, int n_segs,
                             int32_t* __restrict__ frontier, unsigned int* __restrict__ count) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  if (seg_count[s] >= 2u) frontier[atomicAdd(count, 1u)] = (int32_t)seg_root[s];  // the segment's root inner node
}

// ---- SAH-ordered top over the Karras subtrees (HLBVH) -----------------------------------------------------------
// The Morton splits of the LBVH are worst at the top of a large hierarchy (on the 1 M-triangle bench mesh a full SAH
// build needs 13 % fewer node visits, tools/bvh_quality_experiment.cpp).  For a large segment the Karras subtrees of
// at most `T` primitives are therefore kept as they are ("clusters": 99.7 % of the nodes at T = 512) and only the few
// thousand nodes above them are replaced: the cluster roots are listed here, the host builds a binned-SAH binary tree
// over their boxes (lbvh_sah_top, a few thousand items) and the new nodes are appended behind the Karras nodes.  The
// Karras products themselves (what rr_bvh_read returns and the build-order test compares) are not touched.
__global__ void k_find_clusters(uint64_t n, const uint32_t* __restrict__ seg_sfirst, const uint32_t* __restrict__ seg_count, int n_segs,
                                const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                const int32_t* __restrict__ parent, const uint32_t* __restrict__ range_count,
                                const float* __restrict__ bounds, const uint32_t* __restrict__ order,
                                const float* __restrict__ prim_box, uint32_t T, uint32_t min_count, uint32_t capacity,
                                int32_t* __restrict__ c_ref, uint32_t* __restrict__ c_seg, uint32_t* __restrict__ c_cnt,
                                float* __restrict__ c_box, unsigned int* __restrict__ counter) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int s = seg_search(seg_sfirst, n_segs, (uint32_t)g);
  const uint32_t first = seg_sfirst[s], N = seg_count[s];
  if (N < min_count || g - first >= N - 1) return;  // small segment, or not an inner node of it
  const uint32_t rc = range_count[g];
  auto emit = [&](int32_t ref, uint32_t cnt, const float* box) {
    const unsigned int k = atomicAdd(counter, 1u);
    if (k >= capacity) return;
    c_ref[k] = ref; c_seg[k] = (uint32_t)s; c_cnt[k] = cnt;
    for (int a = 0; a < 6; ++a) c_box[6 * (size_t)k + a] = box[a];
  };
  if (rc <= T) {
    const int32_t par = parent[g];
    if (par >= 0 && range_count[par] > T) emit((int32_t)g, rc, bounds + 6 * g);
  } else {  // a node that stays above the cut: a child that is a single primitive is a cluster of its own
    const int32_t kids[2] = {left[g], right[g]};
    for (int k = 0; k < 2; ++k)
      if (kids[k] < 0) emit(kids[k], 1u, prim_box + 6 * (size_t)order[~kids[k]]);
  }
}

namespace {
struct TopItem { int32_t ref; uint32_t cnt; float box[6]; };
struct TopNode { int32_t left, right; uint32_t cnt; float box[6]; };

inline float top_area(const float* b) {
  const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
  return dx * dy + dy * dz + dz * dx;
}
inline void top_grow(float* b, const float* o) {
  for (int a = 0; a < 3; ++a) { b[a] = std::min(b[a], o[a]); b[3 + a] = std::max(b[3 + a], o[3 + a]); }
}
inline void top_init(float* b) {
  for (int a = 0; a < 3; ++a) { b[a] = INFINITY; b[3 + a] = -INFINITY; }
}

// Binned SAH (16 bins per axis, cost = box area x primitives) over items[lo, hi); appends the inner nodes to `out`
// (children first) and returns the reference of the subtree root: an item's own reference, or node_base + index in `out`.
int32_t lbvh_sah_top(std::vector<TopItem>& items, size_t lo, size_t hi, int32_t node_base, std::vector<TopNode>& out) {
  if (hi - lo == 1) return items[lo].ref;
  constexpr int NB = 16;
  float box[6], cbox[6];
  top_init(box); top_init(cbox);
  uint32_t total = 0;
  for (size_t i = lo; i < hi; ++i) {
    top_grow(box, items[i].box);
    float c[6];
    for (int a = 0; a < 3; ++a) c[a] = c[3 + a] = 0.5f * (items[i].box[a] + items[i].box[3 + a]);
    top_grow(cbox, c);
    total += items[i].cnt;
  }
  float best = INFINITY;
  int best_axis = -1, best_bin = -1;
  for (int a = 0; a < 3; ++a) {
    const float ext = cbox[3 + a] - cbox[a];
    if (!(ext > 0.0f)) continue;
    float bb[NB][6];
    uint32_t bc[NB];
    for (int k = 0; k < NB; ++k) { top_init(bb[k]); bc[k] = 0; }
    for (size_t i = lo; i < hi; ++i) {
      const float c = 0.5f * (items[i].box[a] + items[i].box[3 + a]);
      const int k = std::min(NB - 1, (int)((c - cbox[a]) / ext * NB));
      top_grow(bb[k], items[i].box);
      bc[k] += items[i].cnt;
    }
    float la[NB], ra[NB], acc[6];
    uint32_t lc[NB], rc[NB], c = 0;
    top_init(acc);
    for (int k = 0; k < NB; ++k) { top_grow(acc, bb[k]); c += bc[k]; la[k] = c ? top_area(acc) : 0.0f; lc[k] = c; }
    top_init(acc); c = 0;
    for (int k = NB - 1; k >= 0; --k) { top_grow(acc, bb[k]); c += bc[k]; ra[k] = c ? top_area(acc) : 0.0f; rc[k] = c; }
    for (int k = 0; k + 1 < NB; ++k) {
      if (!lc[k] || !rc[k + 1]) continue;
      const float cost = la[k] * (float)lc[k] + ra[k + 1] * (float)rc[k + 1];
      if (cost < best) { best = cost; best_axis = a; best_bin = k; }
    }
  }
  size_t mid = (lo + hi) / 2;
  if (best_axis >= 0) {
    const float ext = cbox[3 + best_axis] - cbox[best_axis];
    auto it = std::partition(items.begin() + lo, items.begin() + hi, [&](const TopItem& t) {
      const float c = 0.5f * (t.box[best_axis] + t.box[3 + best_axis]);
      return std::min(NB - 1, (int)((c - cbox[best_axis]) / ext * NB)) <= best_bin;
    });
    const size_t m = (size_t)(it - items.begin());
    if (m > lo && m < hi) mid = m;
  }
  TopNode nd;
  nd.left = lbvh_sah_top(items, lo, mid, node_base, out);
  nd.right = lbvh_sah_top(items, mid, hi, node_base, out);
  nd.cnt = total;
  for (int a = 0; a < 6; ++a) nd.box[a] = box[a];
  out.push_back(nd);
  return node_base + (int32_t)out.size() - 1;
}
}  // namespace

__device__ __forceinline__ float half_area(const float* __restrict__ b) {
  const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
  return dx * dy + dy * dz + dz * dx;
}

__global__ void k_pack_wide(const int32_t* __restrict__ frontier, unsigned int n_front, int32_t* __restrict__ next,
                            unsigned int* __restrict__ next_count, const uint32_t* __restrict__ order,
                            const float* __restrict__ prim_box, const int32_t* __restrict__ left,
                            const int32_t* __restrict__ right, const float* __restrict__ bounds,
                            const uint32_t* __restrict__ seg_sfirst, int n_segs, const float* __restrict__ seg_box,
                            int32_t ref_offset, const uint32_t* __restrict__ range_first,
                            const uint32_t* __restrict__ range_count, uint32_t leaf_max, uint64_t n_karras,
                            const uint32_t* __restrict__ top_seg, float4* __restrict__ nodes) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_front) return;
  const int32_t g = frontier[i];
  int32_t kids[4] = {left[g], right[g], 0, 0};
  int cnt = 2;
  auto opens = [&](int32_t kid) { return kid >= 0 && range_count[kid] > leaf_max; };  // an inner node that stays a node
  while (cnt < 4) {
    int best = -1;
    float best_area = -1.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < cnt && opens(kids[k])) {
        const float a = half_area(bounds + 6 * (uint64_t)kids[k]);
        if (a > best_area) { best_area = a; best = k; }  // ties: the earlier child
      }
    }
    if (best < 0) break;
    const int32_t c = kids[best];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k == best) kids[k] = left[c];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k == cnt) kids[k] = right[c];
    cnt++;
  }
  const int s = (uint64_t)g >= n_karras ? (int)top_seg[(uint64_t)g - n_karras] : seg_search(seg_sfirst, n_segs, (uint32_t)g);
  float sb[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) sb[k] = __ldg(seg_box + 6 * s + k);
  const float d = box_delta(sb);
  float lo[3][4], hi[3][4];
  int32_t ref[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ref[k] = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a][k] = __int_as_float(0x7fc00000); hi[a][k] = __int_as_float(0x7fc00000); }
    if (k < cnt) {
      float bx[6];
      load_ref_box(kids[k], order, prim_box, bounds, bx);
#pragma unroll
      for (int a = 0; a < 3; ++a) { lo[a][k] = bx[a] - d; hi[a][k] = bx[3 + a] + d; }
      if (opens(kids[k])) {
        ref[k] = kids[k] + ref_offset;
        next[atomicAdd(next_count, 1u)] = kids[k];
      } else {  // one primitive, or a whole subtree of at most leaf_max primitives
        const uint32_t first = kids[k] >= 0 ? range_first[kids[k]] : (uint32_t)~kids[k];  // (top nodes never get here: count > leaf_max)
        const uint32_t count = kids[k] >= 0 ? range_count[kids[k]] : 1u;
        ref[k] = -(int32_t)(((first << 2) | (count - 1u)) + 2u);
      }
    }
  }
  float4* o = nodes + RR_NODE_QUADS * (uint64_t)g;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    o[a] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
    o[3 + a] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
  }
  o[6] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), __int_as_float(ref[2]), __int_as_float(ref[3]));
  o[7] = make_float4(__int_as_float(cnt), 0.0f, 0.0f, 0.0f);
}

// Sorted triangle arrays.  geom: (A, prim) (B-A) (C-A) -- the edge vectors are
// the same subtractions the reference does per test (src/Trace.cl:277-278).
__global__ void k_pack_tris(const rr_triangle* __restrict__ tris, const uint32_t* __restrict__ order, uint64_t n,
                            float4* __restrict__ geom, float4* __restrict__ nrm) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  uint32_t prim = order[g];
  const float4* t = reinterpret_cast<const float4*>(tris + prim);
  float4 A = __ldg(t), B = __ldg(t + 1), Cc = __ldg(t + 2);
  geom[3 * g + 0] = make_float4(A.x, A.y, A.z, __uint_as_float(prim));
  geom[3 * g + 1] = make_float4(B.x - A.x, B.y - A.y, B.z - A.z, 0.0f);
  geom[3 * g + 2] = make_float4(Cc.x - A.x, Cc.y - A.y, Cc.z - A.z, 0.0f);
  nrm[3 * g + 0] = __ldg(t + 3);
  nrm[3 * g + 1] = __ldg(t + 4);
  nrm[3 * g + 2] = __ldg(t + 5);
}

__global__ void k_pack_spheres(const rr_sphere* __restrict__ sph, const uint32_t* __restrict__ order, uint64_t n,
                               float4* __restrict__ geom) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  uint32_t prim = order[g];
  const float4* s = reinterpret_cast<const float4*>(sph + prim);
  float4 c = __ldg(s);
  float r = __ldg(reinterpret_cast<const float*>(s + 1));
  geom[g] = make_float4(c.x, c.y, c.z, r);
}

cudaError_t launch_pack_tris(const rr_triangle* d_tris, const uint32_t* d_order, uint64_t n, float4* geom, float4* nrm,
                             cudaStream_t s) {
  if (!n) return cudaSuccess;
  k_pack_tris<<<grid_for(n, 256), 256, 0, s>>>(d_tris, d_order, n, geom, nrm);
  return cudaGetLastError();
}
cudaError_t launch_pack_spheres(const rr_sphere* d_sph, const uint32_t* d_order, uint64_t n, float4* geom,
                                cudaStream_t s) {
  if (!n) return cudaSuccess;
  k_pack_spheres<<<grid_for(n, 256), 256, 0, s>>>(d_sph, d_order, n, geom);
  return cudaGetLastError();
}

__global__ void k_fill_i32(int32_t* p, uint64_t n, int32_t v) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) p[g] = v;
}

// ---- host driver ----------------------------------------------------------------
void lbvh_free(Lbvh& b) {
  dev_free(b.codes); dev_free(b.order); dev_free(b.left); dev_free(b.right); dev_free(b.parent); dev_free(b.bounds);
  dev_free(b.seg_box); dev_free(b.seg_first); dev_free(b.seg_count); dev_free(b.seg_sfirst); dev_free(b.seg_root); dev_free(b.nodes);
  b = Lbvh();
}

template <class T>
static cudaError_t dalloc(T** p, uint64_t count) {
  return dev_malloc(p, (count ? count : 1) * sizeof(T));  // cached: a re-upload of a scene of the same size allocates nothing
}

// RR_BUILD_TIMING=1 in the environment: host wall clock of the build phases on stderr (each mark synchronises the stream)
static void build_mark(cudaStream_t st, const char* what) {
  static const bool on = getenv("RR_BUILD_TIMING") != nullptr;
  static std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
  if (!on) return;
  cudaStreamSynchronize(st);
  const auto now = std::chrono::steady_clock::now();
  fprintf(stderr, "[build] %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(now - last).count());
  last = now;
}

cudaError_t lbvh_build(Lbvh& out, const float* d_prim_box, uint64_t n_total, const uint32_t* h_seg_first,
                       const uint32_t* h_seg_count, uint32_t n_segs, int32_t ref_offset, uint32_t leaf_max, uint32_t top_cluster,
                       cudaStream_t st) {
  build_mark(st, "(before the build)");
  lbvh_free(out);
  leaf_max = leaf_max < 1u ? 1u : leaf_max > RR_LEAF_MAX ? RR_LEAF_MAX : leaf_max;
  out.n_segs = n_segs;
  uint64_t n = 0;
  uint32_t* h_sfirst = (uint32_t*)malloc((n_segs ? n_segs : 1) * sizeof(uint32_t));
  for (uint32_t s = 0; s < n_segs; ++s) {
    h_sfirst[s] = (uint32_t)n;
    n += h_seg_count[s];
  }
  out.n = n;
  out.n_nodes = n;
  // room for the SAH-ordered top (top_cluster > 0): at most one new node per cluster, clusters bounded by `top_cap`
  uint64_t big = 0;
  for (uint32_t s = 0; s < n_segs; ++s)
    if (top_cluster && h_seg_count[s] >= RR_TOP_MIN_CLUSTERS * (uint64_t)top_cluster) big += h_seg_count[s];
  const uint32_t top_cap = big ? (uint32_t)std::min<uint64_t>(big, std::max<uint64_t>(4096, 16 * big / top_cluster)) : 0u;
  const uint64_t n_alloc = n + top_cap;  // node-indexed arrays
  cudaError_t err = cudaSuccess;
  uint64_t *keys_a = nullptr, *keys_b = nullptr;
  uint32_t *vals_a = nullptr, *vals_b = nullptr, *seg_id = nullptr, *hist = nullptr;
  int* seg_box_ord = nullptr;
  int32_t* leaf_parent = nullptr;
  uint32_t *range_first = nullptr, *range_count = nullptr;
  unsigned int *flags = nullptr, *d_depth = nullptr, *front_count = nullptr;
  int32_t *front_a = nullptr, *front_b = nullptr, *c_ref = nullptr;
  uint32_t *c_seg = nullptr, *c_cnt = nullptr, *top_seg = nullptr;
  float* c_box = nullptr;
  unsigned int* c_counter = nullptr;
  const uint32_t n_tiles = (uint32_t)((n_total + SORT_TILE - 1) / SORT_TILE);
#define RR_TRY(x)                   \
  do {                              \
    err = (x);                      \
    if (err != cudaSuccess) goto done; \
  } while (0)
  RR_TRY(dalloc(&out.codes, n_total));
  RR_TRY(dalloc(&out.order, n_total));
  RR_TRY(dalloc(&out.left, n_alloc));
  RR_TRY(dalloc(&out.right, n_alloc));
  RR_TRY(dalloc(&out.parent, n));
  RR_TRY(dalloc(&out.bounds, (n_alloc ? n_alloc : 1) * 6));
  RR_TRY(dalloc(&out.seg_root, n_segs));
  RR_TRY(dalloc(&out.seg_box, (uint64_t)(n_segs ? n_segs : 1) * 6));
  RR_TRY(dalloc(&out.seg_first, n_segs));
  RR_TRY(dalloc(&out.seg_count, n_segs));
  RR_TRY(dalloc(&out.seg_sfirst, n_segs));
  RR_TRY(dalloc(&out.nodes, (n_alloc ? n_alloc : 1) * RR_NODE_QUADS));
  RR_TRY(dalloc(&keys_a, n_total));
  RR_TRY(dalloc(&keys_b, n_total));
  RR_TRY(dalloc(&vals_a, n_total));
  RR_TRY(dalloc(&vals_b, n_total));
  RR_TRY(dalloc(&seg_id, n_total));
  RR_TRY(dalloc(&hist, (uint64_t)256 * (n_tiles ? n_tiles : 1) + 256));  // per-(digit, tile) counters + 256 digit totals
  RR_TRY(dalloc(&seg_box_ord, (uint64_t)(n_segs ? n_segs : 1) * 6));
  RR_TRY(dalloc(&leaf_parent, n));
  RR_TRY(dalloc(&range_first, n_alloc));
  RR_TRY(dalloc(&range_count, n_alloc));
  RR_TRY(dalloc(&flags, n));
  RR_TRY(dalloc(&d_depth, 1));
  build_mark(st, "allocations");
  if (n_segs) {
    RR_TRY(cudaMemcpyAsync(out.seg_first, h_seg_first, n_segs * 4, cudaMemcpyHostToDevice, st));
    RR_TRY(cudaMemcpyAsync(out.seg_count, h_seg_count, n_segs * 4, cudaMemcpyHostToDevice, st));
    RR_TRY(cudaMemcpyAsync(out.seg_sfirst, h_sfirst, n_segs * 4, cudaMemcpyHostToDevice, st));
    RR_TRY(cudaMemcpyAsync(out.seg_root, h_sfirst, n_segs * 4, cudaMemcpyHostToDevice, st));  // root = the Karras root unless a SAH top replaces it
  }
  RR_TRY(cudaMemsetAsync(flags, 0, (n ? n : 1) * 4, st));
  RR_TRY(cudaMemsetAsync(d_depth, 0, 4, st));
  RR_TRY(cudaMemsetAsync(out.bounds, 0, (n ? n : 1) * 24, st));
  RR_TRY(cudaMemsetAsync(out.left, 0, (n ? n : 1) * 4, st));
  RR_TRY(cudaMemsetAsync(out.right, 0, (n ? n : 1) * 4, st));
  if (n_total && n_segs) {
    k_seg_init<<<grid_for(n_segs * 6, 256), 256, 0, st>>>(seg_box_ord, n_segs);
    k_seg_assign<<<grid_for(n_total, 256), 256, 0, st>>>(d_prim_box, n_total, out.seg_first, out.seg_count, (int)n_segs,
                                                         seg_id, seg_box_ord);
    k_seg_decode<<<grid_for(n_segs * 6, 256), 256, 0, st>>>(seg_box_ord, out.seg_box, n_segs);
    k_morton<<<grid_for(n_total, 256), 256, 0, st>>>(d_prim_box, n_total, seg_id, out.seg_box, n_segs, keys_a, vals_a);
    RR_TRY(cudaGetLastError());
    // sort by key (8 passes), then stably by segment: result is ordered by (segment, key, prim)
    uint64_t *kin = keys_a, *kout = keys_b;
    uint32_t *vin = vals_a, *vout = vals_b;
    for (int pass = 0; pass < 8; ++pass) {
      k_sort_hist<false><<<n_tiles, SORT_THREADS, 0, st>>>(kin, vin, seg_id, n_total, pass * 8, hist, n_tiles);
      k_scan_bins<<<256, 1024, 0, st>>>(hist, n_tiles, hist + (uint64_t)256 * n_tiles);
      k_sort_scatter<false><<<n_tiles, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, seg_id, n_total, pass * 8, hist, n_tiles,
                                                              hist + (uint64_t)256 * n_tiles);
      uint64_t* tk = kin; kin = kout; kout = tk;
      uint32_t* tv = vin; vin = vout; vout = tv;
    }
    const bool one_full_segment = (n_segs == 1 && h_seg_first[0] == 0 && h_seg_count[0] == n_total);
    if (!one_full_segment) {
      for (int shift = 0; (n_segs >> shift) != 0; shift += 8) {
        k_sort_hist<true><<<n_tiles, SORT_THREADS, 0, st>>>(kin, vin, seg_id, n_total, shift, hist, n_tiles);
        k_scan_bins<<<256, 1024, 0, st>>>(hist, n_tiles, hist + (uint64_t)256 * n_tiles);
        k_sort_scatter<true><<<n_tiles, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, seg_id, n_total, shift, hist, n_tiles,
                                                               hist + (uint64_t)256 * n_tiles);
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
      }
    }
    RR_TRY(cudaGetLastError());
    build_mark(st, "boxes, morton, radix sort");
    RR_TRY(cudaMemcpyAsync(out.codes, kin, n_total * 8, cudaMemcpyDeviceToDevice, st));
    RR_TRY(cudaMemcpyAsync(out.order, vin, n_total * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (n) {
    k_fill_i32<<<grid_for(n, 256), 256, 0, st>>>(out.parent, n, -1);
    k_fill_i32<<<grid_for(n, 256), 256, 0, st>>>(leaf_parent, n, -1);
    k_karras<<<grid_for(n, 128), 128, 0, st>>>(out.codes, n, out.seg_sfirst, out.seg_count, (int)n_segs, out.left, out.right,
                                               out.parent, leaf_parent, range_first, range_count);
    k_refit<<<grid_for(n, 128), 128, 0, st>>>(n, out.order, d_prim_box, out.left, out.right, out.parent, leaf_parent,
                                              out.bounds, flags, d_depth);
    RR_TRY(cudaGetLastError());
    build_mark(st, "karras + refit");
    if (top_cap) {  // SAH-ordered top over the Karras subtrees of the large segments (k_find_clusters)
      RR_TRY(dalloc(&c_ref, top_cap));
      RR_TRY(dalloc(&c_seg, top_cap));
      RR_TRY(dalloc(&c_cnt, top_cap));
      RR_TRY(dalloc(&c_box, (uint64_t)top_cap * 6));
      RR_TRY(dalloc(&c_counter, 1));
      RR_TRY(cudaMemsetAsync(c_counter, 0, 4, st));
      k_find_clusters<<<grid_for(n, 128), 128, 0, st>>>(n, out.seg_sfirst, out.seg_count, (int)n_segs, out.left, out.right, out.parent,
                                                        range_count, out.bounds, out.order, d_prim_box, top_cluster,
                                                        RR_TOP_MIN_CLUSTERS * top_cluster, top_cap, c_ref, c_seg, c_cnt, c_box, c_counter);
      RR_TRY(cudaGetLastError());
      unsigned int K = 0;
      RR_TRY(cudaMemcpyAsync(&K, c_counter, 4, cudaMemcpyDeviceToHost, st));
      RR_TRY(cudaStreamSynchronize(st));
      if (K >= 2 && K <= top_cap) {  // (more clusters than room: a degenerate hierarchy, keep the Karras top)
        std::vector<int32_t> h_ref(K);
        std::vector<uint32_t> h_seg(K), h_cnt(K);
        std::vector<float> h_box((size_t)K * 6);
        RR_TRY(cudaMemcpyAsync(h_ref.data(), c_ref, K * 4, cudaMemcpyDeviceToHost, st));
        RR_TRY(cudaMemcpyAsync(h_seg.data(), c_seg, K * 4, cudaMemcpyDeviceToHost, st));
        RR_TRY(cudaMemcpyAsync(h_cnt.data(), c_cnt, K * 4, cudaMemcpyDeviceToHost, st));
        RR_TRY(cudaMemcpyAsync(h_box.data(), c_box, (size_t)K * 24, cudaMemcpyDeviceToHost, st));
        RR_TRY(cudaStreamSynchronize(st));
        std::vector<uint32_t> idx(K);
        for (unsigned int k = 0; k < K; ++k) idx[k] = k;
        // the kernel lists the clusters in atomic order: sort by (segment, reference) so that the top is the same every time
        std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return h_seg[a] != h_seg[b] ? h_seg[a] < h_seg[b] : h_ref[a] < h_ref[b]; });
        std::vector<TopNode> top;
        std::vector<uint32_t> top_seg_h, roots(h_sfirst, h_sfirst + n_segs);
        for (size_t a = 0; a < K;) {
          size_t b = a;
          std::vector<TopItem> items;
          while (b < K && h_seg[idx[b]] == h_seg[idx[a]]) {
            TopItem it;
            it.ref = h_ref[idx[b]]; it.cnt = h_cnt[idx[b]];
            for (int q = 0; q < 6; ++q) it.box[q] = h_box[6 * (size_t)idx[b] + q];
            items.push_back(it);
            ++b;
          }
          const size_t before = top.size();
          const int32_t root = lbvh_sah_top(items, 0, items.size(), (int32_t)n, top);
          if (root >= (int32_t)n) roots[h_seg[idx[a]]] = (uint32_t)root;
          top_seg_h.insert(top_seg_h.end(), top.size() - before, h_seg[idx[a]]);
          a = b;
        }
        const size_t M = top.size();  // < K <= top_cap
        std::vector<int32_t> tl(M), tr(M);
        std::vector<uint32_t> tc(M);
        std::vector<float> tb(M * 6);
        for (size_t k = 0; k < M; ++k) {
          tl[k] = top[k].left; tr[k] = top[k].right; tc[k] = top[k].cnt;
          for (int q = 0; q < 6; ++q) tb[6 * k + q] = top[k].box[q];
        }
        RR_TRY(dalloc(&top_seg, M ? M : 1));
        RR_TRY(cudaMemcpyAsync(out.left + n, tl.data(), M * 4, cudaMemcpyHostToDevice, st));
        RR_TRY(cudaMemcpyAsync(out.right + n, tr.data(), M * 4, cudaMemcpyHostToDevice, st));
        RR_TRY(cudaMemcpyAsync(range_count + n, tc.data(), M * 4, cudaMemcpyHostToDevice, st));
        RR_TRY(cudaMemcpyAsync(out.bounds + 6 * n, tb.data(), M * 24, cudaMemcpyHostToDevice, st));
        RR_TRY(cudaMemcpyAsync(top_seg, top_seg_h.data(), M * 4, cudaMemcpyHostToDevice, st));
        RR_TRY(cudaMemcpyAsync(out.seg_root, roots.data(), n_segs * 4, cudaMemcpyHostToDevice, st));
        RR_TRY(cudaStreamSynchronize(st));  // the host vectors go out of scope
        out.n_nodes = n + M;
      }
    }
    build_mark(st, "SAH top (host)");
    // 4-wide collapse, one launch per level of the wide hierarchy (the frontier size comes back to the host)
    RR_TRY(dalloc(&front_a, n_alloc));
    RR_TRY(dalloc(&front_b, n_alloc));
    RR_TRY(dalloc(&front_count, 2));
    RR_TRY(cudaMemsetAsync(out.nodes, 0, n_alloc * RR_NODE_QUADS * sizeof(float4), st));
    RR_TRY(cudaMemsetAsync(front_count, 0, 8, st));
    k_wide_roots<<<grid_for(n_segs, 128), 128, 0, st>>>(out.seg_root, out.seg_count, (int)n_segs, front_a, front_count);
    unsigned int h_count = 0;
    RR_TRY(cudaMemcpyAsync(&h_count, front_count, 4, cudaMemcpyDeviceToHost, st));
    RR_TRY(cudaStreamSynchronize(st));
    int32_t *fin = front_a, *fout = front_b;
    int level = 0;
    while (h_count) {
      unsigned int* cnt_out = front_count + ((level + 1) & 1);
      RR_TRY(cudaMemsetAsync(cnt_out, 0, 4, st));
      k_pack_wide<<<grid_for(h_count, 128), 128, 0, st>>>(fin, h_count, fout, cnt_out, out.order, d_prim_box, out.left, out.right,
                                                          out.bounds, out.seg_sfirst, (int)n_segs, out.seg_box, ref_offset,
                                                          range_first, range_count, leaf_max, n, top_seg, out.nodes);
      RR_TRY(cudaMemcpyAsync(&h_count, cnt_out, 4, cudaMemcpyDeviceToHost, st));
      RR_TRY(cudaStreamSynchronize(st));
      int32_t* t = fin; fin = fout; fout = t;
      level++;
    }
    out.wide_levels = (uint32_t)level;
    build_mark(st, "wide collapse (per level)");
  }
  RR_TRY(cudaMemcpyAsync(&out.max_depth, d_depth, 4, cudaMemcpyDeviceToHost, st));
  RR_TRY(cudaStreamSynchronize(st));
done:
  dev_free(keys_a); dev_free(keys_b); dev_free(vals_a); dev_free(vals_b); dev_free(seg_id); dev_free(hist);
  dev_free(seg_box_ord); dev_free(leaf_parent); dev_free(range_first); dev_free(range_count); dev_free(flags); dev_free(d_depth);
  dev_free(front_a); dev_free(front_b); dev_free(front_count);
  dev_free(c_ref); dev_free(c_seg); dev_free(c_cnt); dev_free(c_box); dev_free(c_counter); dev_free(top_seg);
  free(h_sfirst);
#undef RR_TRY
  return err;
}

}  // namespace rr
