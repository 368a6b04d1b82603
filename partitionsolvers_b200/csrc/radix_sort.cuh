// radix_sort.cuh -- stable LSD radix sort of (priority key, index) pairs, one or many segments.
//
// Replaces the reference's sort_by_priority (DP_impl.hpp:7-28: std::sort of indices by a[i]/b[i]) and LTSS's
// std::stable_sort (LTSS_impl.hpp:14-35).  Stable: elements with equal keys keep their input order, which is
// exactly std::stable_sort's result and one of the orders std::sort may produce.
//
// Keys are the IEEE quotients mapped to unsigned integers whose order is the numeric order (sign flip; -0 is
// first canonicalised to +0 because the reference's comparator `<` treats them as equal).  8-bit digits, least
// significant first; every pass is three kernels over tiles of TILE elements:
//   histogram  digit counts per (segment, tile)                       -> hist[seg][bin][tile]
//   scan       exclusive scan per segment over (bin-major, tile-minor) -> global offsets, in place
//   scatter    each tile re-read in order; an element's destination is offset[bin][tile] + its rank among the
//              tile's earlier elements with the same digit.  Ranks come from __match_any_sync inside a warp and
//              from per-bin counters that the warps of the CTA update in warp order, so the scatter is stable.
// Sorting is <0.2 % of a solve; the code favours being obviously stable over speed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ps {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

template <typename D>
struct SortKey;
template <>
struct SortKey<float> {
  using U = uint32_t;
  static constexpr int PASSES = 4;
  static __device__ __forceinline__ U encode(float v) {
    if (v == 0.0f) v = 0.0f;  // -0 -> +0
    const U u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
};
template <>
struct SortKey<double> {
  using U = uint64_t;
  static constexpr int PASSES = 8;
  static __device__ __forceinline__ U encode(double v) {
    if (v == 0.0) v = 0.0;
    const U u = (U)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
  }
};

// priority keys a[i]/b[i] (IEEE divide in D, as the reference's comparator computes them) + identity payload
template <typename D>
__global__ void sort_keys_kernel(long long total, int n, const D* __restrict__ a, const D* __restrict__ b,
                                 typename SortKey<D>::U* __restrict__ key, int* __restrict__ idx) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  key[g] = SortKey<D>::encode(a[g] / b[g]);
  idx[g] = (int)(g % n);
}

template <typename U>
__global__ void __launch_bounds__(SORT_THREADS)
sort_histogram_kernel(int n, int tiles, int shift, const U* __restrict__ key, unsigned int* __restrict__ hist) {
  __shared__ unsigned int cnt[256];
  const int seg = blockIdx.y, tile = blockIdx.x;
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const U* k = key + (long long)seg * n;
  const int base = tile * SORT_TILE;
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const int i = base + r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&cnt[(unsigned)(k[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[((long long)seg * 256 + threadIdx.x) * tiles + tile] = cnt[threadIdx.x];
}

// one CTA per segment: exclusive scan of 256*tiles counters in (bin, tile) order
__global__ void __launch_bounds__(SORT_THREADS)
sort_scan_kernel(int tiles, unsigned int* __restrict__ hist) {
  __shared__ unsigned int part[SORT_THREADS];
  unsigned int* h = hist + (long long)blockIdx.x * 256 * tiles;
  const int len = 256 * tiles;
  const int per = (len + SORT_THREADS - 1) / SORT_THREADS;
  const int lo = min(len, (int)threadIdx.x * per), hi = min(len, lo + per);
  unsigned int s = 0;
  for (int i = lo; i < hi; ++i) s += h[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int run = 0;
    for (int q = 0; q < SORT_THREADS; ++q) {
      const unsigned int v = part[q];
      part[q] = run;
      run += v;
    }
  }
  __syncthreads();
  unsigned int run = part[threadIdx.x];
  for (int i = lo; i < hi; ++i) {
    const unsigned int v = h[i];
    h[i] = run;
    run += v;
  }
}

template <typename U>
__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(int n, int tiles, int shift, const U* __restrict__ key_in, const int* __restrict__ idx_in,
                    U* __restrict__ key_out, int* __restrict__ idx_out, const unsigned int* __restrict__ offs) {
  __shared__ unsigned int cnt[256];  // running destination per bin for this tile
  const int seg = blockIdx.y, tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cnt[threadIdx.x] = offs[((long long)seg * 256 + threadIdx.x) * tiles + tile];
  __syncthreads();
  const long long sbase = (long long)seg * n;
  const int base = tile * SORT_TILE;
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const int i = base + r * SORT_THREADS + threadIdx.x;
    const bool valid = i < n;
    U kv = 0;
    int iv = 0;
    unsigned int d = 256;  // invalid lanes share a pseudo-digit and never write
    if (valid) {
      kv = key_in[sbase + i];
      iv = idx_in[sbase + i];
      d = (unsigned)(kv >> shift) & 255u;
    }
    const unsigned int peers = __match_any_sync(0xffffffffu, d);
    const unsigned int lower = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    unsigned int dst = 0;
    // warps take turns in warp order: earlier warps hold earlier elements of the round
    for (int w = 0; w < SORT_THREADS / 32; ++w) {
      if (warp == w) {
        unsigned int b0 = 0;
        if (valid && lane == leader) {
          b0 = cnt[d];
          cnt[d] = b0 + __popc(peers);
        }
        b0 = __shfl_sync(0xffffffffu, b0, leader);
        dst = b0 + lower;
      }
      __syncthreads();
    }
    if (valid) {
      key_out[sbase + dst] = kv;
      idx_out[sbase + dst] = iv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Small instances (n <= SORT_SMALL_MAX): the whole sort -- keys, the sort itself, the final gather of a and b -- in ONE
// kernel, one CTA per instance.  A bitonic network over (key, index) pairs in shared memory, padded to a power of two:
// comparing the pair lexicographically makes the result THE stable order (what the LSD radix passes above produce), and the
// network needs log2(N)(log2(N)+1)/2 barrier-separated steps (78 at N = 4096) where the eight 8-bit radix passes of a double key
// needed several hundred.  One SM's shared-memory bandwidth bounds it (every step moves the whole array); measured under
// ncu, n = 1000 float: 20 -> 14 us, n = 2089 double: 77 -> 56 us.
// Dynamic shared memory: N2 * (sizeof(U) + 4) bytes, N2 = n rounded up to a power of two.
// ---------------------------------------------------------------------------------------------
constexpr int SORT_SMALL_MAX = 4096;
constexpr int SORT_SMALL_THREADS = 1024;  // two compare-exchanges per thread and step at N = 4096

__host__ __device__ inline int sort_small_pow2(int n) {
  int m = 1;
  while (m < n) m <<= 1;
  return m;
}

template <typename D>
__global__ void __launch_bounds__(SORT_SMALL_THREADS)
sort_small_kernel(int n, const D* __restrict__ a, const D* __restrict__ b, int* __restrict__ perm,
                  D* __restrict__ a_sorted, D* __restrict__ b_sorted) {
  using U = typename SortKey<D>::U;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N2 = sort_small_pow2(n);
  U* key = reinterpret_cast<U*>(smem_raw);
  int* idx = reinterpret_cast<int*>(key + N2);
  const long long base = (long long)blockIdx.x * n;
  const int tid = threadIdx.x;
  for (int i = tid; i < N2; i += SORT_SMALL_THREADS) {
    // padding sorts behind every element: largest key, indices beyond n
    key[i] = (i < n) ? SortKey<D>::encode(a[base + i] / b[base + i]) : ~(U)0;
    idx[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= N2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (N2 >> 1); t += SORT_SMALL_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const U ki = key[i], kl = key[l];
        const int ii = idx[i], il = idx[l];
        const bool gt = (ki > kl) || (ki == kl && ii > il);  // pair i sorts after pair l
        if (gt == ((i & k) == 0)) {                           // ascending run: swap when out of order
          key[i] = kl;
          key[l] = ki;
          idx[i] = il;
          idx[l] = ii;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += SORT_SMALL_THREADS) {
    const int src_i = idx[i];
    perm[base + i] = src_i;
    a_sorted[base + i] = a[base + src_i];
    b_sorted[base + i] = b[base + src_i];
  }
}

}  // namespace ps
