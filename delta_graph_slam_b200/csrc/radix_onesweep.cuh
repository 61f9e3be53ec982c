// b200reg — stable LSD radix sort of (key, 32-bit value) pairs with ONE sweep over the data per digit:
// chained-scan ("decoupled look-back") digit passes in the manner of Onesweep (Adinets & Merrill 2022).
//
// voxel_sort.cuh sorts with three launches per digit (tile histograms -> one-block scan -> scatter): every key is
// read twice per digit and the one-block scan caps the tile count at 256 (2 M points).  The cooperative kernel of
// voxel_coop.cuh removes the launches but pays 13 grid barriers and is limited to one tile per SM.  Here:
//
//   k_os_histogram : ONE read of the keys builds the digit histograms of ALL passes (per-CTA shared-memory
//                    histograms, one atomicAdd per bin and CTA into hist[pass][256]);
//   k_os_pass      : per digit, one launch.  A CTA takes the next tile from a ticket counter (so every lower tile has
//                    started), counts its digits per warp, publishes the tile's 256 counts as "aggregate" status
//                    words, looks back over the preceding tiles' status words until it meets an inclusive prefix,
//                    publishes its own inclusive prefix, and scatters — keys and values are read once and written
//                    once per digit.  Inside a tile every warp owns a contiguous slice and ranks its keys in lane
//                    order (ballots over the digit bits), tiles are ordered by the look-back: the sort is stable, so its output is
//                    bit-identical to the three-launch path.
//
// Digit passes beyond the number of significant key bits (device-resident: it depends on the data) return at once,
// as in voxel_sort.cuh; the sorted data then sits in buffer A or B according to the parity of the passes that ran.
// No grid barrier, no cooperative launch, any number of tiles, two to four resident CTAs per SM.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kOsThreads = 256;
constexpr int kOsRadixBits = 8;
constexpr int kOsRadix = 256;
constexpr int kOsLookBack = 8;                     // predecessor tiles inspected per round trip of the look-back
constexpr uint32_t kOsFlagAggregate = 1u << 30, kOsFlagPrefix = 2u << 30, kOsFlagMask = 3u << 30, kOsCountMask = ~kOsFlagMask;

template <typename KeyT>
__device__ __forceinline__ uint32_t os_digit(KeyT k, int shift) { return (uint32_t)(k >> shift) & (kOsRadix - 1); }

// Lanes holding the same digit.  match.any costs a round per distinct value in the warp (the 32 different low digits of
// neighbouring voxel keys are its worst case), so both kernels issue the matches of several keys back to back in front
// of the code that consumes them and run enough warps per SM to cover the rest.  (Nine ballots over the digit bits
// instead — fixed latency, three times the instructions — measured 18 us slower per 1 M-point sort in the digit passes
// and the same in the histogram; profiles/r02_voxelgrid_1m.md.)
__device__ __forceinline__ uint32_t os_match(uint32_t dgt) { return __match_any_sync(0xffffffffu, dgt); }

// Optional tail of the LAST digit pass that runs: the records the sorted values index are gathered into sorted order
// while the tile's values are still in shared memory (the VoxelGrid centroid pass then reads consecutive records and
// no separate gather launch exists).  With a gather attached pass 0 always runs (a key range of one value has no
// significant bits: the pass is then a stable copy) unless *disabled.
struct OsGather {
  const float4* src;    // records indexed by the sorted values
  float4* dst;          // dst[j] = src[vals_sorted[j]]; nullptr = plain sort
  const int* disabled;  // device flag: nothing to sort (keys and values were never written)
};
constexpr OsGather kNoGather = {nullptr, nullptr, nullptr};

// all digit histograms in one read of the keys.  hist: [max_passes][256], zeroed before the launch.  Four keys per thread
// in flight.  The kernel is bound by the chain match -> counter update -> __syncwarp of every warp, not by its atomics
// (taking out the shared or the global ones changed nothing; 148 CTAs took 28 us on a 1 M-key sort, 296 took 20, 592
// took 19): up to four CTAs per SM.
template <typename KeyT>
static __global__ void __launch_bounds__(kOsThreads) k_os_histogram(const KeyT* __restrict__ keys, int n, const uint32_t* __restrict__ nbits_ptr, int max_passes, uint32_t* __restrict__ hist,
                                                             OsGather gather) {
  constexpr int MAXP = (int)sizeof(KeyT);
  constexpr int WARPS = kOsThreads / 32;
  // counters private to a warp: plain read-modify-write by the lane elected per distinct digit
  __shared__ uint32_t s[MAXP > 4 ? 4 : MAXP][WARPS][kOsRadix];
  constexpr int PASSES_PER_SWEEP = MAXP > 4 ? 4 : MAXP;  // 64-bit keys: two sweeps over the keys (32 KB of counters each)
  const uint32_t nbits = *nbits_ptr;
  if (gather.dst && *gather.disabled) return;
  int passes = min(max_passes, (int)((nbits + kOsRadixBits - 1) / kOsRadixBits));
  if (gather.dst && passes < 1) passes = 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int U = 4;
  for (int p0 = 0; p0 < passes; p0 += PASSES_PER_SWEEP) {
    const int np = min(PASSES_PER_SWEEP, passes - p0);
    for (int i = threadIdx.x; i < PASSES_PER_SWEEP * WARPS * kOsRadix; i += kOsThreads) (&s[0][0][0])[i] = 0;
    __syncthreads();
    // whole warps walk the keys together (the trip count is warp-uniform)
    for (int base = (blockIdx.x * kOsThreads + (threadIdx.x & ~31)) * U; base < n; base += gridDim.x * kOsThreads * U) {
      KeyT k[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * 32 + lane;
        ok[u] = i < n;
        k[u] = ok[u] ? keys[i] : (KeyT)0;
      }
      // the digits of one key go to the counters of different passes: their read-modify-writes are independent and
      // overlap; the next key's (another lane may be elected for the same counter) wait behind one __syncwarp
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int p = 0; p < PASSES_PER_SWEEP; ++p) {
          if (p < np) {
            const uint32_t dgt = ok[u] ? os_digit(k[u], (p0 + p) * kOsRadixBits) : (uint32_t)kOsRadix;
            const uint32_t peers = os_match(dgt);
            if (ok[u] && (__ffs(peers) - 1) == lane) s[p][warp][dgt] += __popc(peers);  // one lane per digit: no two lanes share a counter
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int p = 0; p < np; ++p) {
      uint32_t c = 0;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) c += s[p][w][threadIdx.x];
      if (c) atomicAdd(hist + (p0 + p) * kOsRadix + threadIdx.x, c);
    }
    __syncthreads();
  }
}

// One digit pass.  status: [n_tiles][256] words of this pass (zero = not yet published); ticket: this pass's tile
// counter (zero before the launch).  ITEMS keys per thread; the tile is staged in shared memory.
template <typename KeyT, int ITEMS>
static __global__ void __launch_bounds__(kOsThreads) k_os_pass(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n,
                                                        int pass, const uint32_t* __restrict__ nbits_ptr, const uint32_t* __restrict__ hist, uint32_t* status, unsigned int* ticket, OsGather gather) {
  const uint32_t nbits = *nbits_ptr;
  if (gather.dst) {
    if (*gather.disabled || (pass > 0 && (uint32_t)(pass * kOsRadixBits) >= nbits)) return;
  } else if ((uint32_t)(pass * kOsRadixBits) >= nbits) {
    return;
  }
  const bool do_gather = gather.dst != nullptr && (uint32_t)((pass + 1) * kOsRadixBits) >= nbits;  // the last pass that runs
  constexpr int WARPS = kOsThreads / 32;
  constexpr int TILE = kOsThreads * ITEMS;
  __shared__ uint32_t cnt[WARPS][kOsRadix];   // per-warp digit counts, then per-warp LOCAL bases (position inside the staged tile)
  __shared__ uint32_t s_tstart[kOsRadix];     // first position of digit d inside the staged tile
  __shared__ uint32_t s_gbase[kOsRadix];      // first output slot of this tile's digit-d run
  __shared__ KeyT s_keys[TILE];
  __shared__ uint32_t s_vals[TILE];
  __shared__ uint32_t s_warp[8], s_warp_t[8];
  __shared__ unsigned int s_tile;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < WARPS * kOsRadix; i += kOsThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int tile = (int)s_tile;
  const int shift = pass * kOsRadixBits;
  const int tbase = tile * TILE;
  const int wbase = tbase + warp * (32 * ITEMS);
  const int tile_n = min(TILE, n - tbase);
  const bool has_vals = vals_in != nullptr;  // nullptr: a keys-only sort
  KeyT k[ITEMS];
  uint32_t v[ITEMS];
  uint32_t wrank[(ITEMS + 1) / 2];  // rank of every key among its warp slice's keys of the same digit, two per register
  // ---- A: per-warp digit counts and, on the way, every key's rank inside (warp slice, digit): round after round the
  // slice's running count of the digit plus the lower lanes holding it in this round (all loads issued first)
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = wbase + r * 32 + lane;
    k[r] = i < n ? keys_in[i] : (KeyT)0;
  }
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = wbase + r * 32 + lane;
    v[r] = (has_vals && i < n) ? vals_in[i] : 0u;
  }
  // (the votes of all rounds first — they do not depend on each other — then the chain through the counters)
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = wbase + r * 32 + lane;
    const uint32_t dgt = i < n ? os_digit(k[r], shift) : (uint32_t)kOsRadix;  // kOsRadix = "no element"
    const uint32_t peers = os_match(dgt);
    // lower lanes with the digit (5 bits), lanes with the digit (6 bits), elected lane (1 bit)
    const uint32_t info = __popc(peers & ((1u << lane) - 1u)) | (__popc(peers) << 5) | (((__ffs(peers) - 1) == lane ? 1u : 0u) << 11);
    if (r & 1) wrank[r >> 1] |= info << 16;
    else wrank[r >> 1] = info;
  }
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = wbase + r * 32 + lane;
    const bool ok = i < n;
    const uint32_t dgt = os_digit(k[r], shift);
    const uint32_t info = (wrank[r >> 1] >> ((r & 1) * 16)) & 0xffffu;
    const uint32_t prior = ok ? cnt[warp][dgt] : 0u;
    __syncwarp();
    if (ok && (info >> 11)) cnt[warp][dgt] = prior + ((info >> 5) & 63u);
    __syncwarp();
    const uint32_t wr = prior + (info & 31u);
    wrank[r >> 1] = (wrank[r >> 1] & ~(0xffffu << ((r & 1) * 16))) | (wr << ((r & 1) * 16));
  }
  __syncthreads();
  // ---- B: thread d owns digit d.  Tile count -> publish -> look back -> publish the inclusive prefix
  {
    const int d = tid;
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) c += cnt[w][d];
    volatile uint32_t* st = status + (size_t)tile * kOsRadix + d;
    uint32_t excl = 0;
    if (tile == 0) {
      *st = kOsFlagPrefix | c;
    } else {
      *st = kOsFlagAggregate | c;
      // look back kOsLookBack tiles at a time: the loads are independent (one L2 round trip per batch, not per
      // tile) and are then consumed in order; a word that is not published yet ends the batch and is polled again.
      // (When every tile of a small sort is resident at once they all publish aggregates together and the prefix has
      // to travel the whole chain: walking it one dependent load at a time would cost a round trip per tile.)
      int t = tile - 1;
      bool done = false;
      while (!done) {
        uint32_t w[kOsLookBack];
#pragma unroll
        for (int j = 0; j < kOsLookBack; ++j) {
          const int tt = t - j;
          w[j] = tt >= 0 ? *(volatile const uint32_t*)(status + (size_t)tt * kOsRadix + d) : kOsFlagPrefix;  // in front of tile 0: an empty prefix
        }
        int used = 0;
#pragma unroll
        for (int j = 0; j < kOsLookBack; ++j) {
          if (!done && used == j && (w[j] & kOsFlagMask) != 0u) {
            excl += w[j] & kOsCountMask;
            used = j + 1;
            if (w[j] & kOsFlagPrefix) done = true;
          }
        }
        t -= used;
      }
      *st = kOsFlagPrefix | (excl + c);
    }
    // two exclusive scans over the digits: the global histogram (keys with a smaller digit anywhere) and this tile's
    // own counts (where the digit's run starts inside the staged tile)
    const uint32_t total_d = hist[pass * kOsRadix + d];
    uint32_t incl_g = total_d, incl_t = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t tg = __shfl_up_sync(0xffffffffu, incl_g, o);
      const uint32_t tt = __shfl_up_sync(0xffffffffu, incl_t, o);
      if (lane >= o) { incl_g += tg; incl_t += tt; }
    }
    if (lane == 31) { s_warp[warp] = incl_g; s_warp_t[warp] = incl_t; }
    __syncthreads();
    uint32_t off_g = 0, off_t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w < warp) { off_g += s_warp[w]; off_t += s_warp_t[w]; }
    const uint32_t tstart = off_t + incl_t - c;
    s_tstart[d] = tstart;
    s_gbase[d] = off_g + incl_g - total_d + excl;
    uint32_t run = tstart;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const uint32_t cw = cnt[w][d];
      cnt[w][d] = run;
      run += cw;
    }
  }
  __syncthreads();
  // ---- C: position of every key inside the tile = start of (digit, warp slice) + the rank from A (warp slices in
  // order, rounds in order, lanes in order: stable); the tile is staged in shared memory in that order ...
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int i = wbase + r * 32 + lane;
    if (i < n) {
      const uint32_t lpos = cnt[warp][os_digit(k[r], shift)] + ((wrank[r >> 1] >> ((r & 1) * 16)) & 0xffffu);
      s_keys[lpos] = k[r];
      s_vals[lpos] = v[r];
    }
  }
  __syncthreads();
  // ... and written out from there: consecutive threads hold consecutive positions of the staged tile, i.e. (mostly) the
  // same digit and therefore consecutive output slots — the scatter of a 256-way partition as runs of coalesced stores
  // instead of 32 different sectors per store instruction
#pragma unroll 4
  for (int r = 0; r < ITEMS; ++r) {
    const int p = r * kOsThreads + tid;
    if (p < tile_n) {
      const KeyT key = s_keys[p];
      const uint32_t d = os_digit(key, shift);
      const uint32_t dst = s_gbase[d] + ((uint32_t)p - s_tstart[d]);
      keys_out[dst] = key;
      if (has_vals) {
        const uint32_t val = s_vals[p];
        vals_out[dst] = val;
        if (do_gather) gather.dst[dst] = __ldg(gather.src + val);
      }
    }
  }
}

// Scratch of the sort: digit histograms, tile status words, tickets — one region per pass, zeroed with a single
// memset in front of the histogram kernel.
struct OneSweepScratch {
  DevBuf<uint32_t> buf;
  static int tiles(int n, int items) { return n > 0 ? (n + kOsThreads * items - 1) / (kOsThreads * items) : 1; }
  static size_t words(int n, int max_passes, int items) { return (size_t)max_passes * (kOsRadix + 32 + (size_t)tiles(n, items) * kOsRadix); }
};
template <typename KeyT>
constexpr int os_items() { return 8; }  // 2048-key tiles: three to four CTAs per SM, more warps to cover each other's latencies than 4096-key tiles

// enqueue: sorts (keys_a, vals_a)[0..n) — vals_a == nullptr: the keys alone — by the low *nbits_ptr key bits (device-resident count, <= 8 * max_passes);
// pass p reads A when p is even, B when odd, so the result lies in B after an odd number of passes that ran.
template <typename KeyT>
inline cudaError_t onesweep_sort(cudaStream_t st, OneSweepScratch& sc, KeyT* keys_a, uint32_t* vals_a, KeyT* keys_b, uint32_t* vals_b, int n, const uint32_t* nbits_ptr, int max_passes,
                                 OsGather gather = kNoGather) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e;
  constexpr int ITEMS = os_items<KeyT>();
  const int n_tiles = sc.tiles(n, ITEMS);
  const size_t words = sc.words(n, max_passes, ITEMS);
  if ((e = sc.buf.reserve(words)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(sc.buf.p, 0, words * sizeof(uint32_t), st)) != cudaSuccess) return e;
  uint32_t* hist = sc.buf.p;                                   // [max_passes][256]
  unsigned int* tickets = sc.buf.p + (size_t)max_passes * kOsRadix;  // [max_passes][32] (one 128-byte line each)
  uint32_t* status = sc.buf.p + (size_t)max_passes * (kOsRadix + 32);
  int hb = (n + kOsThreads * 8 - 1) / (kOsThreads * 8);
  if (hb > kNumSM * 4) hb = kNumSM * 4;
  launch_counter() += 1 + max_passes;
  k_os_histogram<KeyT><<<hb, kOsThreads, 0, st>>>(keys_a, n, nbits_ptr, max_passes, hist, gather);
  for (int p = 0; p < max_passes; ++p) {
    const KeyT* ki = (p & 1) ? keys_b : keys_a;
    const uint32_t* vi = vals_a ? ((p & 1) ? vals_b : vals_a) : nullptr;
    KeyT* ko = (p & 1) ? keys_a : keys_b;
    uint32_t* vo = vals_a ? ((p & 1) ? vals_a : vals_b) : nullptr;
    k_os_pass<KeyT, ITEMS><<<n_tiles, kOsThreads, 0, st>>>(ki, vi, ko, vo, n, p, nbits_ptr, hist, status + (size_t)p * n_tiles * kOsRadix, tickets + p * 32, gather);
  }
  return cudaGetLastError();
}

}  // namespace b200
