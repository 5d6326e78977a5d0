// minppo_b200 -- PRNG-derived minibatch permutation on device
// (replaces jax.random.split / jax.random.permutation at /root/reference/minppo/train.py:252,258).
//
//   permutation(key, B) = _shuffle: `rounds` x { key, sub = split(key); bits = random_bits(sub, 32, B);
//                                                 stable sort_key_val(bits, x) }          (jax/_src/random.py)
// Bit-exact with the oracle (oracle/threefry.py): Threefry-2x32/20 in both JAX bit-stream
// modes, and a STABLE least-significant-digit radix sort (4 passes of 8 bits).  Duplicate
// 32-bit sort keys do occur at B >= 2^18, so stability is observable and required.
//
// All `E` epochs of one update are generated in one batch (blockIdx.y = epoch): the key
// chain depends only on the input key, never on data.
#include "common.cuh"
#include "minppo_internal.h"
#include "threefry.cuh"

namespace minppo {

// Derive, for epoch e and sort round r, the subkey used for random bits, from the update's
// input key.  Chain: rng_{e+1}, k_e = split(rng_e);  inside permutation: k, s_r = split(k).
__host__ __device__ inline void round_subkey(const uint32_t key_in[2], int mode, int epoch, int round,
                                             uint32_t sub[2]) {
  uint32_t rng[2] = {key_in[0], key_in[1]}, a[2], b[2];
  for (int e = 0; e <= epoch; ++e) {
    key_split(rng, mode, a, b);
    rng[0] = a[0]; rng[1] = a[1];
  }
  uint32_t k[2] = {b[0], b[1]};                      // k_e
  for (int r = 0; r <= round; ++r) {
    key_split(k, mode, a, b);
    k[0] = a[0]; k[1] = a[1];
  }
  sub[0] = b[0]; sub[1] = b[1];
}

__global__ void key_advance_kernel(const uint32_t* __restrict__ key_in, uint32_t* __restrict__ key_out, int mode,
                                   int epochs) {
  griddep_wait();
  griddep_launch();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t rng[2] = {key_in[0], key_in[1]}, a[2], b[2];
    for (int e = 0; e < epochs; ++e) {
      key_split(rng, mode, a, b);
      rng[0] = a[0]; rng[1] = a[1];
    }
    key_out[0] = rng[0];
    key_out[1] = rng[1];
  }
}

// ---------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits; tile = 256 threads x ITEMS consecutive elements per warp
// ---------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                         // 32-element strips per warp
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;      // 2048

// bits for this round are produced on the fly in pass 0 (keys_in == nullptr)
struct SortArgs {
  const uint32_t* key_in_dev;   // update input key [2]
  int mode, round;
  int epoch_first, epoch_step;  // blockIdx.y = j sorts epoch epoch_first + j * epoch_step of the update's key chain
  uint32_t n;
  const uint32_t* keys_in;      // [E][n] (pass > 0)
  uint32_t* keys_out;           // [E][n]
  const int32_t* vals_in;       // [E][n] or nullptr (round 0, pass 0: iota)
  int32_t* vals_out;            // [E][n]
  uint32_t* hist;               // [E][256][tiles]
  uint32_t* totals;             // [E][256] digit totals of the current pass
  int shift;
  int tiles;
};

__device__ __forceinline__ uint32_t rs_load_key(const SortArgs& a, int epoch, uint32_t i, const uint32_t sub[2]) {
  if (a.keys_in) return a.keys_in[static_cast<size_t>(epoch) * a.n + i];
  return random_bits_at(sub, a.mode, i, a.n);
}

// Per-warp digit ranking of this warp's RS_ITEMS strips (in element order -> stable).
// wcount[warp][digit] ends as the warp's digit histogram; rank[s] = # earlier elements of the
// same digit within the warp.
__device__ __forceinline__ void rs_rank(const uint32_t (&key)[RS_ITEMS], const bool (&valid)[RS_ITEMS], int shift,
                                        uint32_t* wc /*[256] for this warp*/, uint32_t (&rank)[RS_ITEMS]) {
  const uint32_t lane = lane_id();
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int s = 0; s < RS_ITEMS; ++s) {
    const uint32_t d = (key[s] >> shift) & 0xFFu;
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid[s]);
    uint32_t peers = __match_any_sync(0xffffffffu, valid[s] ? d : 0x100u + lane) & vmask;
    uint32_t before = 0;
    if (valid[s]) {
      before = wc[d];
      rank[s] = before + __popc(peers & lt);
    }
    __syncwarp();
    if (valid[s] && (peers & lt) == 0) wc[d] = before + __popc(peers);    // group leader
    __syncwarp();
  }
}

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(SortArgs a) {
  __shared__ uint32_t wc[RS_WARPS][256];
  const int epoch = blockIdx.y, tile = blockIdx.x, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
  griddep_wait();                                   // launched with programmatic serialization (launch_chain)
  griddep_launch();
  __syncthreads();
  uint32_t sub[2] = {0, 0};
  if (!a.keys_in) {
    uint32_t kin[2] = {a.key_in_dev[0], a.key_in_dev[1]};
    round_subkey(kin, a.mode, a.epoch_first + epoch * a.epoch_step, a.round, sub);
  }
  // counts only (the ranks are the scatter kernel's business): per-warp shared-memory atomics
  const uint32_t base = static_cast<uint32_t>(tile) * RS_TILE + warp * (32 * RS_ITEMS) + lane_id();
#pragma unroll
  for (int s = 0; s < RS_ITEMS; ++s) {
    const uint32_t i = base + s * 32;
    if (i < a.n) atomicAdd(&wc[warp][(rs_load_key(a, epoch, i, sub) >> a.shift) & 0xFFu], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < 256; d += RS_THREADS) {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) c += wc[w][d];
    a.hist[(static_cast<size_t>(epoch) * 256 + d) * a.tiles + tile] = c;
  }
}

// Exclusive scan of every digit's row of per-tile counts, one block per (digit, epoch): hist[e][d][0..tiles) becomes
// the number of elements with digit d in earlier tiles; totals[e][d] = number of elements with digit d.  (The scatter
// kernel adds the exclusive scan over digits of `totals`.)  A first version scanned all 256 x tiles counters of an
// epoch in ONE block: 58 us per pass at B = 2^18 and linear in B, i.e. milliseconds for env-sharded global batches.
__global__ void __launch_bounds__(RS_THREADS) rs_scan_kernel(uint32_t* hist, uint32_t* totals, int tiles) {
  __shared__ uint32_t wsum[RS_WARPS];
  griddep_wait();
  griddep_launch();
  const int d = blockIdx.x, epoch = blockIdx.y;
  uint32_t* h = hist + (static_cast<size_t>(epoch) * 256 + d) * tiles;
  const int per = (tiles + RS_THREADS - 1) / RS_THREADS;
  const int b = threadIdx.x * per, e = min(tiles, b + per);
  uint32_t s = 0;
  for (int i = b; i < e; ++i) s += h[i];
  // block-wide exclusive scan of the per-thread sums
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= static_cast<uint32_t>(o)) inc += v;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) {
    if (static_cast<uint32_t>(w) < warp) wbase += wsum[w];
    total += wsum[w];
  }
  uint32_t run = wbase + inc - s;
  for (int i = b; i < e; ++i) {
    const uint32_t v = h[i];
    h[i] = run;
    run += v;
  }
  if (threadIdx.x == 0) totals[epoch * 256 + d] = total;
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(SortArgs a) {
  __shared__ uint32_t wc[RS_WARPS][256];
  const int epoch = blockIdx.y, tile = blockIdx.x, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
  griddep_wait();
  griddep_launch();
  __syncthreads();
  uint32_t sub[2] = {0, 0};
  if (!a.keys_in) {
    uint32_t kin[2] = {a.key_in_dev[0], a.key_in_dev[1]};
    round_subkey(kin, a.mode, a.epoch_first + epoch * a.epoch_step, a.round, sub);
  }
  uint32_t key[RS_ITEMS], rank[RS_ITEMS];
  int32_t val[RS_ITEMS];
  bool valid[RS_ITEMS];
  const uint32_t base = static_cast<uint32_t>(tile) * RS_TILE + warp * (32 * RS_ITEMS) + lane_id();
#pragma unroll
  for (int s = 0; s < RS_ITEMS; ++s) {
    const uint32_t i = base + s * 32;
    valid[s] = i < a.n;
    key[s] = valid[s] ? rs_load_key(a, epoch, i, sub) : 0u;
    val[s] = valid[s] ? (a.vals_in ? a.vals_in[static_cast<size_t>(epoch) * a.n + i] : static_cast<int32_t>(i)) : 0;
    rank[s] = 0;
  }
  rs_rank(key, valid, a.shift, wc[warp], rank);
  // exclusive scan over digits of the digit totals (256 threads == 256 digits)
  __shared__ uint32_t dscan[256];
  __shared__ uint32_t dws[RS_WARPS];
  {
    const uint32_t tot = a.totals[epoch * 256 + threadIdx.x];
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane_id() >= static_cast<uint32_t>(o)) inc += v;
    }
    if (lane_id() == 31) dws[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) if (w < warp) wbase += dws[w];
    dscan[threadIdx.x] = wbase + inc - tot;
  }
  __syncthreads();
  // per digit d (thread d): tbase[d] = elements of this tile with a smaller digit (exclusive scan over digits of the
  // tile's counts), gdst[d] = where the tile's run of digit d starts in the output; wc[w][d] becomes the LOCAL start of
  // warp w's elements of digit d.  Elements are first placed in shared memory in (digit, original order) order --
  // which is exactly their order in the output -- and then written out by consecutive threads: the elements of one
  // digit go to consecutive addresses instead of one 4-byte store per 32-byte sector.
  __shared__ uint32_t tbase[256];
  __shared__ uint32_t gdst[256];
  __shared__ uint32_t skey[RS_TILE];
  __shared__ int32_t sval[RS_TILE];
  {
    const int d = threadIdx.x;
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) cnt += wc[w][d];
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane_id() >= static_cast<uint32_t>(o)) inc += v;
    }
    __syncthreads();                                   // dws is reused (the digit-total scan above has been consumed)
    if (lane_id() == 31) dws[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) if (w < warp) wbase += dws[w];
    const uint32_t excl = wbase + inc - cnt;
    tbase[d] = excl;
    gdst[d] = dscan[d] + a.hist[(static_cast<size_t>(epoch) * 256 + d) * a.tiles + tile];
    uint32_t run = excl;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = wc[w][d];
      wc[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < RS_ITEMS; ++s) {
    if (valid[s]) {
      const uint32_t d = (key[s] >> a.shift) & 0xFFu;
      const uint32_t lp = wc[warp][d] + rank[s];
      skey[lp] = key[s];
      sval[lp] = val[s];
    }
  }
  __syncthreads();
  const uint32_t tile_n = min(static_cast<uint32_t>(RS_TILE), a.n - static_cast<uint32_t>(tile) * RS_TILE);
  for (uint32_t i = threadIdx.x; i < tile_n; i += RS_THREADS) {
    const uint32_t k = skey[i];
    const uint32_t d = (k >> a.shift) & 0xFFu;
    const size_t pos = static_cast<size_t>(epoch) * a.n + gdst[d] + (i - tbase[d]);
    a.keys_out[pos] = k;
    a.vals_out[pos] = sval[i];
  }
}

size_t perm_workspace_bytes(int epochs, long long B) {
  const size_t n = static_cast<size_t>(B), E = static_cast<size_t>(epochs);
  const size_t tiles = (n + RS_TILE - 1) / RS_TILE;
  // keys ping/pong + vals pong (+ final values go to caller's perm) + histograms
  return 2 * E * n * 4 + E * n * 4 + E * 256 * tiles * 4 + E * 256 * 4 + 1024;
}

// perm_out: int32 [E][B].  ws as sized above.  key_out (device, [2]) receives the key after E splits.
// Sorted epochs: epoch_first + j * epoch_step, j = 0 .. epochs-1 (perm_out[j]); key_out always advances `key_epochs` splits.
int perm_launch(const uint32_t* key_in_dev, uint32_t* key_out_dev, int mode, int epochs, long long B,
                int32_t* perm_out, void* ws, size_t ws_bytes, cudaStream_t stream, int epoch_first, int epoch_step,
                int key_epochs) {
  if (key_epochs < 0) key_epochs = epochs;
  if (epochs == 0) {
    if (key_out_dev) launch_chain(key_advance_kernel, dim3(1), dim3(32), 0, stream, key_in_dev, key_out_dev, mode, key_epochs);
    return cudaGetLastError() == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
  }
  if (B <= 0 || B > 0x7fffffffLL || epochs < 0) return MINPPO_ERR_ARG;
  if (ws_bytes < perm_workspace_bytes(epochs, B)) return MINPPO_ERR_WORKSPACE;
  const size_t n = static_cast<size_t>(B), E = static_cast<size_t>(epochs);
  const int tiles = static_cast<int>((n + RS_TILE - 1) / RS_TILE);
  uint32_t* k0 = reinterpret_cast<uint32_t*>(ws);
  uint32_t* k1 = k0 + E * n;
  int32_t* v1 = reinterpret_cast<int32_t*>(k1 + E * n);
  uint32_t* hist = reinterpret_cast<uint32_t*>(v1 + E * n);
  uint32_t* totals = hist + E * 256 * static_cast<size_t>(tiles);
  // rounds = ceil(3 ln B / ln(2^32 - 1))
  double lr = 3.0 * log(static_cast<double>(B > 1 ? B : 1)) / log(4294967295.0);
  int rounds = static_cast<int>(ceil(lr));
  dim3 grid(tiles, epochs);
  if (rounds == 0) cudaMemsetAsync(perm_out, 0, E * n * 4, stream);   // B == 1: identity
  // value buffers alternate between perm_out and v1 so that the last pass lands in perm_out
  const int total_passes = rounds * 4;
  for (int r = 0; r < rounds; ++r) {
    for (int pass = 0; pass < 4; ++pass) {
      const int gp = r * 4 + pass;
      SortArgs a;
      a.key_in_dev = key_in_dev;
      a.mode = mode;
      a.round = r;
      a.epoch_first = epoch_first; a.epoch_step = epoch_step;
      a.n = static_cast<uint32_t>(n);
      a.keys_in = pass == 0 ? nullptr : ((pass & 1) ? k0 : k1);
      a.keys_out = (pass & 1) ? k1 : k0;
      const bool out_is_perm = ((total_passes - 1 - gp) % 2) == 0;
      a.vals_out = out_is_perm ? perm_out : v1;
      a.vals_in = gp == 0 ? nullptr : (out_is_perm ? v1 : perm_out);
      a.hist = hist;
      a.totals = totals;
      a.shift = pass * 8;
      a.tiles = tiles;
      launch_chain(rs_hist_kernel, grid, dim3(RS_THREADS), 0, stream, a);
      launch_chain(rs_scan_kernel, dim3(256, epochs), dim3(RS_THREADS), 0, stream, hist, totals, tiles);
      launch_chain(rs_scatter_kernel, grid, dim3(RS_THREADS), 0, stream, a);
    }
  }
  if (key_out_dev) launch_chain(key_advance_kernel, dim3(1), dim3(32), 0, stream, key_in_dev, key_out_dev, mode, key_epochs);
  return cudaGetLastError() == cudaSuccess ? 0 : MINPPO_ERR_CUDA;
}

int perm_launch_count(long long B) {
  const int rounds = static_cast<int>(ceil(3.0 * log(static_cast<double>(B > 1 ? B : 1)) / log(4294967295.0)));
  return rounds * 4 * 3;
}

}  // namespace minppo
