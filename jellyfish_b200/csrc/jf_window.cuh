// jf_window.cuh -- K2, the shared-memory window insert (the default form of K2 for 32-bit slots and 4-byte records).
//
// The L2 form of K2 (insert_chunks32_kernel) performs ~2 L2 operations per k-mer (first-probe CAS, reprobes,
// look-ahead loads).  Here the probing happens in shared memory:
//   win_hist / win_scan / win_scatter  split the 4-byte records of a group of regions by WINDOW
//        (2^WIN_LG slots = 64 KB of 32-bit slots) with a shared-memory staged tile sort, so that each
//        window's records are contiguous (12 B of traffic per record);
//   win_insert2  persistent, one CTA per SM, two stages: while the CTA applies the records of one window with
//        shared-memory CAS/add along the reference's probe sequence pos + i(i+1)/2, the TMA engine brings in the
//        next window and its records (cp.async.bulk + mbarrier) and writes the previous window back
//        (cp.async.bulk shared -> global): 8 B of table traffic per slot + 4 B per record, no global-memory
//        latency inside the probing loop.  A probe that would leave the window is DEFERRED (position + key
//        bits appended to a list); a counter carry goes to the side overflow table exactly as in the L2 kernels;
//   win_deferred  the deferred records, with the ordinary global probe sequence, after win_insert
//        of the same group has completed (stream order), so no slot is ever touched by a window CTA
//        and by the global path at the same time.
// Slots only ever fill up, so a key that left its window because every slot of its sequence inside
// the window belongs to other keys finds the same situation on every later visit: all its
// occurrences are deferred, and it cannot end up in two slots.
#pragma once

namespace jfk {

constexpr uint32_t WIN_LG = 14;                    // slots per window
constexpr uint32_t WIN_SLOTS = 1u << WIN_LG;
constexpr uint32_t WIN_TILE_UNITS = 4;             // chunks per partition tile: 8192 records
constexpr uint32_t WIN_MAX_G = 64;                 // regions per group
constexpr uint32_t WIN_MAX_WPR = 2048;             // windows per region (region_bits - WIN_LG <= 11)
constexpr uint32_t WIN_NTH = 512;

struct WinDev {
  uint32_t g0, G, wpr_lg, n_tiles;
  uint32_t tile_first[WIN_MAX_G + 1];              // prefix sum of tiles per region of the group (win_hist: WIN_TILE_UNITS chunks)
  uint32_t stile_first[WIN_MAX_G + 1];             // the same for win_scatter's larger tiles (WIN_ST_UNITS chunks)
  uint32_t unit_first[WIN_MAX_G + 1];              // first unit (index into `order`) of each region; [G] = end
  uint32_t* wstart;                                // [(G << wpr_lg) + 1] exclusive offsets into wrec after win_scan
  uint32_t* wcursor;                               // [G << wpr_lg] counts (win_hist), then write cursors (win_scatter)
  uint32_t* wcnt;                                  // [G << wpr_lg] records per window (win_scan); runs start on 16-byte boundaries
  uint32_t* wrec; uint64_t wrec_cap;               // records grouped by (region, window)
  uint64_t* def_pos; uint32_t* def_high; unsigned long long* def_n; uint64_t def_cap;
};

__device__ __forceinline__ uint32_t win_region_of_tile(const WinDev& wd, uint32_t tile) {
  uint32_t r = 0;
  while(r + 1 < wd.G && wd.tile_first[r + 1] <= tile) ++r;
  return r;
}

// ---- counts per (region, window) ---------------------------------------------------------------
__global__ void __launch_bounds__(WIN_NTH) win_hist_kernel(PartDev pd, WinDev wd, const uint32_t* __restrict__ order, uint32_t hb) {
  __shared__ uint32_t cnt[WIN_MAX_WPR];
  const uint32_t wpr = 1u << wd.wpr_lg;
  for(uint32_t i = threadIdx.x; i < wpr; i += WIN_NTH) cnt[i] = 0;
  __syncthreads();
  const uint32_t r = win_region_of_tile(wd, blockIdx.x);
  const uint32_t u0 = wd.unit_first[r] + (blockIdx.x - wd.tile_first[r]) * WIN_TILE_UNITS;
  const uint32_t u1 = min(u0 + WIN_TILE_UNITS, wd.unit_first[r + 1]);
  // the loads of the tile's chunks level by level (order -> directory -> records), so that they overlap
  uint32_t chunk[WIN_TILE_UNITS], n[WIN_TILE_UNITS]; uint4 v[WIN_TILE_UNITS];
#pragma unroll
  for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) chunk[j] = u0 + j < u1 ? __ldg(order + u0 + j) : 0u;
#pragma unroll
  for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) n[j] = u0 + j < u1 ? __ldg(&pd.dir[chunk[j]].y) : 0u;
#pragma unroll
  for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) {
    v[j] = make_uint4(0, 0, 0, 0);
    if(threadIdx.x * 4 < n[j]) v[j] = __ldg(reinterpret_cast<const uint4*>(pd.pool + (size_t)chunk[j] * CHUNK_BYTES) + threadIdx.x);
  }
#pragma unroll
  for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) {
    const uint32_t i = threadIdx.x * 4;
    const uint32_t rec[4] = { v[j].x, v[j].y, v[j].z, v[j].w };
#pragma unroll
    for(uint32_t q = 0; q < 4; ++q) if(i + q < n[j]) atomicAdd(&cnt[((rec[q] >> hb) >> WIN_LG) & (wpr - 1)], 1u);
  }
  __syncthreads();
  for(uint32_t i = threadIdx.x; i < wpr; i += WIN_NTH) if(cnt[i]) atomicAdd(&wd.wcursor[(r << wd.wpr_lg) + i], cnt[i]);
}

// ---- exclusive scan of the counts (one CTA) ------------------------------------------------------
__global__ void __launch_bounds__(1024) win_scan_kernel(WinDev wd, unsigned long long* __restrict__ stats) {
  __shared__ uint32_t part[1024];
  const uint32_t n = wd.G << wd.wpr_lg;
  const uint32_t per = (n + 1023) / 1024;
  const uint32_t b = threadIdx.x * per, e = min(b + per, n);
  uint32_t s = 0;
  for(uint32_t i = b; i < e; ++i) s += (wd.wcursor[i] + 3u) & ~3u;
  part[threadIdx.x] = s;
  __syncthreads();
  for(uint32_t d = 1; d < 1024; d <<= 1) {
    const uint32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  // every window's run starts on a 16-byte boundary (the TMA engine copies it): counts are rounded up to 4 records
  uint32_t run = part[threadIdx.x] - s;            // exclusive prefix of this thread's segment
  for(uint32_t i = b; i < e; ++i) {
    const uint32_t c = wd.wcursor[i];
    wd.wstart[i] = run; wd.wcursor[i] = run; wd.wcnt[i] = c;
    run += (c + 3u) & ~3u;
  }
  if(threadIdx.x == 1023) {
    wd.wstart[n] = part[1023];
    if(part[1023] > wd.wrec_cap) atomicAdd(&stats[STAT_POOL_FULL], 1ull);   // the host sizes groups so that this cannot happen
  }
}

// ---- tile sort by window, runs written to the group buffer ----------------------------------------
// A tile = WIN_ST_UNITS chunks of one region (24 K records): histogram by window in shared memory, scan, stable placement in
// a shared-memory staging buffer, then the runs (one per window, ~24 records) go to their places in the group buffer with
// coalesced stores.  The GPU retires a bounded number of store transactions per second whatever their size
// (scripts/micro/scatter_store.cu), so the tile is as large as two resident CTAs per SM allow.  The chunks are read twice
// (second time from L2) instead of being held in registers across the scan.
constexpr uint32_t WIN_ST_UNITS = 12;
constexpr uint32_t WIN_ST_NTH = 1024;

__global__ void __launch_bounds__(WIN_ST_NTH, 2) win_scatter_kernel(PartDev pd, WinDev wd, const uint32_t* __restrict__ order, uint32_t hb) {
  extern __shared__ __align__(16) uint32_t wsm[];
  const uint32_t wpr = 1u << wd.wpr_lg;
  uint32_t* cnt = wsm; uint32_t* lbase = cnt + wpr; uint32_t* lcur = lbase + wpr; uint32_t* gbase = lcur + wpr;
  uint32_t* stage = gbase + wpr;                   // WIN_ST_UNITS * chunk_recs records
  __shared__ uint32_t warp_tot[WIN_ST_NTH / 32];
  __shared__ uint32_t s_chunk[WIN_ST_UNITS], s_n[WIN_ST_UNITS];
  const uint32_t tid = threadIdx.x;
  for(uint32_t i = tid; i < wpr; i += WIN_ST_NTH) cnt[i] = 0;
  // which region this tile belongs to (tiles are numbered region by region)
  uint32_t r = 0;
  while(r + 1 < wd.G && wd.stile_first[r + 1] <= blockIdx.x) ++r;
  const uint32_t u0 = wd.unit_first[r] + (blockIdx.x - wd.stile_first[r]) * WIN_ST_UNITS;
  const uint32_t u1 = min(u0 + WIN_ST_UNITS, wd.unit_first[r + 1]);
  if(tid < WIN_ST_UNITS) {
    uint32_t c = 0, n = 0;
    if(u0 + tid < u1) { c = __ldg(order + u0 + tid); n = __ldg(&pd.dir[c].y); }
    s_chunk[tid] = c; s_n[tid] = n;
  }
  __syncthreads();
  const uint32_t half = tid >> 9, piece = tid & 511u;      // two chunks per trip, 512 x 16 bytes each
  const uint32_t wmask = wpr - 1;
#pragma unroll
  for(uint32_t it = 0; it < WIN_ST_UNITS / 2; ++it) {
    const uint32_t j = 2 * it + half, n = s_n[j];
    if(piece * 4 < n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(pd.pool + (size_t)s_chunk[j] * CHUNK_BYTES) + piece);
      const uint32_t rec[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for(uint32_t q = 0; q < 4; ++q) if(piece * 4 + q < n) atomicAdd(&cnt[((rec[q] >> hb) >> WIN_LG) & wmask], 1u);
    }
  }
  __syncthreads();
  // exclusive scan of cnt[0 .. wpr): each thread owns `per` consecutive windows
  const uint32_t per = (wpr + WIN_ST_NTH - 1) / WIN_ST_NTH;
  const uint32_t b = tid * per;
  uint32_t s = 0;
  for(uint32_t i = b; i < min(b + per, wpr); ++i) s += cnt[i];
  uint32_t incl = s;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if((tid & 31) >= (uint32_t)o) incl += v; }
  if((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  uint32_t woff = 0, total = 0;
  for(uint32_t w = 0; w < WIN_ST_NTH / 32; ++w) { const uint32_t x = warp_tot[w]; if(w < (tid >> 5)) woff += x; total += x; }
  uint32_t run = woff + incl - s;
  for(uint32_t i = b; i < min(b + per, wpr); ++i) {
    const uint32_t c = cnt[i];
    lbase[i] = run; lcur[i] = run;
    gbase[i] = (c ? atomicAdd(&wd.wcursor[(r << wd.wpr_lg) + i], c) : 0u) - run;      // (output position = this + index in the staging buffer)
    run += c;
  }
  __syncthreads();
#pragma unroll
  for(uint32_t it = 0; it < WIN_ST_UNITS / 2; ++it) {
    const uint32_t j = 2 * it + half, n = s_n[j];
    if(piece * 4 < n) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(pd.pool + (size_t)s_chunk[j] * CHUNK_BYTES) + piece);   // (L2 hit: read a moment ago)
      const uint32_t rec[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for(uint32_t q = 0; q < 4; ++q) if(piece * 4 + q < n) stage[atomicAdd(&lcur[((rec[q] >> hb) >> WIN_LG) & wmask], 1u)] = rec[q];
    }
  }
  __syncthreads();
  for(uint32_t i = tid; i < total; i += WIN_ST_NTH) {
    const uint32_t v = stage[i], w = ((v >> hb) >> WIN_LG) & wmask;
    const uint32_t dst = gbase[w] + i;
    if(dst < wd.wrec_cap) wd.wrec[dst] = v;
  }
}

// ---- persistent window insert: probe in shared memory, windows and records moved by the TMA engine -------------
// Warp 0 is the PRODUCER: it finds this CTA's next non-empty window, brings the window and its records into one of two
// stages (cp.async.bulk + mbarrier `full`), and, once the consumers have released a stage (mbarrier `empty`), writes the
// window back (cp.async.bulk shared -> global) and refills the stage.  Warps 1..31 are CONSUMERS: they never meet at a
// block-wide barrier -- a warp that runs out of records in one stage moves on to the other one.
constexpr uint32_t WIN2_NTH = 1024;
constexpr uint32_t WIN2_CONS = WIN2_NTH - 32;      // consumer threads
constexpr uint32_t WIN2_RB  = 10240;               // records per batch (a window of iid input holds ~0.6 * WIN_SLOTS)
constexpr uint32_t WIN2_BLK = 256;                 // records a consumer warp claims at a time
constexpr uint32_t WIN2_NONE = 0xFFFFFFFFu;
constexpr size_t   WIN2_SMEM = (size_t)2 * WIN_SLOTS * 4 + (size_t)2 * WIN2_RB * 4;

struct Win2Info { uint32_t task, b, n, pad; };

template<int KW>
__global__ void __launch_bounds__(WIN2_NTH, 1) win_insert2_kernel(TableDev T, PartDev pd, WinDev wd, const uint64_t* __restrict__ inv_lut_g, uint32_t nbytes) {
  extern __shared__ __align__(128) uint8_t w2smem[];
  __shared__ __align__(8) uint64_t full[2];
  __shared__ __align__(8) uint64_t empty[2];
  __shared__ Win2Info info[2];
  __shared__ uint32_t cursor[2];                   // next unclaimed record of the batch in each stage
  uint32_t* const winb0 = reinterpret_cast<uint32_t*>(w2smem);
  uint32_t* const recb0 = winb0 + 2 * WIN_SLOTS;
  uint32_t* tab = (uint32_t*)T.slots;
  const uint32_t n_tasks = wd.G << wd.wpr_lg;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;

  auto slot_base_of = [&](uint32_t task) -> uint64_t {
    return ((uint64_t)(wd.g0 + (task >> wd.wpr_lg)) << pd.region_bits) + ((uint64_t)(task & ((1u << wd.wpr_lg) - 1)) << WIN_LG);
  };
  if(tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    mbar_init(&empty[0], WIN2_CONS / 32); mbar_init(&empty[1], WIN2_CONS / 32);     // one arrival per consumer warp
    fence_proxy_async();
  }
  __syncthreads();

  if(tid < 32) {
    // ================================= producer =================================
    if(lane == 0) {
      uint32_t next_from = blockIdx.x;
      uint32_t ephase[2] = { 0, 0 };
      uint32_t cur_task[2] = { WIN2_NONE, WIN2_NONE }, cur_n[2] = { 0, 0 }, cur_b[2] = { 0, 0 };
      auto load_batch = [&](uint32_t s, uint32_t off, bool with_window) {
        const uint32_t nb = min(WIN2_RB, cur_n[s] - off);
        const uint32_t rbytes = ((nb + 3u) & ~3u) * 4u;
        cursor[s] = 0;
        mbar_expect_tx(&full[s], rbytes + (with_window ? WIN_SLOTS * 4u : 0u));
        if(with_window) tma_load_1d(winb0 + s * WIN_SLOTS, tab + slot_base_of(cur_task[s]), WIN_SLOTS * 4u, &full[s]);
        tma_load_1d(recb0 + s * WIN2_RB, wd.wrec + cur_b[s] + off, rbytes, &full[s]);
      };
      auto next_window = [&](uint32_t s) {           // the next non-empty window of this CTA's stride into stage s
        uint32_t t = next_from, c = 0;
        while(t < n_tasks && (c = wd.wcnt[t]) == 0) t += gridDim.x;
        if(t >= n_tasks) { next_from = t; cur_task[s] = WIN2_NONE; info[s].task = WIN2_NONE; mbar_arrive(&full[s]); return; }
        next_from = t + gridDim.x;
        cur_task[s] = t; cur_n[s] = c; cur_b[s] = wd.wstart[t];
        info[s].task = t; info[s].b = cur_b[s]; info[s].n = c;
        load_batch(s, 0, true);
      };
      next_window(0); next_window(1);
      for(uint32_t s = 0; cur_task[s] != WIN2_NONE; s ^= 1u) {
        // the consumers work through the batches of stage s in order
        for(uint32_t off = WIN2_RB; off < cur_n[s]; off += WIN2_RB) {
          mbar_wait(&empty[s], ephase[s]); ephase[s] ^= 1u;
          load_batch(s, off, false);
        }
        mbar_wait(&empty[s], ephase[s]); ephase[s] ^= 1u;   // every consumer warp is done with this window
        tma_store_1d(tab + slot_base_of(cur_task[s]), winb0 + s * WIN_SLOTS, WIN_SLOTS * 4u);
        tma_commit_group();
        tma_wait_group_read0();                    // the stage may be overwritten
        next_window(s);
      }
      tma_wait_group0();                           // every window is back in the table before the kernel ends
    }
    return;
  }

  // ================================= consumers =================================
  const uint32_t fb = T.fbits, rb = T.rbits, hb = fb - rb;
  const uint32_t fmask = (1u << fb) - 1u, one = 1u << fb, cb = 32 - fb;
  const uint32_t hmask = hb ? ((1u << hb) - 1u) : 0u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t n_ins = 0, n_new = 0, n_rep = 0;
  uint32_t fphase[2] = { 0, 0 };
  for(uint32_t s = 0; ; s ^= 1u) {
    mbar_wait(&full[s], fphase[s]); fphase[s] ^= 1u;
    const Win2Info inf = info[s];
    if(inf.task == WIN2_NONE) break;               // (the same for every consumer)
    uint32_t* const win = winb0 + s * WIN_SLOTS;
    const uint32_t* const recs = recb0 + s * WIN2_RB;
    const uint64_t slot_base = slot_base_of(inf.task);
    for(uint32_t off = 0; off < inf.n; off += WIN2_RB) {
      const uint32_t nb = min(WIN2_RB, inf.n - off);
      if(off) { mbar_wait(&full[s], fphase[s]); fphase[s] ^= 1u; }
      // Every lane keeps one record in flight and performs ONE probe per trip of the loop; the lanes whose record is settled
      // take the next records of the warp's current block (WIN2_BLK records, claimed with one shared-memory atomic per block;
      // inside a block the indices come from a warp-uniform register counter and a ballot), so a warp stays full until the
      // batch is exhausted whatever the lengths of the probe sequences.
      bool have = false, want = true;
      uint32_t rec = 0, local = 0, kf = 0, at = 0, p = 0;
      uint32_t blk_next = 0, blk_end = 0;          // (warp-uniform)
      bool more = true;                            // the batch may still have unclaimed blocks (warp-uniform)
      for(;;) {
        const uint32_t need = __ballot_sync(0xffffffffu, want);
        if(need && more) {
          if(blk_next >= blk_end) {                // claim the next block
            uint32_t base = 0;
            if(lane == 0) base = atomicAdd(&cursor[s], WIN2_BLK);
            blk_next = __shfl_sync(0xffffffffu, base, 0);
            blk_end = min(blk_next + WIN2_BLK, nb);
            more = blk_next < nb;
          }
          if(want && more) {
            const uint32_t i = blk_next + __popc(need & lt_mask);
            if(i < blk_end) {
              rec = recs[i];
              local = (rec >> hb) & (WIN_SLOTS - 1); kf = ((rec & hmask) << rb) | 1u;
              at = local; p = 0;
              have = true; want = false;
            }                                      // (else: asks again on the next trip, from the next block)
          }
          blk_next = min(blk_next + (uint32_t)__popc(need), blk_end);
        }
        if(!__any_sync(0xffffffffu, have)) { if(!more) break; else continue; }
        if(have) {
          if(at < WIN_SLOTS) {
            const uint32_t o = atomicCAS(&win[at], 0u, kf | one);
            if(o == 0u) { ++n_new; ++n_ins; n_rep += p; want = true; }
            else if((o & fmask) == kf) {
              const uint32_t o2 = atomicAdd(&win[at], one);
              if((((o2 >> fb) + 1) >> cb) != 0) k2_carry(T.ovf_keys, T.ovf_vals, T.ovf_mask, T.stats, slot_base + at);
              ++n_ins; n_rep += p; want = true;
            } else if(p < T.max_reprobe) { ++p; at += p; ++kf; }            // pos + i(i+1)/2, reprobe field + 1
            else {
              k2_fail<KW>(T.shard_index, T.local_lsize, T.lsize, T.stats, T.fail_keys, T.fail_counts, T.fail_cap, slot_base + local, rec & hmask, inv_lut_g, nbytes);
              want = true;
            }
          } else {                                 // leaves the window: the global path takes it after this kernel
            const unsigned long long d = atomicAdd(wd.def_n, 1ull);
            if(d < wd.def_cap) { wd.def_pos[d] = slot_base + local; wd.def_high[d] = rec & hmask; }
            else atomicAdd(&T.stats[STAT_POOL_FULL], 1ull);
            want = true;
          }
          have = !want;
        }
      }
      fence_proxy_async();                         // this thread's writes to the window, before the TMA engine reads it
      __syncwarp();
      if(lane == 0) mbar_arrive(&empty[s]);        // this warp is done with the batch
    }
  }
  unsigned long long v[3] = { n_ins, n_new, n_rep };
#pragma unroll
  for(int q = 0; q < 3; ++q) {
#pragma unroll
    for(int o = 16; o; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  if(lane == 0) {
    if(v[0]) atomicAdd(&T.stats[STAT_INSERTED], v[0]);
    if(v[1]) atomicAdd(&T.stats[STAT_DISTINCT], v[1]);
    if(v[2]) atomicAdd(&T.stats[STAT_REPROBES], v[2]);
  }
}

// ---- the deferred records: ordinary probe sequence in global memory ---------------------------------
template<int KW>
__global__ void __launch_bounds__(256) win_deferred_kernel(TableDev T, WinDev wd, const uint64_t* __restrict__ inv_lut_g, uint32_t nbytes) {
  const uint32_t fb = T.fbits, rb = T.rbits;
  const uint32_t fmask = (1u << fb) - 1u, one = 1u << fb, cb = 32 - fb;
  uint32_t* tab = (uint32_t*)T.slots;
  const unsigned long long n = min(*wd.def_n, (unsigned long long)wd.def_cap);
  uint32_t n_ins = 0, n_new = 0, n_rep = 0;
  for(unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint64_t base = wd.def_pos[i];
    const uint32_t high = wd.def_high[i], kf0 = high << rb;
    bool ok = true;
    const uint32_t o = atomicCAS(&tab[base], 0u, (kf0 | 1u) | one);
    if(o == 0u) ++n_new;
    else if((o & fmask) == (kf0 | 1u)) {
      const uint32_t o2 = atomicAdd(&tab[base], one);
      if((((o2 >> fb) + 1) >> cb) != 0) k2_carry(T.ovf_keys, T.ovf_vals, T.ovf_mask, T.stats, base);
    } else {
      const uint32_t w = k2_walk(tab, base, kf0, fb, T.max_reprobe);
      ok = w != 0;
      if(ok) {
        const uint32_t p = (w & 0xFFFFu) - 1;
        n_rep += p; if(w & 0x10000u) ++n_new;
        if(w & 0x20000u) k2_carry(T.ovf_keys, T.ovf_vals, T.ovf_mask, T.stats, base + tri(p));
      }
    }
    if(ok) ++n_ins;
    else k2_fail<KW>(T.shard_index, T.local_lsize, T.lsize, T.stats, T.fail_keys, T.fail_counts, T.fail_cap, base, high, inv_lut_g, nbytes);
  }
  unsigned long long v[3] = { n_ins, n_new, n_rep };
#pragma unroll
  for(int q = 0; q < 3; ++q) {
#pragma unroll
    for(int o2 = 16; o2; o2 >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o2);
  }
  if((threadIdx.x & 31) == 0) {
    if(v[0]) atomicAdd(&T.stats[STAT_INSERTED], v[0]);
    if(v[1]) atomicAdd(&T.stats[STAT_DISTINCT], v[1]);
    if(v[2]) atomicAdd(&T.stats[STAT_REPROBES], v[2]);
  }
}

}  // namespace jfk
