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
  uint32_t tile_first[WIN_MAX_G + 1];              // prefix sum of tiles per region of the group
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
__global__ void __launch_bounds__(WIN_NTH) win_scatter_kernel(PartDev pd, WinDev wd, const uint32_t* __restrict__ order, uint32_t hb) {
  extern __shared__ __align__(16) uint32_t wsm[];
  const uint32_t wpr = 1u << wd.wpr_lg;
  uint32_t* cnt = wsm; uint32_t* lbase = cnt + wpr; uint32_t* lcur = lbase + wpr; uint32_t* gbase = lcur + wpr;
  uint32_t* stage = gbase + wpr;                   // WIN_TILE_UNITS * chunk_recs records
  __shared__ uint32_t warp_tot[WIN_NTH / 32];
  for(uint32_t i = threadIdx.x; i < wpr; i += WIN_NTH) cnt[i] = 0;
  __syncthreads();
  const uint32_t r = win_region_of_tile(wd, blockIdx.x);
  const uint32_t u0 = wd.unit_first[r] + (blockIdx.x - wd.tile_first[r]) * WIN_TILE_UNITS;
  const uint32_t u1 = min(u0 + WIN_TILE_UNITS, wd.unit_first[r + 1]);
  uint32_t rec[WIN_TILE_UNITS][4]; uint32_t nv[WIN_TILE_UNITS];
  {
    uint32_t chunk[WIN_TILE_UNITS], n[WIN_TILE_UNITS];
#pragma unroll
    for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) chunk[j] = u0 + j < u1 ? __ldg(order + u0 + j) : 0u;
#pragma unroll
    for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) n[j] = u0 + j < u1 ? __ldg(&pd.dir[chunk[j]].y) : 0u;
#pragma unroll
    for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) {
      const uint32_t i = threadIdx.x * 4;
      uint4 v = make_uint4(0, 0, 0, 0);
      nv[j] = i < n[j] ? min(4u, n[j] - i) : 0u;
      if(nv[j]) v = __ldcs(reinterpret_cast<const uint4*>(pd.pool + (size_t)chunk[j] * CHUNK_BYTES) + threadIdx.x);
      rec[j][0] = v.x; rec[j][1] = v.y; rec[j][2] = v.z; rec[j][3] = v.w;
    }
  }
#pragma unroll
  for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) {
#pragma unroll
    for(uint32_t q = 0; q < 4; ++q) if(q < nv[j]) atomicAdd(&cnt[((rec[j][q] >> hb) >> WIN_LG) & (wpr - 1)], 1u);
  }
  __syncthreads();
  // exclusive scan of cnt[0 .. wpr): each thread owns `per` consecutive windows
  const uint32_t per = (wpr + WIN_NTH - 1) / WIN_NTH;
  const uint32_t b = threadIdx.x * per;
  uint32_t s = 0;
  for(uint32_t i = b; i < min(b + per, wpr); ++i) s += cnt[i];
  uint32_t incl = s;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if((threadIdx.x & 31) >= (uint32_t)o) incl += v; }
  if((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for(uint32_t w = 0; w < (threadIdx.x >> 5); ++w) woff += warp_tot[w];
  uint32_t run = woff + incl - s;
  for(uint32_t i = b; i < min(b + per, wpr); ++i) {
    const uint32_t c = cnt[i];
    lbase[i] = run; lcur[i] = run;
    gbase[i] = c ? atomicAdd(&wd.wcursor[(r << wd.wpr_lg) + i], c) : 0u;
    run += c;
  }
  __syncthreads();
#pragma unroll
  for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j)
#pragma unroll
    for(uint32_t q = 0; q < 4; ++q)
      if(q < nv[j]) stage[atomicAdd(&lcur[((rec[j][q] >> hb) >> WIN_LG) & (wpr - 1)], 1u)] = rec[j][q];
  __syncthreads();
  uint32_t total = 0;
  for(uint32_t w = 0; w < WIN_NTH / 32; ++w) total += warp_tot[w];
  for(uint32_t i = threadIdx.x; i < total; i += WIN_NTH) {
    const uint32_t v = stage[i], w = ((v >> hb) >> WIN_LG) & (wpr - 1);
    const uint64_t dst = (uint64_t)gbase[w] + (i - lbase[w]);
    if(dst < wd.wrec_cap) wd.wrec[dst] = v;
  }
}

// ---- persistent window insert: probe in shared memory, windows and records moved by the TMA engine -------------
constexpr uint32_t WIN2_NTH = 1024;
constexpr uint32_t WIN2_RB  = 10240;               // records per batch (a window of iid input holds ~0.6 * WIN_SLOTS)
constexpr uint32_t WIN2_NONE = 0xFFFFFFFFu;
constexpr size_t   WIN2_SMEM = (size_t)2 * WIN_SLOTS * 4 + (size_t)2 * WIN2_RB * 4;

struct Win2Info { uint32_t task, b, n, pad; };

template<int KW>
__global__ void __launch_bounds__(WIN2_NTH, 1) win_insert2_kernel(TableDev T, PartDev pd, WinDev wd, const uint64_t* __restrict__ inv_lut_g, uint32_t nbytes) {
  extern __shared__ __align__(128) uint8_t w2smem[];
  __shared__ __align__(8) uint64_t full[2];
  __shared__ __align__(8) uint64_t extra;
  __shared__ Win2Info info[2];
  __shared__ uint32_t cursor[2];                   // next unclaimed record of the batch in each stage
  uint32_t* const winb0 = reinterpret_cast<uint32_t*>(w2smem);
  uint32_t* const recb0 = winb0 + 2 * WIN_SLOTS;
  const uint32_t fb = T.fbits, rb = T.rbits, hb = fb - rb;
  const uint32_t fmask = (1u << fb) - 1u, one = 1u << fb, cb = 32 - fb;
  const uint32_t hmask = hb ? ((1u << hb) - 1u) : 0u;
  uint32_t* tab = (uint32_t*)T.slots;
  const uint32_t n_tasks = wd.G << wd.wpr_lg;
  const uint32_t tid = threadIdx.x;
  uint32_t n_ins = 0, n_new = 0, n_rep = 0;

  auto slot_base_of = [&](uint32_t task) -> uint64_t {
    return ((uint64_t)(wd.g0 + (task >> wd.wpr_lg)) << pd.region_bits) + ((uint64_t)(task & ((1u << wd.wpr_lg) - 1)) << WIN_LG);
  };
  // thread 0: the next non-empty window of this CTA's stride at or after `from`, and the copies that bring it into stage s
  uint32_t next_from = blockIdx.x;                 // (thread 0 only)
  auto prefetch = [&](uint32_t s) {
    uint32_t t = next_from, c = 0;
    while(t < n_tasks && (c = wd.wcnt[t]) == 0) t += gridDim.x;
    if(t >= n_tasks) { info[s].task = WIN2_NONE; next_from = t; mbar_arrive(&full[s]); return; }
    next_from = t + gridDim.x;
    const uint32_t b = wd.wstart[t];
    info[s].task = t; info[s].b = b; info[s].n = c; cursor[s] = WIN2_NTH;
    const uint32_t rbytes = ((min(c, WIN2_RB) + 3u) & ~3u) * 4u;
    mbar_expect_tx(&full[s], WIN_SLOTS * 4u + rbytes);
    tma_load_1d(winb0 + s * WIN_SLOTS, tab + slot_base_of(t), WIN_SLOTS * 4u, &full[s]);
    tma_load_1d(recb0 + s * WIN2_RB, wd.wrec + b, rbytes, &full[s]);
  };
  if(tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&extra, 1); }
  __syncthreads();
  if(tid == 0) { fence_proxy_async(); prefetch(0); prefetch(1); }

  uint32_t xphase = 0;
  for(uint32_t it = 0; ; ++it) {
    const uint32_t s = it & 1u;
    mbar_wait(&full[s], (it >> 1) & 1u);
    const Win2Info inf = info[s];
    if(inf.task == WIN2_NONE) break;               // (the same for every thread of the CTA)
    uint32_t* const win = winb0 + s * WIN_SLOTS;
    const uint32_t* const recs = recb0 + s * WIN2_RB;
    const uint64_t slot_base = slot_base_of(inf.task);
    for(uint32_t off = 0; off < inf.n; off += WIN2_RB) {
      const uint32_t nb = min(WIN2_RB, inf.n - off);
      if(off) {                                    // a window with more than one batch of records (skewed input)
        __syncthreads();
        if(tid == 0) {
          const uint32_t rbytes = ((nb + 3u) & ~3u) * 4u;
          cursor[s] = WIN2_NTH;
          mbar_expect_tx(&extra, rbytes);
          tma_load_1d(recb0 + s * WIN2_RB, wd.wrec + inf.b + off, rbytes, &extra);
        }
        mbar_wait(&extra, xphase);
        xphase ^= 1u;
      }
      // Every lane keeps one record in flight and performs ONE probe per trip of the loop; a lane whose record is settled
      // takes the next unclaimed record of the batch (shared cursor), so the lanes of a warp stay busy until the batch is
      // exhausted whatever the lengths of their probe sequences.
      uint32_t i = tid;
      bool have = i < nb;
      uint32_t rec = have ? recs[i] : 0u;
      uint32_t local = (rec >> hb) & (WIN_SLOTS - 1), kf = ((rec & hmask) << rb) | 1u;
      uint32_t at = local, p = 0;
      while(have) {
        bool fetch = true;
        if(at < WIN_SLOTS) {
          const uint32_t o = atomicCAS(&win[at], 0u, kf | one);
          if(o == 0u) { ++n_new; ++n_ins; n_rep += p; }
          else if((o & fmask) == kf) {
            const uint32_t o2 = atomicAdd(&win[at], one);
            if((((o2 >> fb) + 1) >> cb) != 0) k2_carry(T.ovf_keys, T.ovf_vals, T.ovf_mask, T.stats, slot_base + at);
            ++n_ins; n_rep += p;
          } else if(p < T.max_reprobe) { ++p; at += p; ++kf; fetch = false; }          // pos + i(i+1)/2, reprobe field + 1
          else k2_fail<KW>(T.shard_index, T.local_lsize, T.lsize, T.stats, T.fail_keys, T.fail_counts, T.fail_cap, slot_base + local, rec & hmask, inv_lut_g, nbytes);
        } else {                                   // leaves the window: the global path takes it after this kernel
          const unsigned long long d = atomicAdd(wd.def_n, 1ull);
          if(d < wd.def_cap) { wd.def_pos[d] = slot_base + local; wd.def_high[d] = rec & hmask; }
          else atomicAdd(&T.stats[STAT_POOL_FULL], 1ull);
        }
        if(fetch) {
          i = atomicAdd(&cursor[s], 1u);
          have = i < nb;
          rec = have ? recs[i] : 0u;
          local = (rec >> hb) & (WIN_SLOTS - 1); kf = ((rec & hmask) << rb) | 1u;
          at = local; p = 0;
        }
      }
    }
    fence_proxy_async();                           // this thread's writes to the window, before the TMA engine reads it
    __syncthreads();
    if(tid == 0) {
      tma_store_1d(tab + slot_base, win, WIN_SLOTS * 4u);
      tma_commit_group();
      tma_wait_group_read0();                      // the stage may be overwritten
      prefetch(s);
    }
  }
  if(tid == 0) tma_wait_group0();                  // every window is back in the table before the kernel ends
  unsigned long long v[3] = { n_ins, n_new, n_rep };
#pragma unroll
  for(int q = 0; q < 3; ++q) {
#pragma unroll
    for(int o = 16; o; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  if((threadIdx.x & 31) == 0) {
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
