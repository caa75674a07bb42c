// jf_window.cuh -- EXPERIMENTAL second form of K2 ("window insert"), off unless JFGPU_K2_WINDOW is set.
//
// Status: written at the end of round 1 after the GPU budget was spent -- compiled for sm_100a, never
// run.  Nothing here is reachable in the default configuration; it is the starting point for round 2
// ("what comes next").
//
// K2 today performs ~2 L2 operations per k-mer (first-probe CAS, reprobes, look-ahead loads) and sits
// at the L2 ceiling for that mix.  Here the probing moves into shared memory:
//   win_hist / win_scan / win_scatter  split the 4-byte records of a group of regions by WINDOW
//        (2^WIN_LG slots = 64 KB of 32-bit slots) with a shared-memory staged tile sort, so that each
//        window's records are contiguous (12 B of traffic per record);
//   win_insert   one CTA per window: load the window's slots into shared memory, apply its records
//        with shared-memory CAS/add along the reference's probe sequence pos + i(i+1)/2, store the
//        window back (8 B of table traffic per slot + 4 B per record).  A probe that would leave the
//        window is DEFERRED (position + key bits appended to a list); a counter carry goes to the
//        side overflow table exactly as in the L2 kernels;
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
  for(uint32_t u = u0; u < u1; ++u) {
    const uint32_t chunk = order[u];
    const uint32_t n = pd.dir[chunk].y;
    const uint32_t i = threadIdx.x * 4;
    if(i < n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(pd.pool + (size_t)chunk * CHUNK_BYTES) + threadIdx.x);
      const uint32_t rec[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for(uint32_t q = 0; q < 4; ++q) if(i + q < n) atomicAdd(&cnt[((rec[q] >> hb) >> WIN_LG) & (wpr - 1)], 1u);
    }
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
  for(uint32_t i = b; i < e; ++i) s += wd.wcursor[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for(uint32_t d = 1; d < 1024; d <<= 1) {
    const uint32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;            // exclusive prefix of this thread's segment
  for(uint32_t i = b; i < e; ++i) { const uint32_t c = wd.wcursor[i]; wd.wstart[i] = run; wd.wcursor[i] = run; run += c; }
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
#pragma unroll
  for(uint32_t j = 0; j < WIN_TILE_UNITS; ++j) {
    nv[j] = 0;
    if(u0 + j < u1) {
      const uint32_t chunk = order[u0 + j];
      const uint32_t n = pd.dir[chunk].y;
      const uint32_t i = threadIdx.x * 4;
      if(i < n) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(pd.pool + (size_t)chunk * CHUNK_BYTES) + threadIdx.x);
        rec[j][0] = v.x; rec[j][1] = v.y; rec[j][2] = v.z; rec[j][3] = v.w;
        nv[j] = min(4u, n - i);
      }
    }
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

// ---- one CTA per window: probe in shared memory -----------------------------------------------------
template<int KW>
__global__ void __launch_bounds__(WIN_NTH) win_insert_kernel(TableDev T, PartDev pd, WinDev wd, const uint64_t* __restrict__ inv_lut_g, uint32_t nbytes) {
  extern __shared__ __align__(16) uint32_t win[];
  const uint32_t fb = T.fbits, rb = T.rbits, hb = fb - rb;
  const uint32_t fmask = (1u << fb) - 1u, one = 1u << fb, cb = 32 - fb;
  const uint32_t hmask = hb ? ((1u << hb) - 1u) : 0u;
  uint32_t* tab = (uint32_t*)T.slots;
  const uint32_t n_tasks = wd.G << wd.wpr_lg;
  uint32_t n_ins = 0, n_new = 0, n_rep = 0;
  for(uint32_t task = blockIdx.x; task < n_tasks; task += gridDim.x) {
    const uint32_t b = wd.wstart[task], e = wd.wstart[task + 1];
    if(b == e) continue;                               // (uniform over the CTA)
    const uint64_t slot_base = ((uint64_t)(wd.g0 + (task >> wd.wpr_lg)) << pd.region_bits) + ((uint64_t)(task & ((1u << wd.wpr_lg) - 1)) << WIN_LG);
    uint4* gw = reinterpret_cast<uint4*>(tab + slot_base);
    uint4* sw = reinterpret_cast<uint4*>(win);
    for(uint32_t i = threadIdx.x; i < WIN_SLOTS / 4; i += WIN_NTH) sw[i] = __ldcs(gw + i);
    __syncthreads();
    // Every lane keeps one record in flight and performs ONE probe per trip of the loop; a lane whose record
    // is settled takes the next record of its stride at once (the one after it is already on its way from
    // memory), so lanes with long probe sequences do not idle the rest of the warp.
    {
      uint32_t i = b + threadIdx.x;
      bool have = i < e;
      uint32_t rec = have ? __ldcs(wd.wrec + i) : 0u;
      bool have_n = have && i + WIN_NTH < e;
      uint32_t nxt = have_n ? __ldcs(wd.wrec + i + WIN_NTH) : 0u;
      uint32_t local = (hb < 32 ? rec >> hb : 0u) & (WIN_SLOTS - 1), high = rec & hmask, kf0 = high << rb;
      uint32_t at = local, p = 0;
      while(have) {
        bool done = false;
        if(at >= WIN_SLOTS) {                              // leaves the window: the global path takes it after this kernel
          const unsigned long long d = atomicAdd(wd.def_n, 1ull);
          if(d < wd.def_cap) { wd.def_pos[d] = slot_base + local; wd.def_high[d] = high; }
          else atomicAdd(&T.stats[STAT_POOL_FULL], 1ull);
          done = true;
        } else {
          const uint32_t kf = kf0 | (p + 1);
          const uint32_t o = atomicCAS(&win[at], 0u, kf | one);
          if(o == 0u) { ++n_new; ++n_ins; n_rep += p; done = true; }
          else if((o & fmask) == kf) {
            const uint32_t o2 = atomicAdd(&win[at], one);
            if((((o2 >> fb) + 1) >> cb) != 0) k2_carry(T.ovf_keys, T.ovf_vals, T.ovf_mask, T.stats, slot_base + at);
            ++n_ins; n_rep += p; done = true;
          } else if(p >= T.max_reprobe) {
            k2_fail<KW>(T.shard_index, T.local_lsize, T.lsize, T.stats, T.fail_keys, T.fail_counts, T.fail_cap, slot_base + local, high, inv_lut_g, nbytes);
            done = true;
          } else { ++p; at += p; }                         // pos + i(i+1)/2
        }
        if(done) {
          rec = nxt; have = have_n; i += WIN_NTH;
          have_n = have && i + WIN_NTH < e;
          if(have_n) nxt = __ldcs(wd.wrec + i + WIN_NTH);
          local = (hb < 32 ? rec >> hb : 0u) & (WIN_SLOTS - 1); high = rec & hmask; kf0 = high << rb;
          at = local; p = 0;
        }
      }
    }
    __syncthreads();
    for(uint32_t i = threadIdx.x; i < WIN_SLOTS / 4; i += WIN_NTH) __stcs(gw + i, sw[i]);
    __syncthreads();
  }
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
