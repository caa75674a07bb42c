// jf_kernels.cuh -- the sm_100a kernels of the counting pipeline.
//
//   K0a nl_scan_kernel      last '\n' of every tile (two ranges, see tile_state_kernel)
//   K0b tile_state_kernel   parser state at the start of every staged window
//   K1  count_kernel        fused: TMA-staged FASTA window -> classify -> compact symbols
//                           -> rolling canonical k-mers -> GF(2) hash -> CAS insert/increment
//                           (or, in ROUTE mode, bucket by owning shard for the all-to-all)
//   K2  insert_keys_kernel  packed keys -> hash -> insert (multi-GPU receive side, regrow)
//   K3  collect_kernel      table segment -> (key, count, sort key) records
//   K4  serialize_kernel    sorted records -> on-disk record bytes
//   K5  lookup / histogram / synth_fasta
#ifndef JF_KERNELS_CUH
#define JF_KERNELS_CUH
#include "jf_device.cuh"

namespace jfk {

// ---------------------------------------------------------------------------------------
// K0a: per tile t (bytes [t*TILE, (t+1)*TILE)), position of the last '\n' in
//      A = [start, start+TILE-HALO) and in B = [start+TILE-HALO, start+TILE); -1 if none.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nl_scan_kernel(const uint8_t* __restrict__ in, uint64_t n, uint64_t n_tiles,
                                                      long long* __restrict__ nlA, long long* __restrict__ nlB) {
  __shared__ long long sA[8], sB[8];
  for(uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const uint64_t start = t * (uint64_t)TILE;
    const uint64_t end = min(n, start + (uint64_t)TILE);
    const uint64_t split = start + (uint64_t)(TILE - HALO);
    long long a = -1, b = -1;
    // 16-byte vectors; the tile start is 16-byte aligned
    const uint64_t nvec = (end - start) / 16;
    const uint4* v = reinterpret_cast<const uint4*>(in + start);
    for(uint64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
      uint4 x = v[i];
      uint32_t w[4] = { x.x, x.y, x.z, x.w };
#pragma unroll
      for(int j = 0; j < 4; ++j) {
        uint32_t y = w[j] ^ 0x0A0A0A0Au;
        if(((y - 0x01010101u) & ~y & 0x80808080u) != 0u) {     // some byte of w[j] is '\n'
#pragma unroll
          for(int q = 0; q < 4; ++q) {
            if(((w[j] >> (8 * q)) & 0xFFu) == 0x0Au) {
              long long p = (long long)(start + i * 16 + j * 4 + q);
              if((uint64_t)p < split) a = max(a, p); else b = max(b, p);
            }
          }
        }
      }
    }
    for(uint64_t p = start + nvec * 16 + threadIdx.x; p < end; p += blockDim.x) {
      if(in[p] == '\n') { if(p < split) a = max(a, (long long)p); else b = max(b, (long long)p); }
    }
#pragma unroll
    for(int o = 16; o; o >>= 1) {
      a = max(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if((threadIdx.x & 31) == 0) { sA[threadIdx.x >> 5] = a; sB[threadIdx.x >> 5] = b; }
    __syncthreads();
    if(threadIdx.x == 0) {
      for(int i = 1; i < (int)(blockDim.x >> 5); ++i) { a = max(a, sA[i]); b = max(b, sB[i]); }
      nlA[t] = a; nlB[t] = b;
    }
    __syncthreads();
  }
}

// state after the bytes [from, to) when entering them at a line start / with state s0
__device__ __forceinline__ uint32_t line_start_state(const uint8_t* in, uint64_t from, uint64_t to, uint32_t s0) {
  if(s0 != ST_L) return s0;
  uint64_t p = from;
  while(p < to && in[p] == '\r') ++p;
  if(p >= to) return ST_L;
  return in[p] == '>' ? ST_H : ST_S;
}

// ---------------------------------------------------------------------------------------
// K0b: window t starts at h_t = t*TILE - HALO.  Its entry state depends on the last '\n'
//      before h_t:  max(nlA[t-1], max_{u<t-1} max(nlA[u], nlB[u])).  One CTA, chunked scan:
//      the thread that owns tile u produces the state of window u+1.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) tile_state_kernel(const uint8_t* __restrict__ in, uint64_t n_tiles,
                                                          const long long* __restrict__ nlA, const long long* __restrict__ nlB,
                                                          const Carry* __restrict__ carry_in, uint8_t* __restrict__ tile_state) {
  __shared__ long long part[1024];
  const uint64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
  const uint64_t lo = min(n_tiles, per * threadIdx.x), hi = min(n_tiles, lo + per);
  long long m = -1;
  for(uint64_t u = lo; u < hi; ++u) m = max(m, max(nlA[u], nlB[u]));
  part[threadIdx.x] = m;
  __syncthreads();
  if(threadIdx.x == 0) {               // exclusive running max over the partials
    long long run = -1;
    for(int i = 0; i < (int)blockDim.x; ++i) { long long x = part[i]; part[i] = run; run = max(run, x); }
  }
  __syncthreads();
  long long mu = part[threadIdx.x];    // last newline in tiles < u   (u = lo initially)
  const uint32_t cstate = carry_in->state;
  if(threadIdx.x == 0 && n_tiles) tile_state[0] = (uint8_t)cstate;
  for(uint64_t u = lo; u < hi; ++u) {
    if(u + 1 < n_tiles) {
      const long long before = max(mu, nlA[u]);               // last newline before h_{u+1}
      const uint64_t h = (u + 1) * (uint64_t)TILE - HALO;
      tile_state[u + 1] = (uint8_t)(before >= 0 ? line_start_state(in, (uint64_t)before + 1, h, ST_L)
                                                : line_start_state(in, 0, h, cstate));
    }
    mu = max(mu, max(nlA[u], nlB[u]));
  }
}

// ---------------------------------------------------------------------------------------
// parser state machine helpers (semantics: mer_overlap_sequence_parser.hpp:161-185,260-287)
//   state L (line start): '\n','\r' stay; '>' -> H (new record: emits a window reset, the 'N'
//                          of :174-176); anything else -> S and is a sequence character
//   state S (in sequence): '\n' -> L; '\r' dropped when the run of '\r' ends in '\n'/EOF,
//                          otherwise it is a window reset; other bytes: base or reset
//   state H (in header)  : '\n' -> L; everything else ignored
// A transition function over {H,S,L} is packed 2 bits per input state.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fn_const(uint32_t s) { return s | (s << 2) | (s << 4); }
__device__ __forceinline__ uint32_t fn_apply(uint32_t f, uint32_t s) { return (f >> (2 * s)) & 3u; }
// first f, then g
__device__ __forceinline__ uint32_t fn_compose(uint32_t f, uint32_t g) {
  return fn_apply(g, fn_apply(f, 0)) | (fn_apply(g, fn_apply(f, 1)) << 2) | (fn_apply(g, fn_apply(f, 2)) << 4);
}
constexpr uint32_t FN_ID = 0u | (1u << 2) | (2u << 4);

// is the '\r' at position p (inside a sequence line) dropped?  True when the run of '\r'
// it belongs to is followed by '\n' or by the end of the file.
__device__ __forceinline__ bool cr_dropped(const uint8_t* in, uint64_t p, uint64_t n) {
  uint64_t q = p + 1;
  while(q < n && in[q] == '\r') ++q;
  return q >= n || in[q] == '\n';
}

struct CountArgs {
  const uint8_t* in;          // batch bytes (16-byte aligned)
  uint64_t       n;           // bytes in the batch
  uint64_t       n_look;      // bytes readable from `in` (>= n): look-ahead for '\r' runs
  uint64_t       n_tiles;
  const uint8_t* tile_state;
  const Carry*   carry_in;
  Carry*         carry_out;
  const uint64_t* lut;        // nb*256 hash table entries (global)
  uint32_t       k;
  uint32_t       canonical;
  uint32_t       nbytes;      // ceil(2k/8)
  uint32_t       mode;        // 0 = insert, 1 = route
  TableDev       T;
  // route mode
  uint64_t*      route_keys;
  unsigned long long* route_counts;
  uint64_t       route_cap;
  uint32_t       shard_bits;
};

// Slow path: walk backwards from byte position `end` (exclusive) of the batch collecting
// the symbols emitted before it, newest first, until `need` symbols or a reset is found.
// out[need-1-j] receives the j-th newest symbol, so out[0..need) ends up oldest-first;
// missing leading entries are SYM_BREAK.  `line_nl` = position of the last '\n' before `end`
// if known (>= -1), or -2 when unknown.
__device__ void backfill_symbols(const uint8_t* in, uint64_t n, const Carry* cin, long long end, long long line_nl,
                                 int need, uint8_t* out) {
  for(int i = 0; i < need; ++i) out[i] = SYM_BREAK;
  int got = 0;
  long long cur_end = end;           // exclusive end of the line piece under inspection
  long long q = line_nl;
  bool first = true;
  while(got < need) {
    if(!(first && q >= -1)) { q = cur_end - 1; while(q >= 0 && in[q] != '\n') --q; }
    first = false;
    const long long ls = q + 1;      // line start (0 when q == -1: batch start)
    // classify the line
    uint32_t s0 = (q >= 0) ? ST_L : cin->state;
    long long fnc = ls;              // first non-'\r' byte of the line piece (only meaningful when s0 == L)
    uint32_t st = s0;
    if(s0 == ST_L) {
      while(fnc < cur_end && in[fnc] == '\r') ++fnc;
      st = fnc >= cur_end ? ST_L : (in[fnc] == '>' ? ST_H : ST_S);
    }
    if(st == ST_H) return;           // a header precedes: window reset
    if(st == ST_S) {
      const long long seq_from = (s0 == ST_L) ? fnc : ls;
      for(long long p = cur_end - 1; p >= seq_from && got < need; --p) {
        uint32_t b = in[p];
        uint32_t sy;
        if(b == '\r') { if(cr_dropped(in, (uint64_t)p, n)) continue; sy = SYM_BREAK; }
        else sy = base_symbol(b);
        if(sy == SYM_BREAK) return;
        out[need - 1 - got] = (uint8_t)sy; ++got;
      }
      if(got >= need) return;
    }
    if(q < 0) {                      // reached the batch start: continue into the previous batch's carry
      for(int j = PRE - 1; j >= 0 && got < need; --j) {
        uint32_t sy = cin->sym[j];
        if(sy == SYM_BREAK) return;
        out[need - 1 - got] = (uint8_t)sy; ++got;
      }
      return;
    }
    cur_end = q;                     // continue with the previous line (the '\n' itself emits nothing)
  }
}

// ---------------------------------------------------------------------------------------
// K1: the fused counting kernel.  Persistent CTAs, double-buffered TMA windows.
// ---------------------------------------------------------------------------------------
struct __align__(16) CountSmem {
  uint8_t  win[2][WIN];            // TMA destinations
  uint8_t  sym[PRE + WIN + 16];    // carried prefix + compacted symbols of the window
  uint64_t bar[2];
  uint32_t warp_fn[NWARP];
  uint32_t warp_cnt[NWARP];
  uint32_t idx0, nsym, halo_break, total_state;
  unsigned long long red[8];
};

template<int KW, int SB>
__global__ void __launch_bounds__(NT, 2) count_kernel(const CountArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  CountSmem& sm = *reinterpret_cast<CountSmem*>(smem_raw);
  uint64_t* lut = reinterpret_cast<uint64_t*>(smem_raw + ((sizeof(CountSmem) + 15) & ~(size_t)15));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t k = a.k;
  const uint64_t n = a.n;

  // hash tables -> shared memory
  for(uint32_t i = tid; i < a.nbytes * 256u; i += NT) lut[i] = a.lut[i];
  if(tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); }
  __syncthreads();

  // issue the TMA copy of window t into buffer b (thread 0 only)
  auto issue = [&](uint64_t t, int b) {
    long long h = (long long)(t * (uint64_t)TILE) - HALO;
    long long from = h < 0 ? 0 : h;
    uint64_t avail = n - (uint64_t)from;
    uint64_t want = (uint64_t)((h + WIN) - from);
    uint32_t bytes = (uint32_t)((avail < want ? avail : want) & ~(uint64_t)15);
    if(bytes) {
      mbar_expect_tx(&sm.bar[b], bytes);
      tma_load_1d(&sm.win[b][from - h], a.in + from, bytes, &sm.bar[b]);
    } else {
      mbar_arrive(&sm.bar[b]);
    }
  };

  LocalStats ls = { 0, 0, 0, 0, 0 };
  uint32_t phase_bits = 0;          // bit b = parity to wait for on buffer b
  int buf = 0;
  uint64_t t = blockIdx.x;
  if(t < a.n_tiles && tid == 0) issue(t, 0);

  const uint64_t kmask_hi = (k * 2) % 64 ? ((1ull << ((k * 2) % 64)) - 1ull) : ~0ull;   // mask of the top key word
  for(; t < a.n_tiles; t += gridDim.x, buf ^= 1) {
    const uint64_t tn = t + gridDim.x;
    if(tn < a.n_tiles && tid == 0) issue(tn, buf ^ 1);

    const long long h = (long long)(t * (uint64_t)TILE) - HALO;      // global position of window byte 0
    const long long wend_ll = (long long)n < h + WIN ? (long long)n : h + WIN;
    // tail bytes that the 16-byte granular TMA copy left out
    {
      long long from = h < 0 ? 0 : h;
      long long copied = ((wend_ll - from) & ~15ll);
      long long g = from + copied + tid;
      if(tid < 16 && g < wend_ll) sm.win[buf][g - h] = a.in[g];
    }
    mbar_wait(&sm.bar[buf], (phase_bits >> buf) & 1u);
    phase_bits ^= 1u << buf;
    __syncthreads();

    // ---- phase B: classify 32 bytes per thread, build the state transition function ----
    const uint8_t* wb = sm.win[buf];
    uint32_t w[8];
    {
      const uint4* p4 = reinterpret_cast<const uint4*>(wb + tid * BPT);
      uint4 x0 = p4[0], x1 = p4[1];
      w[0] = x0.x; w[1] = x0.y; w[2] = x0.z; w[3] = x0.w; w[4] = x1.x; w[5] = x1.y; w[6] = x1.z; w[7] = x1.w;
    }
    const long long g0 = h + (long long)tid * BPT;      // global position of this thread's first byte
    int vlo = g0 < 0 ? (int)(-g0 < BPT ? -g0 : BPT) : 0;
    int vhi = (wend_ll - g0) < 0 ? 0 : ((wend_ll - g0) > BPT ? BPT : (int)(wend_ll - g0));
    if(vhi < vlo) vhi = vlo;

    uint32_t f;
    {
      uint32_t st = ST_L; bool seen_nl = false;
#pragma unroll
      for(int i = 0; i < BPT; ++i) {
        if(i >= vlo && i < vhi) {
          uint32_t b = (w[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
          if(b == '\n') { st = ST_L; seen_nl = true; }
          else if(st == ST_L && b != '\r') st = (b == '>') ? ST_H : ST_S;
        }
      }
      f = seen_nl ? fn_const(st) : ((uint32_t)ST_H | ((uint32_t)ST_S << 2) | (st << 4));
      if(vhi == vlo) f = FN_ID;
    }
    // block-wide exclusive scan of the transition functions
    uint32_t inc = f;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) {
      uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
      if(lane >= o) inc = fn_compose(up, inc);
    }
    if(lane == 31) sm.warp_fn[warp] = inc;
    __syncthreads();
    uint32_t entry = (t == 0) ? a.carry_in->state : (uint32_t)a.tile_state[t];
    uint32_t wpre = FN_ID;
    for(int i = 0; i < warp; ++i) wpre = fn_compose(wpre, sm.warp_fn[i]);
    uint32_t excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if(lane == 0) excl = FN_ID;
    uint32_t st_in = fn_apply(fn_compose(wpre, excl), entry);
    if(tid == NT - 1) sm.total_state = fn_apply(fn_compose(wpre, inc), entry);

    // ---- phase C: emit symbols (4 bits each, 32 max) ----
    uint64_t pk0 = 0, pk1 = 0;
    uint32_t cnt = 0; bool brk = false;
    {
      uint32_t st = st_in;
#pragma unroll
      for(int i = 0; i < BPT; ++i) {
        if(i >= vlo && i < vhi) {
          uint32_t b = (w[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
          uint32_t sy = 8;   // 8 = nothing
          if(st == ST_H) { if(b == '\n') st = ST_L; }
          else if(b == '\n') st = ST_L;
          else if(st == ST_L) {
            if(b == '\r') { }
            else if(b == '>') { st = ST_H; sy = SYM_BREAK; }
            else { st = ST_S; sy = base_symbol(b); }
          } else {           // ST_S
            if(b == '\r') { if(!cr_dropped(a.in, (uint64_t)(g0 + i), a.n_look)) sy = SYM_BREAK; }
            else sy = base_symbol(b);
          }
          if(sy != 8) {
            if(cnt < 16) pk0 |= (uint64_t)sy << (4 * cnt); else pk1 |= (uint64_t)sy << (4 * (cnt - 16));
            brk |= (sy == SYM_BREAK);
            ++cnt;
          }
        }
      }
    }
    // block-wide exclusive scan of counts
    uint32_t cinc = cnt;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) {
      uint32_t up = __shfl_up_sync(0xffffffffu, cinc, o);
      if(lane >= o) cinc += up;
    }
    if(lane == 31) sm.warp_cnt[warp] = cinc;
    if(tid == 0) sm.halo_break = 0;
    __syncthreads();
    uint32_t woff = 0;
    for(int i = 0; i < warp; ++i) woff += sm.warp_cnt[i];
    const uint32_t off = woff + cinc - cnt;
    if(tid == HALO / BPT) sm.idx0 = off;               // symbols emitted by the halo bytes
    if(tid == NT - 1) sm.nsym = off + cnt;
    if(tid < HALO / BPT && brk) sm.halo_break = 1;
    {
      uint8_t* dst = sm.sym + PRE + off;
      for(uint32_t j = 0; j < cnt; ++j) {
        uint32_t sy = (uint32_t)((j < 16 ? pk0 >> (4 * j) : pk1 >> (4 * (j - 16))) & 0xF);
        dst[j] = (uint8_t)sy;
      }
    }
    __syncthreads();
    const uint32_t idx0 = sm.idx0, nsym = sm.nsym;

    // ---- phase D: the k-1 symbols that precede the window ----
    if(warp == 0) {
      if(t == 0) {
        sm.sym[lane] = a.carry_in->sym[lane]; sm.sym[lane + 32] = a.carry_in->sym[lane + 32];
      } else {
        sm.sym[lane] = SYM_BREAK; sm.sym[lane + 32] = SYM_BREAK;
        __syncwarp();
        if(lane == 0 && a.tile_state[t] != ST_H && idx0 < k - 1 && !sm.halo_break) {
          // pathological input (very short lines / long runs of blank lines): exact slow path
          backfill_symbols(a.in, a.n_look, a.carry_in, h, -2, PRE, sm.sym);
        }
      }
    }
    __syncthreads();

    // hand the parser state to the next batch
    if(t == a.n_tiles - 1 && warp == 1) {
      // (need at least PRE symbols of history: if this window is short, look further back)
      uint8_t* cs = a.carry_out->sym;
      if(nsym >= (uint32_t)PRE || t == 0 || sm.halo_break || a.tile_state[t] == ST_H) {
        cs[lane] = sm.sym[nsym + lane]; cs[lane + 32] = sm.sym[nsym + lane + 32];
      } else if(lane == 0) {
        backfill_symbols(a.in, a.n_look, a.carry_in, (long long)n, -2, PRE, cs);
      }
      if(lane == 0) a.carry_out->state = sm.total_state;
    }

    // ---- phase E: roll canonical k-mers over the compacted symbols, hash, insert ----
    if(nsym > idx0) {
      const uint32_t n_chunks = (nsym - idx0 + QSYM - 1) / QSYM;
      for(uint32_t c = tid; c < n_chunks; c += NT) {
        const uint32_t j0 = idx0 + c * QSYM;
        const uint32_t j1 = min(nsym, j0 + (uint32_t)QSYM);
        uint64_t m[KW], rc[KW];
#pragma unroll
        for(int q = 0; q < KW; ++q) { m[q] = 0; rc[q] = 0; }
        uint32_t run = 0;
        const uint8_t* sp = sm.sym + PRE + j0 - (k - 1);
        const uint32_t total = (k - 1) + (j1 - j0);
        for(uint32_t j = 0; j < total; ++j) {
          const uint32_t sy = sp[j];
          if(sy < 4) {
            // m = (m << 2 | sy) & mask ; rc = rc >> 2 | (3 - sy) << (2k - 2)   (mer_dna.hpp:322-370)
            if(KW == 1) {
              m[0] = ((m[0] << 2) | sy) & kmask_hi;
              rc[0] = (rc[0] >> 2) | ((uint64_t)(3 - sy) << (2 * k - 2));
            } else {
              m[KW - 1] = ((m[KW - 1] << 2) | (m[0] >> 62)) & kmask_hi;
              m[0] = (m[0] << 2) | sy;
              rc[0] = (rc[0] >> 2) | (rc[KW - 1] << 62);
              rc[KW - 1] = (rc[KW - 1] >> 2) | ((uint64_t)(3 - sy) << ((2 * k - 2) - 64));
            }
            ++run;
          } else run = 0;
          if(j >= k - 1 && run >= k) {
            uint64_t key[KW];
            bool use_rc = false;
            if(a.canonical) {
              if(KW == 1) use_rc = rc[0] < m[0];
              else use_rc = (rc[KW - 1] < m[KW - 1]) || (rc[KW - 1] == m[KW - 1] && rc[0] < m[0]);
            }
#pragma unroll
            for(int q = 0; q < KW; ++q) key[q] = use_rc ? rc[q] : m[q];
            ls.kmers++;
            const uint64_t pos = gf2_hash<KW>(lut, key, (int)a.nbytes);
            if(a.mode == 0) {
              if(table_add<KW, SB>(a.T, key, pos, 1, ls)) ls.inserted++;
              else { ls.failed++; record_failure<KW>(a.T, key, 1); }
            } else {
              const uint32_t owner = a.shard_bits ? (uint32_t)(pos >> (a.T.lsize - a.shard_bits)) : 0u;
              unsigned long long at = atomicAdd(&a.route_counts[owner], 1ull);
              if(at < a.route_cap) {
#pragma unroll
                for(int q = 0; q < KW; ++q) a.route_keys[((uint64_t)owner * a.route_cap + at) * KW + q] = key[q];
              } else atomicAdd(&a.T.stats[STAT_ROUTE_DROPPED], 1ull);
            }
          }
        }
      }
    }
    __syncthreads();     // all reads of win[buf] and sym[] done before they are overwritten
  }

  // ---- statistics: one atomic per counter per CTA ----
  unsigned long long v[4] = { ls.kmers, ls.inserted, ls.distinct, ls.reprobes };
#pragma unroll
  for(int q = 0; q < 4; ++q) {
#pragma unroll
    for(int o = 16; o; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  __shared__ unsigned long long part[NWARP][4];
  if(lane == 0) { part[warp][0] = v[0]; part[warp][1] = v[1]; part[warp][2] = v[2]; part[warp][3] = v[3]; }
  __syncthreads();
  if(tid < 4) {
    unsigned long long s = 0;
    for(int i = 0; i < NWARP; ++i) s += part[i][tid];
    const int which[4] = { STAT_KMERS, STAT_INSERTED, STAT_DISTINCT, STAT_REPROBES };
    if(s) atomicAdd(&a.T.stats[which[tid]], s);
  }
}

// ---------------------------------------------------------------------------------------
// K2: packed keys (+ optional counts) -> hash -> insert
// ---------------------------------------------------------------------------------------
template<int KW, int SB>
__global__ void __launch_bounds__(256) insert_keys_kernel(TableDev T, const uint64_t* __restrict__ lut_g, uint32_t nbytes,
                                                          const uint64_t* __restrict__ keys, const uint64_t* __restrict__ counts,
                                                          uint64_t n) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* lut = reinterpret_cast<uint64_t*>(smem_raw);
  for(uint32_t i = threadIdx.x; i < nbytes * 256u; i += blockDim.x) lut[i] = lut_g[i];
  __syncthreads();
  LocalStats ls = { 0, 0, 0, 0, 0 };
  for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t key[KW];
#pragma unroll
    for(int q = 0; q < KW; ++q) key[q] = keys[i * KW + q];
    const uint64_t cnt = counts ? counts[i] : 1;
    const uint64_t pos = gf2_hash<KW>(lut, key, (int)nbytes);
    if(table_add<KW, SB>(T, key, pos, cnt, ls)) ls.inserted++;
    else { ls.failed++; record_failure<KW>(T, key, cnt); }
  }
  unsigned long long v[3] = { ls.inserted, ls.distinct, ls.reprobes };
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

// ---------------------------------------------------------------------------------------
// K3: collect the records whose ORIGINAL position lies in [seg_lo, seg_hi) (local slot
//     indices).  They live in slots [seg_lo, seg_hi + reprobe margin).  The full key is
//     rebuilt from the stored high bits and the position with the inverse matrix
//     (large_hash_iterator.hpp:164-170; large_hash_array.hpp:851-858).
//     sort key = ((opos - seg_lo) << hb) | high   -- ascending (position, key) order because
//     two keys sharing a position differ in their high bits.
// ---------------------------------------------------------------------------------------
struct CollectArgs {
  TableDev T;
  const uint64_t* inv_lut;     // byte tables of the inverse matrix (low lsize bits of the key)
  uint32_t nbytes;
  uint64_t seg_lo, seg_hi;     // local original positions
  uint64_t scan_hi;            // exclusive end of the slots to scan
  uint64_t lower, upper;       // count filter
  uint32_t hb;                 // number of explicit key bits (2k - lsize, >= 0)
  uint64_t* out_keys;          // KW words per record
  uint64_t* out_counts;
  uint64_t* out_sort_lo;       // low 64 bits of the sort key
  uint64_t* out_sort_hi;       // remaining bits (KW == 2 only; may be NULL)
  unsigned long long* out_n;
  uint64_t out_cap;
};

template<int KW, int SB>
__global__ void __launch_bounds__(256) collect_kernel(const CollectArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* lut = reinterpret_cast<uint64_t*>(smem_raw);
  for(uint32_t i = threadIdx.x; i < a.nbytes * 256u; i += blockDim.x) lut[i] = a.inv_lut[i];
  __syncthreads();
  const TableDev& T = a.T;
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t span = a.scan_hi - a.seg_lo;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t rounds = (span + stride - 1) / stride;
  for(uint64_t r = 0; r < rounds; ++r) {
    const uint64_t i = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t idx = a.seg_lo + i;
    bool have = false;
    u128 high; uint32_t rp = 0; uint64_t cnt = 0; uint64_t opos = 0;
    if(i < span && slot_decode<SB>(T, idx, high, rp, cnt)) {
      opos = idx - (rp ? tri(rp) : 0);
      if(opos >= a.seg_lo && opos < a.seg_hi) {
        if(T.stats[STAT_OVERFLOWED]) {
          const uint32_t cb = (SB == 128) ? (64 - (T.fbits > 64 ? T.fbits - 64 : 0)) : (SB - T.fbits);
          const uint64_t carries = ovf_get(T, idx);
          if(carries) {
            // saturate at 2^64-1 like a 64-bit counter would
            if(cb >= 64 || (carries >> (64 - cb)) != 0) cnt = ~0ull;
            else { uint64_t add = carries << cb; cnt = (cnt + add < cnt) ? ~0ull : cnt + add; }
          }
        }
        have = cnt >= a.lower && cnt <= a.upper;
      }
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, have);
    if(ballot) {
      unsigned long long basei = 0;
      if(lane == 0) basei = atomicAdd(a.out_n, (unsigned long long)__popc(ballot));
      basei = __shfl_sync(0xffffffffu, basei, 0);
      if(have) {
        const uint64_t o = basei + __popc(ballot & ((1u << lane) - 1u));
        if(o < a.out_cap) {
          // global position = shard bits : local original position
          const uint64_t gpos = ((uint64_t)T.shard_index << T.local_lsize) | opos;
          // vector fed to the inverse matrix: [high bits of key : position]
          uint64_t v[KW];
          if(KW == 1) v[0] = (T.lsize >= 64 ? 0 : (high.lo << T.lsize)) | gpos;
          else {
            v[0] = (T.lsize >= 64 ? 0 : (high.lo << T.lsize)) | gpos;
            v[KW - 1] = T.lsize ? ((high.hi << T.lsize) | (high.lo >> (64 - T.lsize))) : high.hi;
          }
          const uint64_t low = gf2_hash<KW>(lut, v, (int)a.nbytes);
          const uint64_t lmask = T.lsize >= 64 ? ~0ull : ((1ull << T.lsize) - 1ull);
          uint64_t key[KW];
#pragma unroll
          for(int q = 0; q < KW; ++q) key[q] = v[q];
          key[0] = (key[0] & ~lmask) | (low & lmask);
#pragma unroll
          for(int q = 0; q < KW; ++q) a.out_keys[o * KW + q] = key[q];
          a.out_counts[o] = cnt;
          const uint64_t rel = opos - a.seg_lo;
          if(KW == 1 || a.out_sort_hi == nullptr) {
            a.out_sort_lo[o] = (a.hb >= 64 ? 0 : (rel << a.hb)) | high.lo;
          } else {
            // 128-bit sort key (rel << hb) | high
            uint64_t lo = high.lo, hi = high.hi;
            if(a.hb < 64) { lo |= rel << a.hb; hi |= a.hb ? (rel >> (64 - a.hb)) : 0; }
            else hi |= rel << (a.hb - 64);
            a.out_sort_lo[o] = lo; a.out_sort_hi[o] = hi;
          }
        }
      }
    }
  }
}

// gather helper for the two-pass 128-bit sort
__global__ void gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t* __restrict__ dst, uint64_t n) {
  for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}
__global__ void iota_u32_kernel(uint32_t* __restrict__ dst, uint64_t n) {
  for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = (uint32_t)i;
}

// ---------------------------------------------------------------------------------------
// K4: record bytes = first ceil(2k/8) bytes of the little-endian key words, then
//     min(count, 2^(8*ocl)-1) as ocl little-endian bytes (binary_dumper.hpp:36-40).
//     A CTA assembles its records in shared memory and writes them with coalesced words.
// ---------------------------------------------------------------------------------------
template<int KW>
__global__ void __launch_bounds__(256) serialize_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ counts,
                                                        const uint32_t* __restrict__ perm, uint64_t n, uint32_t key_bytes,
                                                        uint32_t ocl, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t stage[];
  const uint32_t rec = key_bytes + ocl;
  const uint64_t maxv = ocl >= 8 ? ~0ull : ((1ull << (8 * ocl)) - 1ull);
  const uint64_t per_block = 256;
  for(uint64_t b0 = (uint64_t)blockIdx.x * per_block; b0 < n; b0 += (uint64_t)gridDim.x * per_block) {
    const uint64_t i = b0 + threadIdx.x;
    if(i < n) {
      const uint64_t src = perm ? perm[i] : i;
      uint8_t* d = stage + threadIdx.x * rec;
      uint64_t kw[KW];
#pragma unroll
      for(int q = 0; q < KW; ++q) kw[q] = keys[src * KW + q];
      for(uint32_t b = 0; b < key_bytes; ++b) d[b] = (uint8_t)(kw[b >> 3] >> ((b & 7) * 8));
      uint64_t c = counts[src]; if(c > maxv) c = maxv;
      for(uint32_t b = 0; b < ocl; ++b) d[key_bytes + b] = (uint8_t)(c >> (8 * b));
    }
    __syncthreads();
    const uint64_t nrec = min((uint64_t)per_block, n - b0);
    const uint64_t nb = nrec * rec;
    uint8_t* o = out + b0 * rec;
    for(uint64_t j = threadIdx.x; j < nb; j += blockDim.x) o[j] = stage[j];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// K5: lookups / histogram straight from the resident table
// ---------------------------------------------------------------------------------------
template<int KW, int SB>
__global__ void __launch_bounds__(256) lookup_kernel(TableDev T, const uint64_t* __restrict__ lut_g, uint32_t nbytes,
                                                     const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ vals,
                                                     uint32_t shard_bits) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* lut = reinterpret_cast<uint64_t*>(smem_raw);
  for(uint32_t i = threadIdx.x; i < nbytes * 256u; i += blockDim.x) lut[i] = lut_g[i];
  __syncthreads();
  for(uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t key[KW];
#pragma unroll
    for(int q = 0; q < KW; ++q) key[q] = keys[i * KW + q];
    const uint64_t pos = gf2_hash<KW>(lut, key, (int)nbytes);
    uint64_t res = 0;
    const uint32_t owner = shard_bits ? (uint32_t)(pos >> (T.lsize - shard_bits)) : 0u;
    if(owner == T.shard_index) {
      const u128 want = key_high<KW>(key, T.lsize);
      const uint64_t base = pos & T.local_mask;
      uint64_t idx = base;
      for(uint32_t r = 0; r <= T.max_reprobe; ++r) {
        u128 high; uint32_t rp; uint64_t cnt;
        if(!slot_decode<SB>(T, idx, high, rp, cnt)) break;          // empty slot ends the probe sequence
        if(rp == r && high.lo == want.lo && high.hi == want.hi) {
          const uint32_t cb = (SB == 128) ? (64 - (T.fbits > 64 ? T.fbits - 64 : 0)) : (SB - T.fbits);
          const uint64_t carries = T.stats[STAT_OVERFLOWED] ? ovf_get(T, idx) : 0;
          if(carries) {
            if(cb >= 64 || (carries >> (64 - cb)) != 0) cnt = ~0ull;
            else { uint64_t add = carries << cb; cnt = (cnt + add < cnt) ? ~0ull : cnt + add; }
          }
          res = cnt;
          break;
        }
        idx = base + tri(r + 1);
      }
    }
    vals[i] = res;
  }
}

template<int SB>
__global__ void __launch_bounds__(256) histogram_kernel(TableDev T, uint64_t n_slots, unsigned long long* __restrict__ hist, uint32_t n_bins) {
  for(uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_slots; idx += (uint64_t)gridDim.x * blockDim.x) {
    u128 high; uint32_t rp; uint64_t cnt;
    if(slot_decode<SB>(T, idx, high, rp, cnt)) {
      if(T.stats[STAT_OVERFLOWED]) {
        const uint32_t cb = (SB == 128) ? (64 - (T.fbits > 64 ? T.fbits - 64 : 0)) : (SB - T.fbits);
        const uint64_t carries = ovf_get(T, idx);
        if(carries) {
          if(cb >= 64 || (carries >> (64 - cb)) != 0) cnt = ~0ull;
          else { uint64_t add = carries << cb; cnt = (cnt + add < cnt) ? ~0ull : cnt + add; }
        }
      }
      atomicAdd(&hist[cnt < n_bins ? cnt : n_bins - 1], 1ull);
    }
  }
}

template<int SB>
__global__ void __launch_bounds__(256) max_count_kernel(TableDev T, uint64_t n_slots) {
  unsigned long long m = 0;
  for(uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_slots; idx += (uint64_t)gridDim.x * blockDim.x) {
    u128 high; uint32_t rp; uint64_t cnt;
    if(slot_decode<SB>(T, idx, high, rp, cnt)) {
      if(T.stats[STAT_OVERFLOWED]) {
        const uint32_t cb = (SB == 128) ? (64 - (T.fbits > 64 ? T.fbits - 64 : 0)) : (SB - T.fbits);
        const uint64_t carries = ovf_get(T, idx);
        if(carries) {
          if(cb >= 64 || (carries >> (64 - cb)) != 0) cnt = ~0ull;
          else { uint64_t add = carries << cb; cnt = (cnt + add < cnt) ? ~0ull : cnt + add; }
        }
      }
      m = max(m, (unsigned long long)cnt);
    }
  }
#pragma unroll
  for(int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if((threadIdx.x & 31) == 0 && m) atomicMax(&T.stats[STAT_MAXCOUNT], m);
}

// ---------------------------------------------------------------------------------------
// synthetic FASTA of the shape jellyfish/generate_sequence.cc:119-149 writes:
// ">read1\n" then 70 bases per line.  Base i is drawn from a counter-based generator
// (splitmix64 of seed + i/32 gives 32 bases), so any byte range can be generated independently.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
constexpr int SYNTH_HDR = 7;     // strlen(">read1\n")
constexpr int SYNTH_LINE = 70;

__global__ void __launch_bounds__(256) synth_fasta_kernel(uint8_t* __restrict__ out, uint64_t n_bytes, uint64_t n_bases, uint64_t seed) {
  const char hdr[SYNTH_HDR + 1] = ">read1\n";
  for(uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_bytes; p += (uint64_t)gridDim.x * blockDim.x) {
    uint8_t c;
    if(p < SYNTH_HDR) c = (uint8_t)hdr[p];
    else {
      const uint64_t q = p - SYNTH_HDR;
      const uint64_t line = q / (SYNTH_LINE + 1), col = q % (SYNTH_LINE + 1);
      if(col == SYNTH_LINE) c = '\n';
      else {
        const uint64_t i = line * SYNTH_LINE + col;
        if(i >= n_bases) c = '\n';
        else {
          const uint64_t r = splitmix64(seed + (i >> 5));
          c = (uint8_t)"ACGT"[(r >> (2 * (i & 31))) & 3];
        }
      }
    }
    out[p] = c;
  }
}

}  // namespace jfk
#endif
