// jf_engine.cu -- C-ABI implementation (include/jfgpu.h) of the B200 k-mer counting engine.
// Host orchestration only: every byte of the hot path is processed by the kernels in
// jf_kernels.cuh.  There is no CPU fallback; without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jfgpu.h"
#include "host/jf_matrix.hpp"
#include "jf_kernels.cuh"
#include "jf_extract.cuh"
#include "jf_window.cuh"
#include "jf_dump.cuh"
#include "jf_shard.cuh"

using namespace jfk;

static std::atomic<unsigned long long> g_launches(0);
static thread_local std::string g_create_error;

#define JF_LAUNCHED() g_launches.fetch_add(1, std::memory_order_relaxed)

namespace {

constexpr unsigned SHARD_RESERVED_SMS = 16;   // SMs K1 leaves to NCCL while an exchange runs beside it

unsigned ceil_log2(uint64_t x) { unsigned l = 0; while(l < 64 && ((uint64_t)1 << l) < x) ++l; return l; }
unsigned bitsize(uint64_t x) { unsigned b = 0; while(x) { ++b; x >>= 1; } return b ? b : 1; }

struct DevBuf {
  void* p = nullptr; size_t bytes = 0;
  cudaError_t alloc(size_t n) { free(); bytes = n; return n ? cudaMalloc(&p, n) : cudaSuccess; }
  void free() { if(p) cudaFree(p); p = nullptr; bytes = 0; }
  template<typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// byte-indexed tables of a GF(2) matrix: entry [b*256+v] = product with the vector whose
// byte b equals v (reference column order: bit i selects columns[c-1-i],
// rectangular_binary_matrix.hpp:223-261)
std::vector<uint64_t> build_lut(const jfb::gf2_matrix& m, unsigned nbytes) {
  std::vector<uint64_t> lut((size_t)nbytes * 256, 0);
  const unsigned c = m.c(), r = m.r();
  for(unsigned b = 0; b < nbytes; ++b) {
    uint64_t col[8];
    for(unsigned j = 0; j < 8; ++j) {
      unsigned i = 8 * b + j;
      if(i >= c) col[j] = 0;
      else if(m.is_identity()) col[j] = i < r ? ((uint64_t)1 << i) : 0;
      else col[j] = m[c - 1 - i];
    }
    for(unsigned v = 0; v < 256; ++v) {
      uint64_t x = 0;
      for(unsigned j = 0; j < 8; ++j) if(v & (1u << j)) x ^= col[j];
      lut[(size_t)b * 256 + v] = x;
    }
  }
  return lut;
}

struct Table {
  unsigned lsize = 0, local_lsize = 0, max_reprobe = 0, rbits = 1, fbits = 1, slot_bits = 32, hb = 0;
  uint64_t size = 0, local_size = 0, margin = 0, local_slots = 0, ovf_size = 0;
  jfb::gf2_matrix M, Minv;
  DevBuf slots, lut, inv_lut, ovf_keys, ovf_vals, lut11;
  uint64_t prow[8] = {0,0,0,0,0,0,0,0}; unsigned n_prow = 0; bool hash_fast = false;
  std::vector<uint64_t> reprobes;
  void release() { slots.free(); lut.free(); inv_lut.free(); ovf_keys.free(); ovf_vals.free(); lut11.free(); }
  size_t bytes() const { return (size_t)local_slots * (slot_bits / 8); }
};

}  // namespace

struct PartState {
  uint32_t P = 0, region_bits = 0, rec_bytes = 0, cap = 0, flush_min = 0, stage_bytes = 0, n_chunks = 0, margin = 0, arena_chunks = 0, n_arenas = 0;
  DevBuf pool, dir, order, pool_next, cta_chunk, cta_fill, spill_keys, spill_counts, spill_n, hist, start, cursor, unit_cursor;
  uint64_t spill_cap = 0;
  // window form of K2 (jf_window.cuh)
  DevBuf w_start, w_cursor, w_cnt, w_rec, w_def_pos, w_def_high, w_def_n;
  uint64_t w_rec_cap = 0, w_def_cap = 0;
  uint64_t bound_chunks = 0;     // host-side upper bound of the chunks in use in any one arena
  bool pending = false;          // records sit in the pool
};

// Sharded counting, record exchange (jf_shard.cuh): the send pool (global regions, one chunk arena per owning shard, two
// banks) and the receive pool live in buffers the caller registers (they are the NCCL send / receive buffers).
struct ShardState {
  bool on = false;
  uint32_t P = 0, sbits = 0, own_regions = 0, split_lg = 0, owner_shift = 0;
  uint64_t arena_chunks = 0, seg_chunks = 0;
  uint8_t* send_pool = nullptr; uint2* send_dir = nullptr; uint8_t* recv_pool = nullptr; uint2* recv_dir = nullptr;
  DevBuf pool_next[2], cta_chunk, cta_fill;
  unsigned int* h_counts = nullptr;       // pinned
};

struct BloomState {
  uint32_t mode = BLOOM_NONE, k = 0;
  uint64_t m = 0, inv = 0, n_words = 0;
  bool drawn = false;              // the two hash matrices have been drawn (lazily: after the --if pass, count_main.cc:288-321)
  jfb::gf2_matrix M1, M2;
  std::vector<uint64_t> cols1, cols2;
  DevBuf bits, locks, lut1, lut2;
  void release() { bits.free(); locks.free(); lut1.free(); lut2.free(); }
};

struct jfgpu_engine {
  jfgpu_params p;
  PartState part;
  ShardState sh;
  BloomState bloom;
  int device = 0;
  unsigned k = 0, kw = 1, nbytes = 0, shard_bits = 0;
  cudaStream_t cs = nullptr, hs = nullptr;
  int n_sm = 148;
  jfb::glibc_random rng;
  Table tab;
  DevBuf stats, carry[2], fail_keys[2], fail_counts[2];
  uint64_t fail_cap = 0;
  int carry_cur = 0, fail_cur = 0;
  unsigned long long* h_stats = nullptr;    // pinned mirror
  // staging for host feeds
  size_t batch_bytes = 0;
  DevBuf stage[2]; cudaEvent_t ev_copied[2] = { nullptr, nullptr }, ev_done[2] = { nullptr, nullptr };
  int stage_cur = 0;
  // per-batch scratch
  DevBuf nlA, nlB, cntA, cntB, tstate; uint64_t scratch_tiles = 0;
  int format = 0;                 // 0 = FASTA, 1 = FASTQ: format of the file being fed
  // -Q on FASTQ: a batch must end on a record boundary (the qualities of a read are looked up two lines further down);
  // lines seen so far in the file (mod 4) and the incomplete last record of the previous feed
  uint32_t q_lines = 0; std::string q_tail;
  // --disk: what hash_counter::handle_full_ary does when the table cannot double (hash_counter.hpp:187-192): the caller's
  // hook dumps the resident table (jfgpu_dump from inside the hook), the engine zeroes it and goes on counting
  jfgpu_spill_fn spill_fn = nullptr; void* spill_ctx = nullptr; bool in_spill = false; uint64_t spills = 0;
  uint32_t op = 0;                // JFGPU_OP_*
  // bookkeeping
  bool in_file = false;
  uint64_t bytes_fed = 0, regrows = 0;
  double count_ms = 0;
  unsigned eff_val_len = 7;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  std::string err;
  std::vector<uint64_t> matrix_cols_host;   // for jfgpu_table_info_get
  std::vector<cudaEvent_t> kev;             // event pairs around count_kernel launches
  size_t kev_used = 0;
  double kernel_ms = 0; uint64_t kernel_launches = 0;
  double drain_ms = 0; cudaEvent_t ev_d0 = nullptr, ev_d1 = nullptr;
  int count_smem = 0;
  // CUDA events around the window kernels of a drain: [4 per group] hist begin, scatter begin, insert begin, insert end
  std::vector<cudaEvent_t> wev; size_t wev_used = 0;
  double win_ms[3] = { 0, 0, 0 };
  // failure counter watched one group behind (hash_counter::add -> handle_full_ary), without draining the stream
  unsigned long long* h_watch = nullptr; cudaEvent_t ev_watch[2] = { nullptr, nullptr };
};

namespace {

int fail(jfgpu_engine* e, int code, const std::string& msg) {
  if(e) e->err = msg; else g_create_error = msg;
  return code;
}
#define CUDA_OK(e, call) do { cudaError_t _c = (call); if(_c != cudaSuccess) \
  return fail(e, JFGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_c)); } while(0)

TableDev table_dev(const jfgpu_engine* e, const Table& t) {
  TableDev d;
  memset(&d, 0, sizeof(d));
  d.slots = t.slots.p;
  d.local_mask = t.local_size - 1;
  d.local_lsize = t.local_lsize;
  d.lsize = t.lsize;
  d.shard_index = e->p.shard_index;
  d.kbits = 2 * e->k;
  d.rbits = t.rbits;
  d.fbits = t.fbits;
  d.max_reprobe = t.max_reprobe;
  d.op = e->op;
  d.ovf_keys = t.ovf_keys.as<unsigned long long>();
  d.ovf_vals = t.ovf_vals.as<unsigned long long>();
  d.ovf_mask = t.ovf_size - 1;
  d.stats = e->stats.as<unsigned long long>();
  d.fail_keys = e->fail_keys[e->fail_cur].as<uint64_t>();
  d.fail_counts = e->fail_counts[e->fail_cur].as<uint64_t>();
  d.fail_cap = e->fail_cap;
  return d;
}

// Geometry of a table of 2^lsize GLOBAL slots -- large_hash_array.hpp:150-173,29-39.
// `reprobe_limit` is -p for the first table and the CLIPPED limit of the previous table after a doubling: the reference
// hands ary_->max_reprobe() to the new array (hash_counter.hpp:205-209), so a limit clipped by a tiny table never grows back.
int table_setup(jfgpu_engine* e, Table& t, unsigned lsize, const jfb::gf2_matrix& M, unsigned reprobe_limit) {
  const unsigned kbits = 2 * e->k;
  t.lsize = lsize;
  t.size = (uint64_t)1 << lsize;
  t.local_lsize = lsize - e->shard_bits;
  t.local_size = (uint64_t)1 << t.local_lsize;
  t.hb = kbits > lsize ? kbits - lsize : 0;
  unsigned limit = kbits > lsize ? reprobe_limit : 0;
  // reprobes[0] = 1, reprobes[i] = i(i+1)/2 (lib/storage.cc:13-41); clip so that reprobes[limit] < size
  auto rp = [](unsigned i) -> uint64_t { return i == 0 ? 1 : tri(i); };
  while(limit >= 1 && rp(limit) >= t.size) --limit;
  t.max_reprobe = limit;
  t.reprobes.resize(limit + 1);
  for(unsigned i = 0; i <= limit; ++i) t.reprobes[i] = rp(i);
  t.rbits = bitsize(limit + 1);
  t.fbits = t.hb + t.rbits;
  // at least 10 counter bits in a slot, so that only counts beyond ~1000 need the carry side table
  if(t.fbits <= 22) t.slot_bits = 32;
  else if(t.fbits <= 56) t.slot_bits = 64;
  else if(t.fbits <= 120) t.slot_bits = 128;
  else return fail(e, JFGPU_ERR_ARG, "key too long for this table size (key field > 120 bits)");
  if(e->kw == 2 && t.slot_bits == 32) t.slot_bits = 64;
  t.margin = limit ? tri(limit) : 0;
  t.local_slots = t.local_size + t.margin + 8;
  t.M = M;
  t.Minv = M.pseudo_inverse();
  if(cudaMalloc(&t.slots.p, t.bytes()) != cudaSuccess) {
    cudaGetLastError();
    t.slots.p = nullptr;
    char buf[128]; snprintf(buf, sizeof(buf), "Failed to allocate %zu bytes of device memory", t.bytes());
    return fail(e, JFGPU_ERR_NOMEM, buf);
  }
  t.slots.bytes = t.bytes();
  CUDA_OK(e, cudaMemsetAsync(t.slots.p, 0, t.bytes(), e->cs));
  // counter-carry side table: one entry per slot whose counter field wrapped; sized with the table
  t.ovf_size = (uint64_t)1 << 20;
  while(t.ovf_size < ((uint64_t)1 << 26) && t.ovf_size * 64 < t.local_size) t.ovf_size <<= 1;
  CUDA_OK(e, t.ovf_keys.alloc(t.ovf_size * 8));
  CUDA_OK(e, t.ovf_vals.alloc(t.ovf_size * 8));
  CUDA_OK(e, cudaMemsetAsync(t.ovf_keys.p, 0, t.ovf_size * 8, e->cs));
  CUDA_OK(e, cudaMemsetAsync(t.ovf_vals.p, 0, t.ovf_size * 8, e->cs));
  std::vector<uint64_t> l1 = build_lut(t.M, e->nbytes), l2 = build_lut(t.Minv, e->nbytes);
  CUDA_OK(e, t.lut.alloc(l1.size() * 8));
  CUDA_OK(e, t.inv_lut.alloc(l2.size() * 8));
  CUDA_OK(e, cudaMemcpyAsync(t.lut.p, l1.data(), l1.size() * 8, cudaMemcpyHostToDevice, e->cs));
  CUDA_OK(e, cudaMemcpyAsync(t.inv_lut.p, l2.data(), l2.size() * 8, cudaMemcpyHostToDevice, e->cs));
  // fast hash tables: 2k <= 44 -> four 11-bit chunks; entries = low 32 bits of the partial products,
  // position bits 32.. come from parity rows
  t.hash_fast = false; t.n_prow = 0;
  for(unsigned i = 0; i < 8; ++i) t.prow[i] = 0;            // (unused rows must be zero: K1 evaluates a fixed number of them)
  std::vector<uint32_t> l11;
  if(e->kw == 1 && kbits <= 44 && lsize <= 40) {
    l11.assign(4 * 2048, 0);
    auto colsel = [&](unsigned i) -> uint64_t {       // contribution of key bit i
      if(i >= t.M.c()) return 0;
      if(t.M.is_identity()) return i < t.M.r() ? ((uint64_t)1 << i) : 0;
      return t.M[t.M.c() - 1 - i];
    };
    for(unsigned tb = 0; tb < 4; ++tb)
      for(unsigned v = 0; v < 2048; ++v) {
        uint64_t x = 0;
        for(unsigned j = 0; j < 11; ++j) if(v & (1u << j)) x ^= colsel(tb * 11 + j);
        l11[tb * 2048 + v] = (uint32_t)x;
      }
    for(unsigned ob = 32; ob < lsize; ++ob) {
      uint64_t row = 0;
      for(unsigned i = 0; i < kbits; ++i) if((colsel(i) >> ob) & 1) row |= (uint64_t)1 << i;
      t.prow[t.n_prow++] = row;
    }
    CUDA_OK(e, t.lut11.alloc(l11.size() * 4));
    CUDA_OK(e, cudaMemcpyAsync(t.lut11.p, l11.data(), l11.size() * 4, cudaMemcpyHostToDevice, e->cs));
    t.hash_fast = true;
  }
  CUDA_OK(e, cudaStreamSynchronize(e->cs));     // the host vectors are about to go out of scope
  return JFGPU_OK;
}

// The matrix large_hash::array draws for a table of 2^lsize slots (large_hash_array.hpp:992-1002)
jfb::gf2_matrix draw_matrix(jfgpu_engine* e, uint64_t requested_size, unsigned lsize) {
  const unsigned kbits = 2 * e->k;
  const bool smaller = kbits >= 64 || requested_size < ((uint64_t)1 << kbits);
  if(!smaller) return jfb::gf2_matrix::identity(kbits);
  jfb::gf2_matrix m(lsize, kbits);
  return m.randomize_pseudo_inverse(e->rng);
}

template<typename F>
int dispatch(jfgpu_engine* e, unsigned kw, unsigned sb, F&& f) {
  if(kw == 1 && sb == 32)  return f(std::integral_constant<int, 1>(), std::integral_constant<int, 32>());
  if(kw == 1 && sb == 64)  return f(std::integral_constant<int, 1>(), std::integral_constant<int, 64>());
  if(kw == 1 && sb == 128) return f(std::integral_constant<int, 1>(), std::integral_constant<int, 128>());
  if(kw == 2 && sb == 64)  return f(std::integral_constant<int, 2>(), std::integral_constant<int, 64>());
  if(kw == 2 && sb == 128) return f(std::integral_constant<int, 2>(), std::integral_constant<int, 128>());
  return fail(e, JFGPU_ERR_ARG, "unsupported key/slot combination");
}

template<int NTH>
size_t count_smem_bytes(size_t lut_bytes, size_t stage_bytes, size_t bloom_bytes = 0, bool fast = false) {
  const size_t part = fast ? (size_t)RING_P * 8 + (size_t)RING_P * RING * 4 : (stage_bytes ? PMAX * 4 + stage_bytes : 0);
  return ((sizeof(ExtractSmemT<NTH>) + 15) & ~(size_t)15) + lut_bytes + part + bloom_bytes;
}

int ensure_scratch(jfgpu_engine* e, uint64_t n_tiles) {
  if(n_tiles <= e->scratch_tiles) return JFGPU_OK;
  uint64_t want = std::max<uint64_t>(n_tiles, 1024);
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  CUDA_OK(e, e->nlA.alloc(want * 8));
  CUDA_OK(e, e->nlB.alloc(want * 8));
  CUDA_OK(e, e->cntA.alloc(want * 4));
  CUDA_OK(e, e->cntB.alloc(want * 4));
  CUDA_OK(e, e->tstate.alloc(want));
  e->scratch_tiles = want;
  return JFGPU_OK;
}

// ---- partitioned insertion: geometry, pool, drain --------------------------------------
// the ring memory of the staging kernels (RING_P * RING records) shared out among P regions
uint32_t ring_len_for(uint32_t P) {
  uint32_t p2 = 1; while(p2 < P) p2 <<= 1;
  const uint32_t len = p2 >= RING_P ? RING : RING * (RING_P / p2);
  return std::min<uint32_t>(len, 1024);
}
PartDev part_dev(const jfgpu_engine* e) {
  PartDev d;
  memset(&d, 0, sizeof(d));
  const PartState& ps = e->part;
  d.P = ps.P; d.region_bits = ps.region_bits; d.rec_bytes = ps.rec_bytes; d.cap = ps.cap; d.flush_min = ps.flush_min;
  d.chunk_recs = CHUNK_BYTES / std::max(1u, ps.rec_bytes); d.n_chunks = ps.n_chunks; d.stage_bytes = ps.stage_bytes;
  d.margin = ps.margin;
  d.arena_chunks = ps.arena_chunks;
  d.ring_len = ring_len_for(ps.P ? ps.P : 1);
  d.pool = ps.pool.as<uint8_t>(); d.pool_next = ps.pool_next.as<unsigned int>(); d.n_units = d.pool_next + ps.n_arenas; d.dir = ps.dir.as<uint2>();
  d.cta_chunk = ps.cta_chunk.as<uint32_t>(); d.cta_fill = ps.cta_fill.as<uint32_t>();
  d.spill_keys = ps.spill_keys.as<uint64_t>(); d.spill_counts = ps.spill_counts.as<uint64_t>();
  d.spill_n = ps.spill_n.as<unsigned long long>(); d.spill_cap = ps.spill_cap;
  return d;
}

// The send pool of sharded counting as K1 sees it: regions of the GLOBAL table, arenas by owning shard.
PartDev shard_send_dev(const jfgpu_engine* e, int bank) {
  PartDev d;
  memset(&d, 0, sizeof(d));
  const ShardState& sh = e->sh;
  const uint32_t G = e->p.n_shards;
  d.P = sh.P; d.region_bits = sh.sbits; d.rec_bytes = 4; d.chunk_recs = CHUNK_BYTES / 4;
  d.n_chunks = (uint32_t)(sh.arena_chunks * G); d.arena_chunks = (uint32_t)sh.arena_chunks;
  d.by_owner = 1; d.owner_shift = sh.owner_shift;
  d.ring_len = ring_len_for(sh.P);
  // a chunk is closed once it might not take the records of one more ring pass (4 k-mers per thread of K1)
  { const double mean = 4096.0 / sh.P; d.margin = std::max<uint32_t>(2 * RING, (uint32_t)(mean + 6.0 * sqrt(mean) + 8.0)); }
  d.pool = sh.send_pool + (size_t)bank * G * sh.arena_chunks * CHUNK_BYTES;
  d.dir = sh.send_dir + (size_t)bank * G * sh.arena_chunks;
  d.pool_next = sh.pool_next[bank].as<unsigned int>(); d.n_units = d.pool_next + G;
  d.cta_chunk = sh.cta_chunk.as<uint32_t>(); d.cta_fill = sh.cta_fill.as<uint32_t>();
  return d;
}

// Decide whether (and how) the current table is filled region by region.
void part_configure(jfgpu_engine* e) {
  PartState& ps = e->part;
  const Table& t = e->tab;
  ps.P = 0;
  if(e->p.no_partition || t.bytes() < ((size_t)(e->p.part_min_mb ? e->p.part_min_mb : 256) << 20)) return;       // small tables live in L2 anyway
  uint32_t P = 256;
  const size_t region_target = (size_t)(e->p.region_mb ? e->p.region_mb : 64) << 20;
  const size_t owned_bytes = (size_t)t.local_size * (t.slot_bits / 8);           // (without the overflow margin)
  while(P < (uint32_t)PMAX && (owned_bytes / P) > region_target) P <<= 1;
  for(;; P >>= 1) {
    if(P < 64 || t.local_lsize < 8 || (1u << (t.local_lsize - 8)) < P) return;
    const uint32_t region_bits = t.local_lsize - ceil_log2(P);
    const uint32_t bits = region_bits + t.hb;
    const uint32_t rec = bits <= 32 ? 4 : bits <= 64 ? 8 : bits <= 128 ? 16 : 0;
    if(!rec) return;
    // records arriving per region between two roll-over passes (one per window): 1024 threads x 32 symbols / P
    const double mean = 1024.0 * 32 / P;
    const uint32_t margin = (uint32_t)(mean + 6.0 * sqrt(mean) + 8.0);
    if(margin * 2 > CHUNK_BYTES / rec) return;               // chunks would be closed half empty: insert directly
    ps.P = P; ps.region_bits = region_bits; ps.rec_bytes = rec; ps.cap = 0; ps.flush_min = 0; ps.margin = margin;
    ps.stage_bytes = PMAX * 4;                               // the open-chunk ids (the counters are accounted separately)
    return;
  }
}

int part_alloc(jfgpu_engine* e) {
  PartState& ps = e->part;
  if(ps.pool.p) return JFGPU_OK;
  size_t free_b = 0, total_b = 0;
  CUDA_OK(e, cudaMemGetInfo(&free_b, &total_b));
  size_t want = e->p.pool_bytes ? (size_t)e->p.pool_bytes : std::min<size_t>((size_t)(free_b * 0.7), (size_t)64 << 30);
  const size_t floor_b = (size_t)e->n_sm * ps.P * CHUNK_BYTES * 2;        // every CTA keeps one open chunk per region
  if(want < floor_b) want = floor_b;
  // one arena per CTA of the staging kernels (persistent, one CTA per SM)
  ps.n_arenas = (uint32_t)e->n_sm;
  ps.arena_chunks = (uint32_t)std::min<size_t>(want / CHUNK_BYTES / ps.n_arenas, 0xFFFFFFF0u / ps.n_arenas);
  ps.n_chunks = ps.arena_chunks * ps.n_arenas;
  ps.spill_cap = (uint64_t)16 << 20;
  bool ok = ps.pool.alloc((size_t)ps.n_chunks * CHUNK_BYTES) == cudaSuccess && ps.dir.alloc((size_t)ps.n_chunks * 8) == cudaSuccess &&
            ps.order.alloc((size_t)ps.n_chunks * 4) == cudaSuccess && ps.pool_next.alloc(((size_t)ps.n_arenas + 2) * 4) == cudaSuccess &&
            ps.cta_chunk.alloc((size_t)e->n_sm * PMAX * 4) == cudaSuccess && ps.cta_fill.alloc((size_t)e->n_sm * PMAX * 4) == cudaSuccess &&
            ps.spill_keys.alloc(ps.spill_cap * 8 * e->kw) == cudaSuccess && ps.spill_counts.alloc(ps.spill_cap * 8) == cudaSuccess &&
            ps.spill_n.alloc(8) == cudaSuccess && ps.hist.alloc(PMAX * 4) == cudaSuccess && ps.start.alloc(PMAX * 4) == cudaSuccess &&
            ps.cursor.alloc(PMAX * 4) == cudaSuccess && ps.unit_cursor.alloc(8) == cudaSuccess;
  if(!ok) { cudaGetLastError(); return fail(e, JFGPU_ERR_NOMEM, "device allocation of the record pool failed"); }
  CUDA_OK(e, cudaMemsetAsync(ps.pool_next.p, 0, ps.pool_next.bytes, e->cs));
  CUDA_OK(e, cudaMemsetAsync(ps.spill_n.p, 0, 8, e->cs));
  CUDA_OK(e, cudaMemsetAsync(ps.cta_chunk.p, 0xFF, ps.cta_chunk.bytes, e->cs));
  CUDA_OK(e, cudaMemsetAsync(ps.cta_fill.p, 0, ps.cta_fill.bytes, e->cs));
  ps.bound_chunks = ps.P;
  ps.pending = false;
  CUDA_OK(e, cudaStreamSynchronize(e->cs));     // callers may continue on another stream
  return JFGPU_OK;
}

void part_release(jfgpu_engine* e) {
  PartState& ps = e->part;
  ps.pool.free(); ps.dir.free(); ps.order.free(); ps.pool_next.free(); ps.cta_chunk.free(); ps.cta_fill.free();
  ps.spill_keys.free(); ps.spill_counts.free(); ps.spill_n.free(); ps.hist.free(); ps.start.free(); ps.cursor.free(); ps.unit_cursor.free();
  ps.w_start.free(); ps.w_cursor.free(); ps.w_cnt.free(); ps.w_rec.free(); ps.w_def_pos.free(); ps.w_def_high.free(); ps.w_def_n.free();
  ps.w_rec_cap = ps.w_def_cap = 0;
  ps.n_chunks = 0; ps.arena_chunks = 0; ps.n_arenas = 0; ps.pending = false;
}

int regrow(jfgpu_engine* e);
int spill_table(jfgpu_engine* e, uint64_t n_failed);
int read_stats(jfgpu_engine* e);
int bloom_draw(jfgpu_engine* e);
BloomDev bloom_dev(const jfgpu_engine* e);

// The window form of K2 (jf_window.cuh): the default for 32-bit slots and 4-byte records.
// Processes whole regions in groups, starting at unit `*done` (which must be the first unit of a
// region), until every unit is inserted or -- with regrow enabled -- a group has reported keys that
// found no slot.  Regions too large for the group buffer are left to the L2 kernel.
// Post a copy of the live failure counter behind the work enqueued so far / wait for an earlier one.
static void watch_post(jfgpu_engine* e, cudaStream_t st, int slot) {
  cudaMemcpyAsync(e->h_watch + slot, e->stats.as<unsigned long long>() + STAT_FAILED, 8, cudaMemcpyDeviceToHost, st);
  cudaEventRecord(e->ev_watch[slot], st);
}
static bool watch_failed(jfgpu_engine* e, int slot) {
  cudaEventSynchronize(e->ev_watch[slot]);
  return e->h_watch[slot] != 0;
}
static cudaEvent_t win_event(jfgpu_engine* e, cudaStream_t st) {
  if(e->wev_used == e->wev.size()) { cudaEvent_t ev; cudaEventCreate(&ev); e->wev.push_back(ev); }
  cudaEvent_t ev = e->wev[e->wev_used++];
  cudaEventRecord(ev, st);
  return ev;
}
// fold the event quadruples of the finished drain into win_ms (the stream must be idle)
static void resolve_win_events(jfgpu_engine* e) {
  for(size_t i = 0; i + 3 < e->wev_used; i += 4)
    for(int q = 0; q < 3; ++q) {
      float ms = 0;
      if(cudaEventElapsedTime(&ms, e->wev[i + q], e->wev[i + q + 1]) == cudaSuccess) e->win_ms[q] += ms; else cudaGetLastError();
    }
  e->wev_used = 0;
}
static bool window_enabled(jfgpu_engine* e, const PartDev& pd) {
  return e->p.k2_mode == 0 && e->op == 0 && e->tab.slot_bits == 32 && pd.rec_bytes == 4 && pd.region_bits > WIN_LG &&
         pd.region_bits - WIN_LG <= 11 && CHUNK_BYTES == WIN_NTH * 16;
}
int read_stats(jfgpu_engine* e);
static int window_drain(jfgpu_engine* e, cudaStream_t st, const PartDev& pd, unsigned n_units, bool careful, unsigned group_units,
                        unsigned* done, bool* failed) {
  PartState& ps = e->part;
  *failed = false;
  const uint32_t wpr_lg = pd.region_bits - WIN_LG;
  if(!ps.w_rec.p) {
    ps.w_rec_cap = (uint64_t)64 << 20;                       // records per group (256 MB)
    ps.w_def_cap = (uint64_t)16 << 20;
    bool ok = ps.w_rec.alloc(ps.w_rec_cap * 4 + 64) == cudaSuccess && ps.w_start.alloc((((size_t)WIN_MAX_G << 11) + 1) * 4) == cudaSuccess &&
              ps.w_cursor.alloc(((size_t)WIN_MAX_G << 11) * 4) == cudaSuccess && ps.w_cnt.alloc(((size_t)WIN_MAX_G << 11) * 4) == cudaSuccess && ps.w_def_pos.alloc(ps.w_def_cap * 8) == cudaSuccess &&
              ps.w_def_high.alloc(ps.w_def_cap * 4) == cudaSuccess && ps.w_def_n.alloc(8) == cudaSuccess;
    if(!ok) { cudaGetLastError(); return fail(e, JFGPU_ERR_NOMEM, "device allocation of the window buffers failed"); }
    CUDA_OK(e, cudaMemsetAsync(ps.w_def_n.p, 0, 8, st));
  }
  // first unit of every region (chunk_scan_kernel wrote it), on the host
  std::vector<uint32_t> start(pd.P + 1);
  CUDA_OK(e, cudaMemcpyAsync(start.data(), ps.start.p, (size_t)pd.P * 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(e, cudaStreamSynchronize(st));
  start[pd.P] = n_units;
  for(uint32_t r = 0; r < pd.P; ++r) start[r] = std::min(start[r], n_units);
  uint32_t r0 = 0;
  while(r0 < pd.P && start[r0] < *done) ++r0;
  const size_t scatter_smem = ((size_t)4 * ((size_t)1 << wpr_lg) + (size_t)WIN_ST_UNITS * pd.chunk_recs) * 4;
  cudaFuncSetAttribute(win_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem);
  // (the runs of a group start on 16-byte boundaries: up to 3 padding records per window)
  const uint64_t max_units = std::min<uint64_t>((ps.w_rec_cap - ((uint64_t)WIN_MAX_G << 13)) / pd.chunk_recs, careful ? group_units : 0xFFFFFFFFu);
  uint32_t gi = 0;
  while(r0 < pd.P && *done < n_units) {
    WinDev wd;
    memset(&wd, 0, sizeof(wd));
    uint32_t G = 0, tiles = 0, stiles = 0;
    while(r0 + G < pd.P && G < WIN_MAX_G) {
      const uint32_t nu = start[r0 + G + 1] - start[r0 + G];
      if((uint64_t)(start[r0 + G + 1] - start[r0]) > max_units) break;
      wd.tile_first[G] = tiles; wd.stile_first[G] = stiles; wd.unit_first[G] = start[r0 + G];
      tiles += (nu + WIN_TILE_UNITS - 1) / WIN_TILE_UNITS;
      stiles += (nu + WIN_ST_UNITS - 1) / WIN_ST_UNITS;
      ++G;
    }
    TableDev T = table_dev(e, e->tab);
    if(G == 0) {
      // a single region holds more records than the group buffer (heavily repeated k-mers): L2 kernel for it
      const unsigned upto = start[r0 + 1];
      cudaMemsetAsync(ps.unit_cursor.p, 0, 8, st);
      if(e->kw == 1) insert_chunks32_kernel<1><<<e->n_sm * 2, 768, 0, st>>>(T, pd, ps.order.as<uint32_t>(), ps.unit_cursor.as<unsigned int>(), *done, upto, e->tab.inv_lut.as<uint64_t>(), e->nbytes);
      else           insert_chunks32_kernel<2><<<e->n_sm * 2, 768, 0, st>>>(T, pd, ps.order.as<uint32_t>(), ps.unit_cursor.as<unsigned int>(), *done, upto, e->tab.inv_lut.as<uint64_t>(), e->nbytes);
      JF_LAUNCHED();
      *done = upto; r0 += 1;
    } else {
      wd.tile_first[G] = tiles; wd.stile_first[G] = stiles; wd.unit_first[G] = start[r0 + G];
      wd.g0 = r0; wd.G = G; wd.wpr_lg = wpr_lg; wd.n_tiles = tiles;
      wd.wstart = ps.w_start.as<uint32_t>(); wd.wcursor = ps.w_cursor.as<uint32_t>(); wd.wcnt = ps.w_cnt.as<uint32_t>();
      wd.wrec = ps.w_rec.as<uint32_t>(); wd.wrec_cap = ps.w_rec_cap;
      wd.def_pos = ps.w_def_pos.as<uint64_t>(); wd.def_high = ps.w_def_high.as<uint32_t>();
      wd.def_n = ps.w_def_n.as<unsigned long long>(); wd.def_cap = ps.w_def_cap;
      const uint32_t hb = e->tab.fbits - e->tab.rbits;
      if(tiles) {
        CUDA_OK(e, cudaMemsetAsync(ps.w_cursor.p, 0, ((size_t)G << wpr_lg) * 4, st));
        win_event(e, st);
        win_hist_kernel<<<tiles, WIN_NTH, 0, st>>>(pd, wd, ps.order.as<uint32_t>(), hb); JF_LAUNCHED();
        win_scan_kernel<<<1, 1024, 0, st>>>(wd, T.stats); JF_LAUNCHED();
        win_event(e, st);
        win_scatter_kernel<<<stiles, WIN_ST_NTH, scatter_smem, st>>>(pd, wd, ps.order.as<uint32_t>(), hb); JF_LAUNCHED();
        win_event(e, st);
        if(e->kw == 1) {
          cudaFuncSetAttribute(win_insert2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WIN2_SMEM);
          win_insert2_kernel<1><<<e->n_sm, WIN2_NTH, WIN2_SMEM, st>>>(T, pd, wd, e->tab.inv_lut.as<uint64_t>(), e->nbytes); JF_LAUNCHED();
          win_deferred_kernel<1><<<e->n_sm * 2, 256, 0, st>>>(T, wd, e->tab.inv_lut.as<uint64_t>(), e->nbytes); JF_LAUNCHED();
        } else {
          cudaFuncSetAttribute(win_insert2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WIN2_SMEM);
          win_insert2_kernel<2><<<e->n_sm, WIN2_NTH, WIN2_SMEM, st>>>(T, pd, wd, e->tab.inv_lut.as<uint64_t>(), e->nbytes); JF_LAUNCHED();
          win_deferred_kernel<2><<<e->n_sm * 2, 256, 0, st>>>(T, wd, e->tab.inv_lut.as<uint64_t>(), e->nbytes); JF_LAUNCHED();
        }
        win_event(e, st);
        CUDA_OK(e, cudaMemsetAsync(ps.w_def_n.p, 0, 8, st));
      }
      *done = start[r0 + G]; r0 += G;
    }
    if(careful) {
      // the failure counter is looked at one group late, so that the device never waits for the host: two groups of
      // failed keys fit the failure list (group = fail_cap / 2 records)
      watch_post(e, st, (int)(gi & 1));
      if(gi > 0 && watch_failed(e, (int)((gi - 1) & 1))) { cudaStreamSynchronize(st); *failed = true; return JFGPU_OK; }
      ++gi;
    }
  }
  if(careful && gi > 0 && watch_failed(e, (int)((gi - 1) & 1))) { cudaStreamSynchronize(st); *failed = true; return JFGPU_OK; }
  CUDA_OK(e, cudaGetLastError());
  return JFGPU_OK;
}

// Insert everything that sits in the record pool (K1b), region by region, then the spill list.
// With regrow enabled the chunks go in groups small enough for the failure list, and the
// failure counter is checked after each group (hash_counter::add -> handle_full_ary).
int part_drain(jfgpu_engine* e, cudaStream_t st) {
  PartState& ps = e->part;
  if(!ps.P || !ps.pool.p || !ps.pending) return JFGPU_OK;
  PartDev pd = part_dev(e);
  const int g = e->n_sm * 4;
  if(!e->ev_d0) { cudaEventCreate(&e->ev_d0); cudaEventCreate(&e->ev_d1); }
  cudaEventRecord(e->ev_d0, st);
  close_chunks_kernel<<<g, 256, 0, st>>>(pd, (uint32_t)e->n_sm); JF_LAUNCHED();
  CUDA_OK(e, cudaMemsetAsync(ps.hist.p, 0, PMAX * 4, st));
  chunk_hist_kernel<<<g, 256, 0, st>>>(pd, ps.hist.as<uint32_t>()); JF_LAUNCHED();
  chunk_scan_kernel<<<1, 1024, 0, st>>>(pd.P, ps.hist.as<uint32_t>(), ps.start.as<uint32_t>(), ps.cursor.as<uint32_t>(), pd.n_units); JF_LAUNCHED();
  chunk_scatter_kernel<<<g, 256, 0, st>>>(pd, ps.cursor.as<uint32_t>(), ps.order.as<uint32_t>()); JF_LAUNCHED();
  CUDA_OK(e, cudaMemsetAsync(ps.unit_cursor.p, 0, 8, st));
  // geometry the records were written with (a regrow in the middle changes e->tab)
  const TableDev T0 = table_dev(e, e->tab);
  const unsigned sb0 = e->tab.slot_bits;
  const bool careful = e->p.allow_regrow != 0 || e->spill_fn != nullptr;
  unsigned int n_units = 0;
  if(careful) {
    CUDA_OK(e, cudaMemcpyAsync(&n_units, pd.n_units, 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(e, cudaStreamSynchronize(st));
    n_units = std::min(n_units, ps.n_chunks);
  }
  const unsigned group = careful ? (unsigned)std::max<uint64_t>(1, e->fail_cap / 2 / pd.chunk_recs) : 0xFFFFFFFFu;
  int rc = JFGPU_OK;
  DevBuf old_inv;     // inverse tables of the geometry the records belong to, once the table has been rebuilt
  unsigned done = 0;
  bool rebuilt = false;
  if(window_enabled(e, pd)) {
    if(!careful) {
      CUDA_OK(e, cudaMemcpyAsync(&n_units, pd.n_units, 4, cudaMemcpyDeviceToHost, st));
      CUDA_OK(e, cudaStreamSynchronize(st));
      n_units = std::min(n_units, ps.n_chunks);
    }
    bool w_failed = false;
    rc = window_drain(e, st, pd, n_units, careful, group, &done, &w_failed);
    if(!rc && w_failed) {        // same as below: keep the old inverse tables, rebuild, the rest goes through the rehash kernel
      if(old_inv.alloc(e->tab.inv_lut.bytes) != cudaSuccess) { cudaGetLastError(); rc = fail(e, JFGPU_ERR_NOMEM, "device allocation failed"); }
      else {
        cudaMemcpyAsync(old_inv.p, e->tab.inv_lut.p, e->tab.inv_lut.bytes, cudaMemcpyDeviceToDevice, e->cs);
        cudaStreamSynchronize(e->cs);
        rebuilt = true;
        rc = regrow(e);
      }
    }
  }
  unsigned gq = 0;
  if(!rc && !(window_enabled(e, pd) && done >= n_units && !rebuilt))
  do {
    const unsigned upto = careful ? (unsigned)std::min<uint64_t>((uint64_t)done + group, n_units) : 0xFFFFFFFFu;
    cudaMemsetAsync(ps.unit_cursor.p, 0, 8, st);
    TableDev T = table_dev(e, e->tab);
    if(!rebuilt) {
      if(e->op == 0 && e->tab.slot_bits == 32 && pd.rec_bytes == 4 && e->p.k2_mode != 2) {
        // lean 32-bit specialisation: 2 CTAs x 1024 threads per SM
        if(e->kw == 1) insert_chunks32_kernel<1><<<e->n_sm * 2, 768, 0, st>>>(T, pd, ps.order.as<uint32_t>(), ps.unit_cursor.as<unsigned int>(), done, upto, e->tab.inv_lut.as<uint64_t>(), e->nbytes);
        else           insert_chunks32_kernel<2><<<e->n_sm * 2, 768, 0, st>>>(T, pd, ps.order.as<uint32_t>(), ps.unit_cursor.as<unsigned int>(), done, upto, e->tab.inv_lut.as<uint64_t>(), e->nbytes);
        rc = JFGPU_OK;
      } else
      rc = dispatch(e, e->kw, e->tab.slot_bits, [&](auto KW, auto SB) -> int {
        insert_chunks_kernel<decltype(KW)::value, decltype(SB)::value><<<e->n_sm * 2, 512, 0, st>>>(
            T, pd, ps.order.as<uint32_t>(), ps.unit_cursor.as<unsigned int>(), done, upto, e->tab.inv_lut.as<uint64_t>(), e->nbytes);
        return JFGPU_OK;
      });
    } else {
      const size_t smem = (size_t)e->nbytes * 256 * 8 * 2;
      rc = dispatch(e, e->kw, e->tab.slot_bits, [&](auto KW, auto SB) -> int {
        auto kern = rehash_chunks_kernel<decltype(KW)::value, decltype(SB)::value>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<e->n_sm, 512, smem, st>>>(T, T0, pd, ps.order.as<uint32_t>(), ps.unit_cursor.as<unsigned int>(), done, upto,
                                            old_inv.as<uint64_t>(), e->tab.lut.as<uint64_t>(), e->nbytes);
        return JFGPU_OK;
      });
    }
    if(rc) break;
    JF_LAUNCHED();
    if(!careful) break;
    done = upto;
    // the failure counter is read one group late (two groups of failed keys fit the failure list), so the device
    // does not idle while the host looks at it; the last group is checked right away
    watch_post(e, st, (int)(gq & 1));
    const bool last = done >= n_units;
    bool failed_now = gq > 0 && watch_failed(e, (int)((gq - 1) & 1));
    if(!failed_now && last) failed_now = watch_failed(e, (int)(gq & 1));
    ++gq;
    if(failed_now) {
      cudaStreamSynchronize(st);
      gq = 0;                    // the groups launched so far are complete: start the look-behind afresh
      if(!rebuilt) {            // keep a copy of the inverse tables the pending records were written against
        if(old_inv.alloc(e->tab.inv_lut.bytes) != cudaSuccess) { cudaGetLastError(); rc = fail(e, JFGPU_ERR_NOMEM, "device allocation failed"); break; }
        cudaMemcpyAsync(old_inv.p, e->tab.inv_lut.p, e->tab.inv_lut.bytes, cudaMemcpyDeviceToDevice, e->cs);
        cudaStreamSynchronize(e->cs);
        rebuilt = true;
      }
      rc = regrow(e);
      if(rc) break;
    }
  } while(done < n_units);
  (void)sb0;
  if(!rc) {
    // the spilled keys carry full keys: plain insertion into whatever the table is now
    TableDev T = table_dev(e, e->tab);
    const size_t smem = (size_t)e->nbytes * 256 * 8;
    rc = dispatch(e, e->kw, e->tab.slot_bits, [&](auto KW, auto SB) -> int {
      auto kern = insert_spill_kernel<decltype(KW)::value, decltype(SB)::value>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<e->n_sm * 2, 256, smem, st>>>(T, e->tab.lut.as<uint64_t>(), e->nbytes, pd);
      return JFGPU_OK;
    });
    if(!rc) JF_LAUNCHED();
  }
  cudaMemsetAsync(ps.pool_next.p, 0, ps.pool_next.bytes, st);
  cudaMemsetAsync(ps.spill_n.p, 0, 8, st);
  cudaEventRecord(e->ev_d1, st);
  cudaStreamSynchronize(st);
  { float ms = 0; if(cudaEventElapsedTime(&ms, e->ev_d0, e->ev_d1) == cudaSuccess) e->drain_ms += ms; else cudaGetLastError(); }
  resolve_win_events(e);
  ps.pending = false;
  if(rebuilt) { cudaStreamSynchronize(st); old_inv.free(); part_configure(e); if(!e->part.P) part_release(e); }
  ps.bound_chunks = ps.P;
  if(rc) return rc;
  CUDA_OK(e, cudaGetLastError());
  return JFGPU_OK;
}

// no more text per launch than an empty arena can take (small pools: tests, tables that leave little memory)
size_t part_cap_len(const jfgpu_engine* e, size_t len) {
  const PartState& ps = e->part;
  if(!ps.P || !ps.arena_chunks) return len;
  const uint64_t usable = CHUNK_BYTES / ps.rec_bytes - ps.margin;
  const uint64_t room = ps.arena_chunks > ps.P + 8 ? ps.arena_chunks - ps.P - 8 : 1;
  const uint64_t tiles_per_cta = std::max<uint64_t>(room * usable / (1024 * 32), 1);
  const uint64_t cap = tiles_per_cta * (1024 * 32 - HALO) * (uint64_t)e->n_sm / 2;
  return len > cap ? (size_t)std::max<uint64_t>(cap & ~(uint64_t)15, 16) : len;
}

// the quality threshold in force: the PRIME pass of --if reads its files without it (count_main.cc:289-295 uses mer_counter there)
static uint32_t eff_min_qual(const jfgpu_engine* e) { return e->op == JFGPU_OP_PRIME ? 0u : e->p.min_qual; }

// One batch of device-resident text through K0a, K0b, K1 on `stream`.
int run_batch(jfgpu_engine* e, const uint8_t* dev, uint64_t n, uint64_t n_look, cudaStream_t stream,
              int mode, uint64_t* route_keys, unsigned long long* route_counts, uint64_t route_cap, uint64_t n_back = 0) {
  if(n == 0) return JFGPU_OK;
  PartState& ps = e->part;
  const bool bc_build = e->bloom.mode == BLOOM_COUNT;
  const bool part = mode == 0 && ps.P != 0 && !bc_build;
  int rc;
  if(e->bloom.mode != BLOOM_NONE && !e->bloom.drawn && (e->op != JFGPU_OP_PRIME || bc_build)) { rc = bloom_draw(e); if(rc) return rc; }
  if(part) {
    rc = part_alloc(e);
    if(rc) return rc;
    // conservative host-side bound on the use of any one arena: a CTA sees ceil(tiles / CTAs) tiles, one record per
    // input byte at most, plus the chunks its roll-over passes may leave nearly empty
    const uint64_t usable = CHUNK_BYTES / ps.rec_bytes - ps.margin;         // records a closed chunk holds at least
    const uint64_t tile_b = 1024 * 32 - HALO;
    const uint64_t tiles = (n + tile_b - 1) / tile_b;
    const uint64_t per_cta = (tiles + e->n_sm - 1) / e->n_sm * tile_b;
    const uint64_t need = per_cta / usable + 2;
    if(ps.bound_chunks + need > ps.arena_chunks) {
      rc = part_drain(e, stream);
      if(rc) return rc;
      if(!e->part.P) return run_batch(e, dev, n, n_look, stream, mode, route_keys, route_counts, route_cap, n_back);
    }
    if(ps.bound_chunks + need > ps.arena_chunks) return fail(e, JFGPU_ERR_NOMEM, "record pool smaller than one batch");
    ps.bound_chunks += need;
    ps.pending = true;
  }
  const bool shard_send = mode == 3;            // K1 writes region records of the GLOBAL table into the send pool (bank = route_cap)
  const uint32_t tile = (part || shard_send ? 1024 : 512) * 32 - HALO;
  const uint64_t n_tiles = (n + tile - 1) / tile;
  rc = ensure_scratch(e, n_tiles);
  if(rc) return rc;
  const int g0 = (int)std::min<uint64_t>(n_tiles, (uint64_t)e->n_sm * 8);
  nl_scan_kernel<<<g0, 256, 0, stream>>>(dev, n, n_tiles, tile, e->nlA.as<long long>(), e->nlB.as<long long>(), e->cntA.as<uint32_t>(), e->cntB.as<uint32_t>());
  JF_LAUNCHED();
  if(e->format == 1)
    tile_state_fastq_kernel<<<1, 1024, 0, stream>>>(n_tiles, e->cntA.as<uint32_t>(), e->cntB.as<uint32_t>(), e->carry[e->carry_cur].as<Carry>(), e->tstate.as<uint8_t>());
  else
    tile_state_kernel<<<1, 1024, 0, stream>>>(dev, n_tiles, tile, e->nlA.as<long long>(), e->nlB.as<long long>(),
                                              e->carry[e->carry_cur].as<Carry>(), e->tstate.as<uint8_t>(), eff_min_qual(e));
  JF_LAUNCHED();
  CountArgs a;
  memset(&a, 0, sizeof(a));
  a.in = dev; a.n = n; a.n_look = n_look; a.n_tiles = n_tiles;
  a.tile_state = e->tstate.as<uint8_t>();
  a.carry_in = e->carry[e->carry_cur].as<Carry>();
  a.carry_out = e->carry[e->carry_cur ^ 1].as<Carry>();
  a.lut = e->tab.hash_fast ? e->tab.lut11.as<uint64_t>() : e->tab.lut.as<uint64_t>();
  a.hash_fast = e->tab.hash_fast ? 1 : 0; a.n_prow = e->tab.n_prow;
  a.lut_bytes = e->tab.hash_fast ? 4 * 2048 * 4 : e->nbytes * 256 * 8;
  for(unsigned i = 0; i < 8; ++i) a.prow[i] = e->tab.prow[i];
  a.min_qual = eff_min_qual(e); a.n_back = n_back;
  a.k = e->k; a.canonical = e->p.canonical; a.nbytes = e->nbytes; a.mode = (uint32_t)(shard_send ? 2 : mode); a.format = (uint32_t)e->format;
  a.T = table_dev(e, e->tab);
  a.bloom = bloom_dev(e);
  const size_t bloom_smem = a.bloom.mode ? (size_t)e->nbytes * 256 * 8 * 2 : 0;
  if(bc_build) { a.lut = nullptr; a.lut_bytes = 0; a.hash_fast = 0; }
  a.route_keys = route_keys; a.route_counts = route_counts; a.route_cap = route_cap; a.shard_bits = e->shard_bits;
  PartDev pd = shard_send ? shard_send_dev(e, (int)route_cap) : part_dev(e);
  auto launch = [&](auto kern, int nth, size_t smem, bool one_per_sm) -> int {
    cudaError_t c = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if(c != cudaSuccess) return fail(e, JFGPU_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(c));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // persistent CTAs: exactly as many as are resident at once (a multiple of the SM count)
    int per_sm = 1;
    if(!one_per_sm && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nth, smem) != cudaSuccess || per_sm < 1)) { cudaGetLastError(); per_sm = 1; }
    // (sharded counting leaves a few SMs to the collective that runs beside K1)
    const int sms = shard_send && e->n_sm > 32 ? e->n_sm - (int)SHARD_RESERVED_SMS : e->n_sm;
    const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)sms * per_sm);
    if(e->kev_used + 2 > e->kev.size()) { cudaEvent_t a0, a1; cudaEventCreate(&a0); cudaEventCreate(&a1); e->kev.push_back(a0); e->kev.push_back(a1); }
    cudaEventRecord(e->kev[e->kev_used], stream);
    kern<<<grid, nth, smem, stream>>>(a, pd);
    cudaEventRecord(e->kev[e->kev_used + 1], stream);
    e->kev_used += 2;
    return JFGPU_OK;
  };
  rc = dispatch(e, e->kw, e->tab.slot_bits, [&](auto KW, auto SB) -> int {
    constexpr int kw = decltype(KW)::value, sb = decltype(SB)::value;
    if(shard_send) {
      if(kw == 1 && e->tab.n_prow <= 2) return launch(extract_kernel<1, sb, 2, 1024, true, 2>, 1024, count_smem_bytes<1024>(a.lut_bytes, 0, 0, true), true);
      if(kw == 1) return launch(extract_kernel<1, sb, 2, 1024, true, 6>, 1024, count_smem_bytes<1024>(a.lut_bytes, 0, 0, true), true);
      return fail(e, JFGPU_ERR_STATE, "internal: record exchange with a two-word key");
    }
    if(part) {
      // the all-32-bit tail: 11-bit-table hash with at most two parity rows, 4-byte records, one shard, region index and
      // record fields inside 32 bits
      const bool fast = kw == 1 && e->tab.hash_fast && e->tab.n_prow <= 6 && ps.rec_bytes == 4 && e->shard_bits == 0 &&
                        ps.region_bits >= 2 && ps.region_bits < 32 && e->tab.lsize <= 38 && e->tab.lsize >= ps.region_bits &&
                        ps.P <= RING_P && !a.bloom.mode;
      if(kw == 1 && fast && e->tab.n_prow <= 2) return launch(extract_kernel<1, sb, 2, 1024, true, 2>, 1024, count_smem_bytes<1024>(a.lut_bytes, ps.stage_bytes, 0, true), true);
      if(kw == 1 && fast) return launch(extract_kernel<1, sb, 2, 1024, true, 6>, 1024, count_smem_bytes<1024>(a.lut_bytes, ps.stage_bytes, 0, true), true);
      return launch(extract_kernel<kw, sb, 2, 1024, false>, 1024, count_smem_bytes<1024>(a.lut_bytes, ps.stage_bytes, bloom_smem), true);
    }
    if(mode == 1) return launch(extract_kernel<kw, sb, 1, 512, false>, 512, count_smem_bytes<512>(a.lut_bytes, 0, bloom_smem), false);
    return launch(extract_kernel<kw, sb, 0, 512, false>, 512, count_smem_bytes<512>(a.lut_bytes, 0, bloom_smem), false);
  });
  if(rc) return rc;
  JF_LAUNCHED();
  CUDA_OK(e, cudaGetLastError());
  e->carry_cur ^= 1;
  return JFGPU_OK;
}

// ---- Bloom filter / counter ------------------------------------------------------------
// bloom_base::opt_m / opt_k (bloom_common.hpp:62-67)
int bloom_setup(jfgpu_engine* e, uint32_t mode, uint64_t n, double fp) {
  BloomState& b = e->bloom;
  if(fp <= 0.0) fp = mode == BLOOM_COUNT ? 0.001 : 0.01;          // bc_main_cmdline.yaggo / count_main_cmdline.yaggo defaults
  if(fp >= 1.0) return fail(e, JFGPU_ERR_ARG, "false positive rate must be in (0, 1)");
  const double LOG2 = 0.6931471805599453, LOG2_SQ = 0.4804530139182014;
  b.m = n * (uint64_t)lrint(-log(fp) / LOG2_SQ);
  b.k = (uint32_t)lrint(-log(fp) / LOG2);
  if(b.m == 0 || b.k == 0) return fail(e, JFGPU_ERR_ARG, "empty Bloom filter (size and false positive rate give no bits)");
  b.mode = mode;
  b.inv = b.m == 1 ? ~(uint64_t)0 : (uint64_t)(((unsigned __int128)1 << 64) / b.m);
  b.n_words = mode == BLOOM_COUNT ? (b.m + 15) / 16 : (b.m + 31) / 32;
  if(b.bits.alloc((size_t)b.n_words * 4 + 16) != cudaSuccess) { cudaGetLastError(); return fail(e, JFGPU_ERR_NOMEM, "Failed to allocate the Bloom filter in device memory"); }
  CUDA_OK(e, cudaMemsetAsync(b.bits.p, 0, b.bits.bytes, e->cs));
  if(mode == BLOOM_FILTER) {
    CUDA_OK(e, b.locks.alloc((size_t)4 << 20));
    CUDA_OK(e, cudaMemsetAsync(b.locks.p, 0, b.locks.bytes, e->cs));
  }
  b.drawn = false;
  return JFGPU_OK;
}
int bloom_upload_matrices(jfgpu_engine* e) {
  BloomState& b = e->bloom;
  std::vector<uint64_t> l1 = build_lut(b.M1, e->nbytes), l2 = build_lut(b.M2, e->nbytes);
  CUDA_OK(e, b.lut1.alloc(l1.size() * 8));
  CUDA_OK(e, b.lut2.alloc(l2.size() * 8));
  CUDA_OK(e, cudaMemcpyAsync(b.lut1.p, l1.data(), l1.size() * 8, cudaMemcpyHostToDevice, e->cs));
  CUDA_OK(e, cudaMemcpyAsync(b.lut2.p, l2.data(), l2.size() * 8, cudaMemcpyHostToDevice, e->cs));
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  b.cols1.assign(b.M1.c(), 0); b.cols2.assign(b.M2.c(), 0);
  for(unsigned i = 0; i < b.M1.c(); ++i) { b.cols1[i] = b.M1[i]; b.cols2[i] = b.M2[i]; }
  b.drawn = true;
  return JFGPU_OK;
}
// hash_pair<mer_dna>() (mer_dna_bloom_counter.hpp:22-27): two 64 x 2k matrices, randomize() only, the next draws of the stream
int bloom_draw(jfgpu_engine* e) {
  BloomState& b = e->bloom;
  if(b.drawn || b.mode == BLOOM_NONE) return JFGPU_OK;
  b.M1 = jfb::gf2_matrix(64, 2 * e->k); b.M1.randomize(e->rng);
  b.M2 = jfb::gf2_matrix(64, 2 * e->k); b.M2.randomize(e->rng);
  return bloom_upload_matrices(e);
}
BloomDev bloom_dev(const jfgpu_engine* e) {
  BloomDev d;
  memset(&d, 0, sizeof(d));
  const BloomState& b = e->bloom;
  // the PRIME pass of --if is not filtered (count_main.cc:288-295 builds that counter without a filter)
  if(b.mode == BLOOM_NONE || !b.drawn || (e->op == JFGPU_OP_PRIME && b.mode != BLOOM_COUNT)) return d;
  d.mode = b.mode; d.k = b.k; d.m = b.m; d.inv = b.inv;
  d.bits = b.bits.as<uint32_t>(); d.locks = b.locks.as<uint32_t>(); d.lock_mask = (uint32_t)(b.locks.bytes / 4 - 1);
  d.lut1 = b.lut1.as<uint64_t>(); d.lut2 = b.lut2.as<uint64_t>();
  return d;
}

int reset_carry(jfgpu_engine* e, cudaStream_t stream) {
  Carry c;
  c.state = e->format == 1 ? (0u | 4u) : (uint32_t)ST_L;    // FASTQ: header line expected, at a line start
  c.pad = 0;
  memset(c.sym, SYM_BREAK, sizeof(c.sym));
  // (stream ordered; source is copied synchronously into the driver's staging for pageable memory)
  CUDA_OK(e, cudaMemcpyAsync(e->carry[e->carry_cur].p, &c, sizeof(c), cudaMemcpyHostToDevice, stream));
  CUDA_OK(e, cudaStreamSynchronize(stream));
  return JFGPU_OK;
}

int read_stats(jfgpu_engine* e) {
  CUDA_OK(e, cudaMemcpyAsync(e->h_stats, e->stats.p, STAT_N * 8, cudaMemcpyDeviceToHost, e->cs));
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  return JFGPU_OK;
}

int insert_keys_into(jfgpu_engine* e, Table& t, const uint64_t* keys, const uint64_t* counts, uint64_t n, cudaStream_t stream) {
  if(n == 0) return JFGPU_OK;
  TableDev T = table_dev(e, t);
  const size_t smem = (size_t)e->nbytes * 256 * 8;
  const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)e->n_sm * 8);
  int rc = dispatch(e, e->kw, t.slot_bits, [&](auto KW, auto SB) -> int {
    auto kern = insert_keys_kernel<decltype(KW)::value, decltype(SB)::value>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 256, smem, stream>>>(T, t.lut.as<uint64_t>(), e->nbytes, keys, counts, n);
    return JFGPU_OK;
  });
  if(rc) return rc;
  JF_LAUNCHED();
  CUDA_OK(e, cudaGetLastError());
  return JFGPU_OK;
}

struct SegScratch {
  DevBuf keys, counts, sort_lo, n_out;
  uint64_t cap = 0;
  void free_all() { keys.free(); counts.free(); sort_lo.free(); n_out.free(); cap = 0; }
};

int seg_alloc(jfgpu_engine* e, SegScratch& s, uint64_t cap) {
  s.cap = cap;
  CUDA_OK(e, s.keys.alloc(cap * 8 * e->kw));
  CUDA_OK(e, s.counts.alloc(cap * 8));
  CUDA_OK(e, s.n_out.alloc(8));
  CUDA_OK(e, s.sort_lo.alloc(cap * 8));
  return JFGPU_OK;
}

// Collect the records of local original positions [lo, hi) of table t; returns their number.
int collect_segment(jfgpu_engine* e, Table& t, SegScratch& s, uint64_t lo, uint64_t hi, uint64_t lower, uint64_t upper, uint64_t* n_rec) {
  CollectArgs a;
  memset(&a, 0, sizeof(a));
  a.T = table_dev(e, t);
  a.inv_lut = t.inv_lut.as<uint64_t>();
  a.nbytes = e->nbytes;
  a.seg_lo = lo; a.seg_hi = hi;
  a.scan_hi = std::min<uint64_t>(hi + t.margin, t.local_size + t.margin);
  a.lower = lower; a.upper = upper;
  a.hb = t.hb;
  a.out_keys = s.keys.as<uint64_t>(); a.out_counts = s.counts.as<uint64_t>();
  a.out_sort_lo = s.sort_lo.as<uint64_t>();
  a.out_sort_hi = nullptr;
  a.out_n = s.n_out.as<unsigned long long>();
  a.out_cap = s.cap;
  CUDA_OK(e, cudaMemsetAsync(s.n_out.p, 0, 8, e->cs));
  const size_t smem = (size_t)e->nbytes * 256 * 8;
  const uint64_t span = a.scan_hi - a.seg_lo;
  const int grid = (int)std::min<uint64_t>((span + 255) / 256, (uint64_t)e->n_sm * 16);
  int rc = dispatch(e, e->kw, t.slot_bits, [&](auto KW, auto SB) -> int {
    auto kern = collect_kernel<decltype(KW)::value, decltype(SB)::value>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 256, smem, e->cs>>>(a);
    return JFGPU_OK;
  });
  if(rc) return rc;
  JF_LAUNCHED();
  unsigned long long n = 0;
  CUDA_OK(e, cudaMemcpyAsync(&n, s.n_out.p, 8, cudaMemcpyDeviceToHost, e->cs));
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  if(n > s.cap) return fail(e, JFGPU_ERR_STATE, "internal: segment overflow in collect");
  *n_rec = n;
  return JFGPU_OK;
}

uint64_t pick_segment(const Table& t) {
  uint64_t seg = std::min<uint64_t>(t.local_size, (uint64_t)1 << 24);
  return seg;
}

// Move every (key, count) of the current table, plus `n_failed` entries of failure list
// `old_fail`, into a fresh table of 2^nl global slots hashed with M.
int rebuild_table_impl(jfgpu_engine* e, unsigned nl, const jfb::gf2_matrix& M, int old_fail, uint64_t n_failed);
int rebuild_table(jfgpu_engine* e, unsigned nl, const jfb::gf2_matrix& M, int old_fail, uint64_t n_failed) {
  // moving (key, count) pairs is plain addition whatever operation the counter is in
  const uint32_t op = e->op;
  e->op = 0;
  const int rc = rebuild_table_impl(e, nl, M, old_fail, n_failed);
  e->op = op;
  return rc;
}
int rebuild_table_impl(jfgpu_engine* e, unsigned nl, const jfb::gf2_matrix& M, int old_fail, uint64_t n_failed) {
  Table nt;
  int rc = table_setup(e, nt, nl, M, e->tab.max_reprobe);
  if(rc) { nt.release(); return rc == JFGPU_ERR_NOMEM ? fail(e, JFGPU_ERR_FULL, "Hash full (" + e->err + ")") : rc; }
  // distinct / reprobes statistics restart for the new table; STAT_INSERTED counts k-mer
  // occurrences and must not change
  unsigned long long inserted_before = 0;
  CUDA_OK(e, cudaMemcpyAsync(&inserted_before, e->stats.as<unsigned long long>() + STAT_INSERTED, 8, cudaMemcpyDeviceToHost, e->cs));
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  CUDA_OK(e, cudaMemsetAsync(e->stats.as<unsigned long long>() + STAT_DISTINCT, 0, 8, e->cs));
  CUDA_OK(e, cudaMemsetAsync(e->stats.as<unsigned long long>() + STAT_REPROBES, 0, 8, e->cs));
  SegScratch s;
  const uint64_t seg = pick_segment(e->tab);
  rc = seg_alloc(e, s, seg + e->tab.margin + 8);
  for(uint64_t lo = 0; lo < e->tab.local_size && !rc; lo += seg) {
    uint64_t n = 0;
    rc = collect_segment(e, e->tab, s, lo, std::min(lo + seg, e->tab.local_size), 0, ~0ull, &n);
    if(!rc) rc = insert_keys_into(e, nt, s.keys.as<uint64_t>(), s.counts.as<uint64_t>(), n, e->cs);
  }
  cudaStreamSynchronize(e->cs);
  s.free_all();
  if(rc) { nt.release(); return rc; }
  // the moved entries are not new k-mer occurrences; the failed ones (below) are
  CUDA_OK(e, cudaMemcpyAsync(e->stats.as<unsigned long long>() + STAT_INSERTED, &inserted_before, 8, cudaMemcpyHostToDevice, e->cs));
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  if(n_failed) {
    rc = insert_keys_into(e, nt, e->fail_keys[old_fail].as<uint64_t>(), e->fail_counts[old_fail].as<uint64_t>(), n_failed, e->cs);
    cudaStreamSynchronize(e->cs);
    if(rc) { nt.release(); return rc; }
  }
  // (each table owns its counter-carry side table: the old one dies with the old slots)
  e->tab.release();
  e->tab = nt;
  if(!e->part.pending) { part_configure(e); if(!e->part.P) part_release(e); }
  return JFGPU_OK;
}

// hash_counter::handle_full_ary without doubling (hash_counter.hpp:187-192): the caller's hook writes the resident table out,
// the table is zeroed, the keys that found no slot go into the empty table, counting continues with the same geometry.
int spill_table(jfgpu_engine* e, uint64_t n_failed) {
  e->in_spill = true;
  const int hrc = e->spill_fn(e->spill_ctx, e);
  e->in_spill = false;
  if(hrc) return fail(e, JFGPU_ERR_SINK, "the spill hook failed (--disk: writing an intermediate file)");
  CUDA_OK(e, cudaMemsetAsync(e->tab.slots.p, 0, e->tab.bytes(), e->cs));
  CUDA_OK(e, cudaMemsetAsync(e->tab.ovf_keys.p, 0, e->tab.ovf_size * 8, e->cs));
  CUDA_OK(e, cudaMemsetAsync(e->tab.ovf_vals.p, 0, e->tab.ovf_size * 8, e->cs));
  unsigned long long* st = e->stats.as<unsigned long long>();
  CUDA_OK(e, cudaMemsetAsync(st + STAT_DISTINCT, 0, 8, e->cs));       // (statistics of the table restart; STAT_INSERTED counts occurrences and goes on)
  CUDA_OK(e, cudaMemsetAsync(st + STAT_REPROBES, 0, 8, e->cs));
  CUDA_OK(e, cudaMemsetAsync(st + STAT_OVERFLOWED, 0, 8, e->cs));
  CUDA_OK(e, cudaMemsetAsync(st + STAT_FAILED, 0, 8, e->cs));
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  const int old_fail = e->fail_cur;
  e->fail_cur ^= 1;                                   // (failures of the re-insertion -- there should be none -- are kept apart)
  if(!e->fail_keys[e->fail_cur].p) {
    CUDA_OK(e, e->fail_keys[e->fail_cur].alloc(e->fail_cap * 8 * e->kw));
    CUDA_OK(e, e->fail_counts[e->fail_cur].alloc(e->fail_cap * 8));
  }
  const uint32_t op = e->op; e->op = 0;
  int rc = insert_keys_into(e, e->tab, e->fail_keys[old_fail].as<uint64_t>(), e->fail_counts[old_fail].as<uint64_t>(), n_failed, e->cs);
  e->op = op;
  if(rc) return rc;
  CUDA_OK(e, cudaStreamSynchronize(e->cs));       // (the keys that had found no slot are counted as inserted now, once)
  e->spills++;
  return JFGPU_OK;
}

// hash_counter::double_size (hash_counter.hpp:200-238): allocate a table twice as large with a
// freshly drawn matrix, re-insert every (key, count) of the old one, then the keys that failed.
int regrow(jfgpu_engine* e) {
  for(;;) {
    int rc = read_stats(e);
    if(rc) return rc;
    if(e->h_stats[STAT_OVF_FULL]) return fail(e, JFGPU_ERR_FULL, "counter overflow side table is full");
    const uint64_t n_failed = e->h_stats[STAT_FAILED];
    if(n_failed == 0) return JFGPU_OK;
    if(e->h_stats[STAT_FAIL_DROPPED]) return fail(e, JFGPU_ERR_FULL, "Hash full (too many keys failed before the table could be doubled)");
    const unsigned kbits = 2 * e->k;
    if(!e->p.allow_regrow || e->shard_bits || (kbits < 64 && e->tab.size >= ((uint64_t)1 << kbits))) {
      if(e->spill_fn && !e->shard_bits) { rc = spill_table(e, n_failed); if(rc) return rc; continue; }
      return fail(e, JFGPU_ERR_FULL, "Hash full");
    }
    const unsigned nl = e->tab.lsize + 1;
    jfb::gf2_matrix M = draw_matrix(e, (uint64_t)1 << nl, nl);
    // switch the failure list so that failures of the re-insertion are kept apart
    const int old_fail = e->fail_cur;
    e->fail_cur ^= 1;
    if(!e->fail_keys[e->fail_cur].p) {
      CUDA_OK(e, e->fail_keys[e->fail_cur].alloc(e->fail_cap * 8 * e->kw));
      CUDA_OK(e, e->fail_counts[e->fail_cur].alloc(e->fail_cap * 8));
    }
    CUDA_OK(e, cudaMemsetAsync(e->stats.as<unsigned long long>() + STAT_FAILED, 0, 8, e->cs));
    rc = rebuild_table(e, nl, M, old_fail, n_failed);
    if(rc == JFGPU_ERR_FULL && e->spill_fn) {          // no memory for the doubled table: dump and zero this one instead
      e->fail_cur = old_fail;
      CUDA_OK(e, cudaMemcpyAsync(e->stats.as<unsigned long long>() + STAT_FAILED, &n_failed, 8, cudaMemcpyHostToDevice, e->cs));
      CUDA_OK(e, cudaStreamSynchronize(e->cs));
      rc = spill_table(e, n_failed);
      if(rc) return rc;
      continue;
    }
    if(rc) return rc;
    e->regrows++;
  }
}

// Direct-indexing regime (table as large as the key space, no reprobing).  The reference
// cannot chain "large" continuation entries there, so a counter that outgrows val_len bits
// makes hash_counter::double_size build a new array with val_len + 1 -- and, the size being
// 4^k, with the IDENTITY matrix (hash_counter.hpp:205-212, large_hash_array.hpp:992-1002).
// The observable result is: val_len = bits of the largest count, identity hash.
int direct_index_fixup(jfgpu_engine* e) {
  const unsigned kbits = 2 * e->k;
  if(e->tab.lsize < kbits || e->shard_bits) return JFGPU_OK;
  CUDA_OK(e, cudaMemsetAsync(e->stats.as<unsigned long long>() + STAT_MAXCOUNT, 0, 8, e->cs));
  TableDev T = table_dev(e, e->tab);
  const uint64_t ns = e->tab.local_size + e->tab.margin;
  const int grid = (int)std::min<uint64_t>((ns + 255) / 256, (uint64_t)e->n_sm * 16);
  switch(e->tab.slot_bits) {
  case 32:  max_count_kernel<32><<<grid, 256, 0, e->cs>>>(T, ns); break;
  case 64:  max_count_kernel<64><<<grid, 256, 0, e->cs>>>(T, ns); break;
  default:  max_count_kernel<128><<<grid, 256, 0, e->cs>>>(T, ns); break;
  }
  JF_LAUNCHED();
  int rc = read_stats(e);
  if(rc) return rc;
  const uint64_t maxc = e->h_stats[STAT_MAXCOUNT];
  if(e->eff_val_len < 64 && (maxc >> e->eff_val_len) != 0) {
    e->eff_val_len = bitsize(maxc);
    if(!e->tab.M.is_identity()) {
      rc = rebuild_table(e, e->tab.lsize, jfb::gf2_matrix::identity(kbits), 0, 0);
      if(rc) return rc;
    }
  }
  return JFGPU_OK;
}

// fold the per-launch event pairs into kernel_ms (the stream must be idle)
void resolve_kernel_events(jfgpu_engine* e) {
  for(size_t i = 0; i + 1 < e->kev_used; i += 2) {
    float ms = 0;
    if(cudaEventElapsedTime(&ms, e->kev[i], e->kev[i + 1]) == cudaSuccess) { e->kernel_ms += ms; e->kernel_launches++; }
    else cudaGetLastError();
  }
  e->kev_used = 0;
}

int check_after_batches(jfgpu_engine* e) {
  int rc = read_stats(e);
  if(rc) return rc;
  if(e->h_stats[STAT_OVF_FULL]) return fail(e, JFGPU_ERR_FULL, "counter overflow side table is full");
  if(e->h_stats[STAT_FAILED]) {
    // records staged against the CURRENT geometry must reach the table before it is rebuilt: the drain copes
    // with a doubling in its middle (old inverse tables + rehash kernel), a plain regrow() would not
    if(e->part.pending) {
      rc = part_drain(e, e->cs);
      if(rc) return rc;
      rc = read_stats(e);
      if(rc) return rc;
      if(!e->h_stats[STAT_FAILED]) return JFGPU_OK;
    }
    return regrow(e);
  }
  return JFGPU_OK;
}

}  // namespace

// =======================================================================================
// C ABI
// =======================================================================================
extern "C" {

const char* jfgpu_version(void) { return "jellyfish-b200 0.1 (sm_100a)"; }
uint64_t jfgpu_kernel_launches(void) { return g_launches.load(); }
const char* jfgpu_last_error(jfgpu_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void* jfgpu_host_alloc(size_t bytes) { void* p = nullptr; if(cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void jfgpu_host_free(void* p) { if(p) cudaFreeHost(p); }
int jfgpu_memcpy_h2d(void* dev_dst, const void* host_src, size_t bytes, void* stream) {
  if(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream) != cudaSuccess) { cudaGetLastError(); return JFGPU_ERR_CUDA; }
  return JFGPU_OK;
}

int jfgpu_reference_matrix(uint32_t r, uint32_t c, uint32_t skip, uint64_t* cols) {
  if(r == 0 || r > 64 || c == 0 || !cols) return JFGPU_ERR_ARG;
  jfb::glibc_random rng;
  jfb::gf2_matrix res;
  for(uint32_t i = 0; i <= skip; ++i) { jfb::gf2_matrix m(r, c); res = m.randomize_pseudo_inverse(rng); }
  for(uint32_t i = 0; i < c; ++i) cols[i] = res[i];
  return JFGPU_OK;
}

int jfgpu_create(const jfgpu_params* params, jfgpu_handle* out) {
  if(!params || !out) return fail(nullptr, JFGPU_ERR_ARG, "null argument");
  if(params->struct_size != sizeof(jfgpu_params)) return fail(nullptr, JFGPU_ERR_ARG, "jfgpu_params size mismatch");
  if(params->k < 1 || params->k > 64) return fail(nullptr, JFGPU_ERR_ARG, "mer length must be in [1, 64]");
  if(params->size == 0) return fail(nullptr, JFGPU_ERR_ARG, "size must be positive");
  uint32_t ns = params->n_shards ? params->n_shards : 1;
  if(ns & (ns - 1)) return fail(nullptr, JFGPU_ERR_ARG, "n_shards must be a power of two");
  if(params->shard_index >= ns) return fail(nullptr, JFGPU_ERR_ARG, "shard_index out of range");
  if(params->max_reprobe > 255) return fail(nullptr, JFGPU_ERR_ARG, "max_reprobe must be <= 255");

  int ndev = 0;
  if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, JFGPU_ERR_CUDA, "no CUDA device: the jellyfish-b200 engine has no CPU fallback");
  }
  if(params->device < 0 || params->device >= ndev) return fail(nullptr, JFGPU_ERR_ARG, "invalid device ordinal");

  jfgpu_engine* e = new jfgpu_engine;
  e->p = *params;
  e->p.n_shards = ns;
  e->device = params->device;
  e->k = params->k;
  e->kw = params->k > 32 ? 2 : 1;
  e->nbytes = (2 * params->k + 7) / 8;
  e->eff_val_len = params->counter_len;
  e->shard_bits = ceil_log2(ns);
  auto bail = [&](int code) { g_create_error = e->err; jfgpu_destroy(e); return code; };

  cudaError_t c;
  if((c = cudaSetDevice(e->device)) != cudaSuccess) { e->err = cudaGetErrorString(c); return bail(JFGPU_ERR_CUDA); }
  cudaDeviceProp prop;
  if((c = cudaGetDeviceProperties(&prop, e->device)) != cudaSuccess) { e->err = cudaGetErrorString(c); return bail(JFGPU_ERR_CUDA); }
  if(prop.major < 10) { e->err = "this engine is built for sm_100a (Blackwell) only"; return bail(JFGPU_ERR_CUDA); }
  e->n_sm = prop.multiProcessorCount;
  if(cudaStreamCreateWithFlags(&e->cs, cudaStreamNonBlocking) != cudaSuccess ||
     cudaStreamCreateWithFlags(&e->hs, cudaStreamNonBlocking) != cudaSuccess) { e->err = "stream creation failed"; return bail(JFGPU_ERR_CUDA); }
  cudaEventCreate(&e->ev_t0); cudaEventCreate(&e->ev_t1);
  cudaHostAlloc((void**)&e->h_watch, 16, cudaHostAllocDefault);
  for(int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&e->ev_watch[i], cudaEventDisableTiming);
  for(int i = 0; i < 2; ++i) { cudaEventCreateWithFlags(&e->ev_copied[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming); }

  if(params->bloom_counter) {
    // `jellyfish bc`: no hash table at all; the two hash matrices are the FIRST draws of the random stream (bc_main.cc:103-106)
    bool ok0 = e->stats.alloc(STAT_N * 8) == cudaSuccess && e->carry[0].alloc(sizeof(Carry)) == cudaSuccess && e->carry[1].alloc(sizeof(Carry)) == cudaSuccess &&
               cudaHostAlloc((void**)&e->h_stats, STAT_N * 8, cudaHostAllocDefault) == cudaSuccess;
    if(!ok0) { cudaGetLastError(); e->err = "device allocation failed"; return bail(JFGPU_ERR_NOMEM); }
    cudaMemsetAsync(e->stats.p, 0, STAT_N * 8, e->cs);
    memset(e->h_stats, 0, STAT_N * 8);
    e->batch_bytes = params->max_batch_bytes ? (size_t)((params->max_batch_bytes + 15) & ~(uint64_t)15) : ((size_t)64 << 20);
    e->tab.slot_bits = 64; e->tab.lsize = 0; e->tab.size = 0;
    int rc0 = bloom_setup(e, BLOOM_COUNT, params->bf_size, params->bf_fp);
    if(!rc0) rc0 = bloom_draw(e);
    if(!rc0) rc0 = reset_carry(e, e->cs);
    if(rc0) return bail(rc0);
    *out = e;
    return JFGPU_OK;
  }

  // table geometry: size rounded up to a power of two, clipped to 4^k (large_hash_array.hpp:992-1002,150-153)
  const unsigned kbits = 2 * e->k;
  uint64_t req = params->size;
  if(kbits < 64 && req > ((uint64_t)1 << kbits)) req = (uint64_t)1 << kbits;
  const unsigned lsize = ceil_log2(req);
  if(lsize < e->shard_bits) { e->err = "table smaller than the number of shards"; return bail(JFGPU_ERR_ARG); }
  for(uint32_t i = 0; i < params->matrix_skip; ++i) (void)draw_matrix(e, params->size, lsize);
  jfb::gf2_matrix M = draw_matrix(e, params->size, lsize);

  // side structures
  e->batch_bytes = params->max_batch_bytes ? (size_t)((params->max_batch_bytes + 15) & ~(uint64_t)15) : ((size_t)64 << 20);
  e->fail_cap = 2 * (uint64_t)e->batch_bytes;      // (two groups of failed keys: the failure counter is read one group late)
  bool ok = e->stats.alloc(STAT_N * 8) == cudaSuccess && e->carry[0].alloc(sizeof(Carry)) == cudaSuccess &&
            e->carry[1].alloc(sizeof(Carry)) == cudaSuccess &&
            e->fail_keys[0].alloc(e->fail_cap * 8 * e->kw) == cudaSuccess && e->fail_counts[0].alloc(e->fail_cap * 8) == cudaSuccess &&
            cudaHostAlloc((void**)&e->h_stats, STAT_N * 8, cudaHostAllocDefault) == cudaSuccess;
  if(!ok) { cudaGetLastError(); e->err = "device allocation failed"; return bail(JFGPU_ERR_NOMEM); }
  cudaMemsetAsync(e->stats.p, 0, STAT_N * 8, e->cs);
  memset(e->h_stats, 0, STAT_N * 8);
  int rc = table_setup(e, e->tab, lsize, M, params->max_reprobe);
  if(rc) return bail(rc);
  part_configure(e);
  // count --bf-size: mer_dna_bloom_filter(bf_fp, bf_size) in front of the table (count_main.cc:317-321); its matrices are
  // drawn when the first unprimed text arrives
  if(params->bf_size) { rc = bloom_setup(e, BLOOM_FILTER, params->bf_size, params->bf_fp); if(rc) return bail(rc); }
  rc = reset_carry(e, e->cs);
  if(rc) return bail(rc);
  *out = e;
  return JFGPU_OK;
}

void jfgpu_destroy(jfgpu_handle e) {
  if(!e) return;
  cudaSetDevice(e->device);
  if(e->cs) cudaStreamSynchronize(e->cs);
  if(e->hs) cudaStreamSynchronize(e->hs);
  e->tab.release();
  part_release(e);
  e->bloom.release();
  e->sh.pool_next[0].free(); e->sh.pool_next[1].free(); e->sh.cta_chunk.free(); e->sh.cta_fill.free();
  if(e->sh.h_counts) cudaFreeHost(e->sh.h_counts);
  e->stats.free();
  for(int i = 0; i < 2; ++i) {
    e->carry[i].free(); e->fail_keys[i].free(); e->fail_counts[i].free(); e->stage[i].free();
    if(e->ev_copied[i]) cudaEventDestroy(e->ev_copied[i]);
    if(e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
  }
  e->nlA.free(); e->nlB.free(); e->cntA.free(); e->cntB.free(); e->tstate.free();
  if(e->h_stats) cudaFreeHost(e->h_stats);
  if(e->h_watch) cudaFreeHost(e->h_watch);
  for(int i = 0; i < 2; ++i) if(e->ev_watch[i]) cudaEventDestroy(e->ev_watch[i]);
  if(e->ev_t0) cudaEventDestroy(e->ev_t0);
  if(e->ev_t1) cudaEventDestroy(e->ev_t1);
  if(e->ev_d0) cudaEventDestroy(e->ev_d0);
  if(e->ev_d1) cudaEventDestroy(e->ev_d1);
  for(cudaEvent_t ev : e->kev) cudaEventDestroy(ev);
  for(cudaEvent_t ev : e->wev) cudaEventDestroy(ev);
  if(e->cs) cudaStreamDestroy(e->cs);
  if(e->hs) cudaStreamDestroy(e->hs);
  delete e;
}

static int begin_feed(jfgpu_engine* e, uint32_t flags, int first_byte, cudaStream_t st) {
  if(flags & JFGPU_FILE_BEGIN) {
    // mer_overlap_sequence_parser.hpp:134-148: the first byte selects the format
    if(first_byte >= 0 && first_byte != '>' && first_byte != '@') return fail(e, JFGPU_ERR_FORMAT, "Unsupported format");
    e->format = first_byte == '@' ? 1 : 0;
    int rc = reset_carry(e, st);
    if(rc) return rc;
    e->in_file = true;
  }
  return JFGPU_OK;
}

static int end_feed(jfgpu_engine* e, uint32_t flags, cudaStream_t st) {
  if(flags & JFGPU_FILE_END) {
    // no k-mer spans two files: mer_overlap_sequence_parser.hpp:111
    e->format = 0;
    int rc = reset_carry(e, st);
    if(rc) return rc;
    e->in_file = false;
  }
  return JFGPU_OK;
}

int jfgpu_feed_device(jfgpu_handle e, const void* dev_bytes, size_t n, uint32_t flags, void* stream) {
  if(!e) return JFGPU_ERR_ARG;
  if(((uintptr_t)dev_bytes & 15) != 0) return fail(e, JFGPU_ERR_ARG, "device text must be 16-byte aligned");
  cudaSetDevice(e->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->cs;
  int first = -1;
  if((flags & JFGPU_FILE_BEGIN) && n) {
    unsigned char b = 0;
    CUDA_OK(e, cudaMemcpyAsync(&b, dev_bytes, 1, cudaMemcpyDeviceToHost, st));
    CUDA_OK(e, cudaStreamSynchronize(st));
    first = b;
  }
  int rc = begin_feed(e, flags, first, st);
  if(rc) return rc;
  const uint8_t* p = (const uint8_t*)dev_bytes;
  if(e->part.P) { rc = part_alloc(e); if(rc) return rc; }
  cudaEventRecord(e->ev_t0, st);
  for(size_t off = 0; off < n; ) {
    size_t len = std::min(e->part.P ? std::max<size_t>(e->batch_bytes, (size_t)512 << 20) : e->batch_bytes, n - off);
    len = part_cap_len(e, len);
    rc = run_batch(e, p + off, len, n - off, st, 0, nullptr, nullptr, 0, off);
    if(rc) return rc;
    off += len;
    if((e->p.allow_regrow || e->spill_fn) && !e->part.P && e->tab.slots.p) {          // the failure list only holds two batches
      if(st != e->cs) CUDA_OK(e, cudaStreamSynchronize(st));
      rc = check_after_batches(e);
      if(rc) return rc;
    }
  }
  cudaEventRecord(e->ev_t1, st);
  CUDA_OK(e, cudaStreamSynchronize(st));
  float ms = 0; cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1); e->count_ms += ms;
  resolve_kernel_events(e);
  e->bytes_fed += n;
  return end_feed(e, flags, st);
}

// length of the longest prefix of [p, p+len) that ends right behind the newline closing a 4-line record, given the number
// of lines (mod 4) in front of p; 0 when there is none.  *lines_out = lines (mod 4) at that point.
static size_t fastq_record_prefix(const char* p, size_t len, uint32_t lines_mod4, uint32_t* lines_out) {
  size_t best = 0; uint32_t lines = lines_mod4, best_lines = lines_mod4;
  const char* q = p; const char* end = p + len;
  while(q < end) {
    const char* nl = (const char*)memchr(q, '\n', (size_t)(end - q));
    if(!nl) break;
    lines = (lines + 1) & 3u;
    q = nl + 1;
    if(lines == 0) { best = (size_t)(q - p); best_lines = 0; }
  }
  *lines_out = best_lines;
  return best;
}

int jfgpu_feed(jfgpu_handle e, const char* bytes, size_t n, uint32_t flags) {
  if(!e) return JFGPU_ERR_ARG;
  cudaSetDevice(e->device);
  int rc = begin_feed(e, flags, n ? (unsigned char)bytes[0] : (e->q_tail.empty() ? -1 : (unsigned char)e->q_tail[0]), e->cs);
  if(rc) return rc;
  if(flags & JFGPU_FILE_BEGIN) { e->q_lines = 0; e->q_tail.clear(); }
  const bool qfastq = eff_min_qual(e) != 0 && e->format == 1;
  std::string joined;                         // (-Q on FASTQ: the incomplete record of the previous feed comes first)
  if(qfastq && !e->q_tail.empty()) {
    joined.swap(e->q_tail);
    joined.append(bytes, n);
    bytes = joined.data(); n = joined.size();
  }
  if(qfastq && !(flags & JFGPU_FILE_END)) {   // keep the incomplete last record for the next feed
    uint32_t l2 = 0;
    const size_t whole = fastq_record_prefix(bytes, n, e->q_lines, &l2);
    e->q_tail.assign(bytes + whole, n - whole);
    n = whole;
  }
  for(int i = 0; i < 2; ++i) if(!e->stage[i].p) CUDA_OK(e, e->stage[i].alloc(e->batch_bytes + 64));
  if(e->part.P) { rc = part_alloc(e); if(rc) return rc; }
  cudaEventRecord(e->ev_t0, e->cs);
  size_t off = 0;
  while(off < n) {
    size_t len = part_cap_len(e, std::min(e->batch_bytes, n - off));
    // never end a chunk on '\r' unless it is the end of the data: the device looks one byte ahead
    if(off + len < n) { size_t l2 = len; while(l2 > 1 && bytes[off + l2 - 1] == '\r') --l2; if(l2 > 1) len = l2; }
    if(qfastq && off + len < n) {              // ... and, under -Q, only on a record boundary
      uint32_t l2 = 0;
      const size_t whole = fastq_record_prefix(bytes + off, len, e->q_lines, &l2);
      if(whole == 0) return fail(e, JFGPU_ERR_FORMAT, "Invalid fastq file: a record is larger than the staging buffer (or has more than 4 lines)");
      len = whole; e->q_lines = l2;
    }
    const int s = e->stage_cur;
    // the previous batch that used this staging buffer must be done before it is overwritten
    CUDA_OK(e, cudaEventSynchronize(e->ev_done[s]));
    if((e->p.allow_regrow || e->spill_fn) && e->tab.slots.p) {
      // peek at the live failure counter without draining the compute stream
      CUDA_OK(e, cudaMemcpyAsync(e->h_stats + STAT_FAILED, e->stats.as<unsigned long long>() + STAT_FAILED, 8, cudaMemcpyDeviceToHost, e->hs));
      CUDA_OK(e, cudaStreamSynchronize(e->hs));
      if(e->h_stats[STAT_FAILED]) { rc = check_after_batches(e); if(rc) return rc; }
    }
    CUDA_OK(e, cudaMemcpyAsync(e->stage[s].p, bytes + off, len, cudaMemcpyHostToDevice, e->hs));
    CUDA_OK(e, cudaEventRecord(e->ev_copied[s], e->hs));
    CUDA_OK(e, cudaStreamWaitEvent(e->cs, e->ev_copied[s], 0));
    rc = run_batch(e, e->stage[s].as<uint8_t>(), len, len, e->cs, 0, nullptr, nullptr, 0);
    if(rc) return rc;
    CUDA_OK(e, cudaEventRecord(e->ev_done[s], e->cs));
    e->stage_cur ^= 1;
    off += len;
  }
  if(qfastq && !(flags & JFGPU_FILE_END)) e->q_lines = 0;       // (the feed was cut behind a complete record)
  cudaEventRecord(e->ev_t1, e->cs);
  e->bytes_fed += n;
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  { float ms = 0; cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1); e->count_ms += ms; }
  resolve_kernel_events(e);
  if(e->tab.slots.p) { rc = check_after_batches(e); if(rc) return rc; }
  return end_feed(e, flags, e->cs);
}

int jfgpu_extract_route(jfgpu_handle e, const void* dev_bytes, size_t n, uint32_t flags, void* dev_keys, uint64_t capacity,
                        uint64_t* dev_counts, void* stream) {
  if(!e) return JFGPU_ERR_ARG;
  if(!e->tab.slots.p || e->bloom.mode != BLOOM_NONE) return fail(e, JFGPU_ERR_STATE, "Bloom filters are not supported on the sharded path");
  if(((uintptr_t)dev_bytes & 15) != 0) return fail(e, JFGPU_ERR_ARG, "device text must be 16-byte aligned");
  cudaSetDevice(e->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->cs;
  int first = -1;
  if((flags & JFGPU_FILE_BEGIN) && n) {
    unsigned char b = 0;
    CUDA_OK(e, cudaMemcpyAsync(&b, dev_bytes, 1, cudaMemcpyDeviceToHost, st));
    CUDA_OK(e, cudaStreamSynchronize(st));
    first = b;
  }
  int rc = begin_feed(e, flags, first, st);
  if(rc) return rc;
  const uint8_t* p = (const uint8_t*)dev_bytes;
  for(size_t off = 0; off < n; ) {
    size_t len = std::min(e->batch_bytes, n - off);
    rc = run_batch(e, p + off, len, n - off, st, 1, (uint64_t*)dev_keys, (unsigned long long*)dev_counts, capacity, off);
    if(rc) return rc;
    off += len;
  }
  e->bytes_fed += n;
  if(stream && !(flags & (JFGPU_FILE_BEGIN | JFGPU_FILE_END))) return JFGPU_OK;   // stream-ordered: the caller synchronises; drops are reported by jfgpu_finish
  CUDA_OK(e, cudaStreamSynchronize(st));
  rc = end_feed(e, flags, st);
  if(rc) return rc;
  rc = read_stats(e);
  if(rc) return rc;
  if(e->h_stats[STAT_ROUTE_DROPPED]) return fail(e, JFGPU_ERR_FULL, "route bucket capacity exceeded");
  return JFGPU_OK;
}

int jfgpu_shard_setup(jfgpu_handle e, const jfgpu_shard_buffers* b) {
  if(!e || !b) return JFGPU_ERR_ARG;
  cudaSetDevice(e->device);
  const uint32_t G = e->p.n_shards;
  ShardState& sh = e->sh;
  sh.on = false;
  // geometry the record exchange covers: one key word, the 32-bit hash tail, 4-byte records of the GLOBAL regions, the
  // receiver's own partition in 4-byte records drained by the window kernels
  const Table& t = e->tab;
  if(G < 2 || G > 8 || !t.slots.p || e->kw != 1 || !t.hash_fast || t.n_prow > 6 || t.lsize > 38 || t.slot_bits != 32 || e->bloom.mode != BLOOM_NONE ||
     !e->part.P || e->part.rec_bytes != 4 || e->part.P > RING_P)
    return fail(e, JFGPU_ERR_ARG, "this table geometry is not covered by the record exchange (use the key exchange)");
  uint32_t P = RING_P;
  while(P > G && t.lsize < ceil_log2(P) + 14) P >>= 1;
  const uint32_t sbits = t.lsize - ceil_log2(P);
  if(sbits + t.hb > 32 || P < G || sbits < e->part.region_bits || (P / G) > e->part.P)
    return fail(e, JFGPU_ERR_ARG, "this table geometry is not covered by the record exchange (use the key exchange)");
  if(!b->send_pool && !b->recv_pool) return JFGPU_OK;          // geometry probe only
  if(!b->send_pool || !b->send_dir || !b->recv_pool || !b->recv_dir || b->send_arena_chunks < 2 * (uint64_t)e->n_sm * (P / G) || b->recv_seg_chunks < b->send_arena_chunks)
    return fail(e, JFGPU_ERR_ARG, "exchange buffers too small: an arena must hold the open chunks of every CTA twice over");
  sh.P = P; sh.sbits = sbits; sh.own_regions = P / G; sh.owner_shift = ceil_log2(P / G); sh.split_lg = sbits - e->part.region_bits;
  sh.arena_chunks = b->send_arena_chunks; sh.seg_chunks = b->recv_seg_chunks;
  sh.send_pool = (uint8_t*)b->send_pool; sh.send_dir = (uint2*)b->send_dir; sh.recv_pool = (uint8_t*)b->recv_pool; sh.recv_dir = (uint2*)b->recv_dir;
  bool ok = sh.pool_next[0].alloc(((size_t)G + 2) * 4) == cudaSuccess && sh.pool_next[1].alloc(((size_t)G + 2) * 4) == cudaSuccess &&
            sh.cta_chunk.alloc((size_t)e->n_sm * RING_P * 4) == cudaSuccess && sh.cta_fill.alloc((size_t)e->n_sm * RING_P * 4) == cudaSuccess &&
            (sh.h_counts || cudaHostAlloc((void**)&sh.h_counts, 16 * sizeof(unsigned int), cudaHostAllocDefault) == cudaSuccess);
  if(!ok) { cudaGetLastError(); return fail(e, JFGPU_ERR_NOMEM, "device allocation failed"); }
  for(int i = 0; i < 2; ++i) CUDA_OK(e, cudaMemsetAsync(sh.pool_next[i].p, 0, sh.pool_next[i].bytes, e->cs));
  CUDA_OK(e, cudaMemsetAsync(sh.cta_chunk.p, 0xFF, sh.cta_chunk.bytes, e->cs));
  CUDA_OK(e, cudaMemsetAsync(sh.cta_fill.p, 0, sh.cta_fill.bytes, e->cs));
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  sh.on = true;
  return JFGPU_OK;
}

uint64_t jfgpu_shard_round_bytes(jfgpu_handle e) {
  if(!e || !e->sh.on) return 0;
  // every byte gives at most one record; iid keys spread evenly over the shards, a third of an arena is kept as slack, and
  // every CTA parks one open chunk per region in the arena of its owner
  const ShardState& sh = e->sh;
  const uint64_t open = (uint64_t)e->n_sm * sh.own_regions;
  const uint64_t room = sh.arena_chunks > open ? sh.arena_chunks - open : 0;
  const uint64_t recs = room * (CHUNK_BYTES / 4 - 2 * RING) * 3 / 4;
  const uint64_t bytes = recs * e->p.n_shards;
  return bytes & ~(uint64_t)0xFFFFF;
}

int jfgpu_shard_extract(jfgpu_handle e, const void* dev_bytes, size_t n, uint32_t flags, uint32_t bank, void* stream) {
  if(!e) return JFGPU_ERR_ARG;
  if(!e->sh.on || bank > 1) return fail(e, JFGPU_ERR_STATE, "jfgpu_shard_setup has not been called");
  if(((uintptr_t)dev_bytes & 15) != 0) return fail(e, JFGPU_ERR_ARG, "device text must be 16-byte aligned");
  cudaSetDevice(e->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->cs;
  int first = -1;
  if((flags & JFGPU_FILE_BEGIN) && n) {
    unsigned char b = 0;
    CUDA_OK(e, cudaMemcpyAsync(&b, dev_bytes, 1, cudaMemcpyDeviceToHost, st));
    CUDA_OK(e, cudaStreamSynchronize(st));
    first = b;
  }
  int rc = begin_feed(e, flags, first, st);
  if(rc) return rc;
  const uint8_t* p = (const uint8_t*)dev_bytes;
  for(size_t off = 0; off < n; ) {
    const size_t len = std::min<size_t>((size_t)512 << 20, n - off);
    rc = run_batch(e, p + off, len, n - off, st, 3, nullptr, nullptr, bank, off);
    if(rc) return rc;
    off += len;
  }
  e->bytes_fed += n;
  if(flags & JFGPU_FILE_END) { CUDA_OK(e, cudaStreamSynchronize(st)); return end_feed(e, flags, st); }
  return JFGPU_OK;                       // stream-ordered
}

int jfgpu_shard_pack(jfgpu_handle e, uint32_t bank, uint64_t* counts, void* stream) {
  if(!e || !counts) return JFGPU_ERR_ARG;
  if(!e->sh.on || bank > 1) return fail(e, JFGPU_ERR_STATE, "jfgpu_shard_setup has not been called");
  cudaSetDevice(e->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->cs;
  const uint32_t G = e->p.n_shards;
  PartDev pd = shard_send_dev(e, (int)bank);
  close_chunks_kernel<<<e->n_sm * 4, 256, 0, st>>>(pd, (uint32_t)e->n_sm); JF_LAUNCHED();
  CUDA_OK(e, cudaMemcpyAsync(e->sh.h_counts, pd.pool_next, G * 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(e, cudaStreamSynchronize(st));
  for(uint32_t d = 0; d < G; ++d) {
    if(e->sh.h_counts[d] > e->sh.arena_chunks) return fail(e, JFGPU_ERR_FULL, "route bucket capacity exceeded (an arena of the send pool overflowed)");
    counts[d] = e->sh.h_counts[d];
  }
  CUDA_OK(e, cudaMemsetAsync(pd.pool_next, 0, ((size_t)G + 2) * 4, st));    // (the chunks stay where they are until the caller has sent them)
  return JFGPU_OK;
}

int jfgpu_shard_unpack(jfgpu_handle e, const uint64_t* counts, uint32_t self_bank, void* stream) {
  if(!e || !counts) return JFGPU_ERR_ARG;
  if(!e->sh.on) return fail(e, JFGPU_ERR_STATE, "jfgpu_shard_setup has not been called");
  cudaSetDevice(e->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->cs;
  const uint32_t G = e->p.n_shards;
  PartState& ps = e->part;
  int rc = part_alloc(e);
  if(rc) return rc;
  uint64_t total = 0;
  for(uint32_t s = 0; s < G; ++s) { if(counts[s] > std::max(e->sh.seg_chunks, e->sh.arena_chunks)) return fail(e, JFGPU_ERR_ARG, "more chunks than a receive segment holds"); total += counts[s]; }
  if(total == 0) return JFGPU_OK;
  // room in the CTAs' arenas of the local pool (as in run_batch): drain first when the bound says they could fill up
  const uint64_t usable = CHUNK_BYTES / ps.rec_bytes - ps.margin;
  const uint64_t per_cta = (total + e->n_sm - 1) / e->n_sm + 1;                  // (chunk j goes to CTA j mod grid)
  const uint64_t need = per_cta * (CHUNK_BYTES / 4) / usable + 2;
  if(ps.bound_chunks + need > ps.arena_chunks) { rc = part_drain(e, st); if(rc) return rc; }
  if(ps.bound_chunks + need > ps.arena_chunks) return fail(e, JFGPU_ERR_NOMEM, "record pool smaller than one exchange round");
  ps.bound_chunks += need;
  ps.pending = true;
  RestageArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.T = table_dev(e, e->tab);
  ra.recv_pool = e->sh.recv_pool; ra.recv_dir = e->sh.recv_dir; ra.n_src = G; ra.seg_chunks = (uint32_t)e->sh.seg_chunks;
  for(uint32_t s = 0; s < G; ++s) ra.count[s] = (uint32_t)counts[s];
  ra.first_region = e->p.shard_index * e->sh.own_regions; ra.split_lg = e->sh.split_lg; ra.sbits = e->sh.sbits;
  ra.inv_lut = e->tab.inv_lut.as<uint64_t>(); ra.nbytes = e->nbytes;
  if(self_bank <= 1) {             // this shard's own chunks stay in its arena of that send bank
    const size_t a0 = ((size_t)self_bank * G + e->p.shard_index) * e->sh.arena_chunks;
    ra.self_pool = e->sh.send_pool + a0 * CHUNK_BYTES; ra.self_dir = e->sh.send_dir + a0; ra.self_src = e->p.shard_index;
  }
  PartDev pd = part_dev(e);
  const size_t smem = (size_t)RING_P * 8 + (size_t)RING_P * RING * 4;
  cudaFuncSetAttribute(restage_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = e->n_sm;
  restage_kernel<1><<<grid, 1024, smem, st>>>(ra, pd); JF_LAUNCHED();
  CUDA_OK(e, cudaGetLastError());
  return JFGPU_OK;
}

int jfgpu_insert_keys(jfgpu_handle e, const void* dev_keys, uint64_t n, void* stream) {
  if(!e) return JFGPU_ERR_ARG;
  if(!e->tab.slots.p) return fail(e, JFGPU_ERR_STATE, "this engine holds a Bloom counter, not a hash table");
  cudaSetDevice(e->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->cs;
  if(n == 0) return JFGPU_OK;
  if(!stream) cudaEventRecord(e->ev_t0, st);
  int rc = JFGPU_OK;
  PartState& ps = e->part;
  if(ps.P) {
    // region-by-region mode: turn the keys into records of the pool (K1c); K2 inserts them at the next drain
    rc = part_alloc(e);
    if(rc) return rc;
    const uint64_t usable = CHUNK_BYTES / ps.rec_bytes - ps.margin;
    const uint64_t per_cta = (n + e->n_sm - 1) / e->n_sm + 1024 * 32;     // keys a CTA of stage_keys_kernel handles at most
    const uint64_t need = per_cta / usable + 2;
    if(ps.bound_chunks + need > ps.arena_chunks) { rc = part_drain(e, st); if(rc) return rc; }
    if(ps.P && ps.bound_chunks + need > ps.arena_chunks) return fail(e, JFGPU_ERR_NOMEM, "record pool smaller than one batch of keys");
  }
  if(ps.P) {
    const uint64_t usable = CHUNK_BYTES / ps.rec_bytes - ps.margin;
    ps.bound_chunks += ((n + e->n_sm - 1) / e->n_sm + 1024 * 32) / usable + 2;
    ps.pending = true;
    PartDev pd = part_dev(e);
    TableDev T = table_dev(e, e->tab);
    const size_t smem = (size_t)e->nbytes * 256 * 8 + PMAX * 8;
    const int grid = (int)std::min<uint64_t>((n + 1024ull * 32 - 1) / (1024ull * 32), (uint64_t)e->n_sm);
    if(e->kw == 1) {
      cudaFuncSetAttribute(stage_keys_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      stage_keys_kernel<1><<<grid, 1024, smem, st>>>(T, pd, e->tab.lut.as<uint64_t>(), e->nbytes, (const uint64_t*)dev_keys, n, nullptr);
    } else {
      cudaFuncSetAttribute(stage_keys_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      stage_keys_kernel<2><<<grid, 1024, smem, st>>>(T, pd, e->tab.lut.as<uint64_t>(), e->nbytes, (const uint64_t*)dev_keys, n, nullptr);
    }
    JF_LAUNCHED();
    CUDA_OK(e, cudaGetLastError());
  } else {
    rc = insert_keys_into(e, e->tab, (const uint64_t*)dev_keys, nullptr, n, st);
    if(rc) return rc;
  }
  if(stream) return JFGPU_OK;          // stream-ordered: the caller synchronises
  cudaEventRecord(e->ev_t1, st);
  CUDA_OK(e, cudaStreamSynchronize(st));
  float ms = 0; cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1); e->count_ms += ms;
  return JFGPU_OK;
}

int jfgpu_set_op(jfgpu_handle e, uint32_t op) {
  if(!e || op > JFGPU_OP_UPDATE) return JFGPU_ERR_ARG;
  // records staged under the previous operation must reach the table under that operation
  int rc = jfgpu_finish(e, nullptr);
  if(rc) return rc;
  e->op = op;
  return JFGPU_OK;
}

int jfgpu_set_spill(jfgpu_handle e, jfgpu_spill_fn fn, void* ctx) {
  if(!e) return JFGPU_ERR_ARG;
  if(e->shard_bits) return fail(e, JFGPU_ERR_STATE, "spilling is not supported on a sharded table");
  e->spill_fn = fn; e->spill_ctx = ctx;
  return JFGPU_OK;
}

int jfgpu_clear(jfgpu_handle e) {
  if(!e) return JFGPU_ERR_ARG;
  cudaSetDevice(e->device);
  CUDA_OK(e, cudaStreamSynchronize(e->hs));
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  resolve_kernel_events(e);
  if(e->tab.slots.p) {
    CUDA_OK(e, cudaMemsetAsync(e->tab.slots.p, 0, e->tab.bytes(), e->cs));
    CUDA_OK(e, cudaMemsetAsync(e->tab.ovf_keys.p, 0, e->tab.ovf_size * 8, e->cs));
    CUDA_OK(e, cudaMemsetAsync(e->tab.ovf_vals.p, 0, e->tab.ovf_size * 8, e->cs));
  }
  if(e->bloom.bits.p && e->bloom.mode != BLOOM_CHECK) CUDA_OK(e, cudaMemsetAsync(e->bloom.bits.p, 0, e->bloom.bits.bytes, e->cs));
  CUDA_OK(e, cudaMemsetAsync(e->stats.p, 0, STAT_N * 8, e->cs));
  if(e->part.pool.p) {
    CUDA_OK(e, cudaMemsetAsync(e->part.pool_next.p, 0, e->part.pool_next.bytes, e->cs));
    CUDA_OK(e, cudaMemsetAsync(e->part.spill_n.p, 0, 8, e->cs));
    CUDA_OK(e, cudaMemsetAsync(e->part.cta_chunk.p, 0xFF, e->part.cta_chunk.bytes, e->cs));
    CUDA_OK(e, cudaMemsetAsync(e->part.cta_fill.p, 0, e->part.cta_fill.bytes, e->cs));
    e->part.bound_chunks = e->part.P;
    e->part.pending = false;
  }
  if(e->sh.on) {
    for(int i = 0; i < 2; ++i) CUDA_OK(e, cudaMemsetAsync(e->sh.pool_next[i].p, 0, e->sh.pool_next[i].bytes, e->cs));
    CUDA_OK(e, cudaMemsetAsync(e->sh.cta_chunk.p, 0xFF, e->sh.cta_chunk.bytes, e->cs));
    CUDA_OK(e, cudaMemsetAsync(e->sh.cta_fill.p, 0, e->sh.cta_fill.bytes, e->cs));
  }
  e->bytes_fed = 0; e->count_ms = 0; e->kernel_ms = 0; e->kernel_launches = 0; e->drain_ms = 0; e->win_ms[0] = e->win_ms[1] = e->win_ms[2] = 0;
  e->eff_val_len = e->p.counter_len;
  return reset_carry(e, e->cs);
}

int jfgpu_get_stats(jfgpu_handle e, jfgpu_stats* s) {
  if(!e || !s) return JFGPU_ERR_ARG;
  cudaSetDevice(e->device);
  int rc = read_stats(e);
  if(rc) return rc;
  s->kmers = e->h_stats[STAT_KMERS];
  s->inserted = e->h_stats[STAT_INSERTED];
  s->distinct = e->h_stats[STAT_DISTINCT];
  s->reprobes = e->h_stats[STAT_REPROBES];
  s->overflowed = e->h_stats[STAT_OVERFLOWED];
  s->regrows = e->regrows;
  s->bytes = e->bytes_fed;
  s->seconds_count = e->count_ms * 1e-3;
  resolve_kernel_events(e);
  s->seconds_count_kernel = e->kernel_ms * 1e-3;
  s->count_kernel_launches = e->kernel_launches;
  s->seconds_drain = e->drain_ms * 1e-3;
  s->seconds_win_hist = e->win_ms[0] * 1e-3; s->seconds_win_scatter = e->win_ms[1] * 1e-3; s->seconds_win_insert = e->win_ms[2] * 1e-3;
  return JFGPU_OK;
}

int jfgpu_finish(jfgpu_handle e, jfgpu_stats* s) {
  if(!e) return JFGPU_ERR_ARG;
  cudaSetDevice(e->device);
  CUDA_OK(e, cudaStreamSynchronize(e->hs));
  int rc = part_drain(e, e->cs);
  if(rc) return rc;
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  CUDA_OK(e, cudaGetLastError());
  rc = e->tab.slots.p ? check_after_batches(e) : read_stats(e);
  if(rc) return rc;
  if(e->h_stats[STAT_FORMAT_ERR]) return fail(e, JFGPU_ERR_FORMAT, "Invalid fastq sequence (the device parser reads 4-line FASTQ records: '@' header, sequence, '+', qualities)");
  if(e->h_stats[STAT_POOL_FULL]) return fail(e, JFGPU_ERR_NOMEM, "internal: k-mer record pool overflow");
  if(e->h_stats[STAT_ROUTE_DROPPED]) return fail(e, JFGPU_ERR_FULL, "route bucket capacity exceeded");
  if(e->tab.slots.p) { rc = direct_index_fixup(e); if(rc) return rc; }
  if(s) return jfgpu_get_stats(e, s);
  return JFGPU_OK;
}

int jfgpu_table_info_get(jfgpu_handle e, jfgpu_table_info* info) {
  if(!e || !info) return JFGPU_ERR_ARG;
  if(!e->tab.slots.p) return fail(e, JFGPU_ERR_STATE, "this engine holds a Bloom counter, not a hash table");
  const Table& t = e->tab;
  info->size = t.size; info->lsize = t.lsize; info->key_len = 2 * e->k; info->val_len = e->eff_val_len;
  info->max_reprobe = t.max_reprobe; info->matrix_r = t.M.r(); info->matrix_c = t.M.c();
  info->matrix_identity = t.M.is_low_identity() ? 1 : 0;
  info->slot_bits = t.slot_bits; info->local_slots = t.local_slots; info->table_bytes = t.bytes();
  e->matrix_cols_host.assign(t.M.c(), 0);
  for(unsigned i = 0; i < t.M.c(); ++i) e->matrix_cols_host[i] = t.M[i];
  info->matrix_columns = t.M.is_identity() ? nullptr : e->matrix_cols_host.data();
  info->reprobes = t.reprobes.data();
  info->part_regions = e->part.P; info->part_rec_bytes = e->part.rec_bytes;
  return JFGPU_OK;
}

int jfgpu_dump(jfgpu_handle e, uint64_t lower, uint64_t upper, uint32_t ocl, jfgpu_sink_fn sink, void* ctx, uint64_t* n_records) {
  if(!e || !sink) return JFGPU_ERR_ARG;
  if(!e->tab.slots.p) return fail(e, JFGPU_ERR_STATE, "this engine holds a Bloom counter, not a hash table");
  if(ocl < 1 || ocl > 8) return fail(e, JFGPU_ERR_ARG, "out_counter_len must be in [1, 8]");
  cudaSetDevice(e->device);
  int rc = e->in_spill ? JFGPU_OK : jfgpu_finish(e, nullptr);      // (from inside the spill hook the table is dumped as it stands)
  if(rc) return rc;
  // Segment by segment, two buffers: while the host hands segment i to the sink, the device sorts and serialises segment
  // i+1 (jf_dump.cuh: no global sort, every tile of 8192 positions is ordered in shared memory) and the copy engine brings
  // it to pinned host memory.
  Table& t = e->tab;
  const unsigned rec = e->nbytes + ocl;
  const uint64_t seg = std::min<uint64_t>(t.local_size, (uint64_t)1 << 24);
  const uint64_t cap = seg + t.margin + 8;                    // records of a segment at most
  const uint32_t max_tiles = (uint32_t)((seg + DUMP_TP - 1) / DUMP_TP);
  DevBuf tile_cnt[2], out[2];
  uint8_t* hbuf[2] = { nullptr, nullptr };
  uint32_t* h_total = nullptr;
  cudaEvent_t ev_emit[2] = { nullptr, nullptr }, ev_copy[2] = { nullptr, nullptr };
  bool ok = cudaHostAlloc((void**)&h_total, 2 * sizeof(uint32_t), cudaHostAllocDefault) == cudaSuccess;
  for(int i = 0; i < 2 && ok; ++i)
    ok = tile_cnt[i].alloc(((size_t)max_tiles + 1) * 4) == cudaSuccess && out[i].alloc(cap * rec + 16) == cudaSuccess &&
         cudaHostAlloc((void**)&hbuf[i], cap * rec + 16, cudaHostAllocDefault) == cudaSuccess &&
         cudaEventCreateWithFlags(&ev_emit[i], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming) == cudaSuccess;
  if(!ok) { cudaGetLastError(); rc = fail(e, JFGPU_ERR_NOMEM, "allocation of the dump buffers failed"); }
  const size_t smem = (size_t)e->nbytes * 256 * 8 + ((size_t)DUMP_TP + 1) * 4 + (size_t)DUMP_MAXC * 2 + (size_t)DUMP_NTH * rec;
  uint64_t total = 0;
  const uint64_t n_seg = (t.local_size + seg - 1) / seg;
  auto launch = [&](uint64_t si) -> int {            // count, scan, emit of segment si on the compute stream; its total follows
    const int b = (int)(si & 1);
    DumpArgs a;
    memset(&a, 0, sizeof(a));
    a.T = table_dev(e, t);
    a.inv_lut = t.inv_lut.as<uint64_t>(); a.nbytes = e->nbytes; a.ocl = ocl;
    a.seg_lo = si * seg; a.seg_hi = std::min(a.seg_lo + seg, t.local_size);
    a.slots_end = t.local_size + t.margin; a.margin = t.margin;
    a.lower = lower; a.upper = upper;
    a.n_tiles = (uint32_t)((a.seg_hi - a.seg_lo + DUMP_TP - 1) / DUMP_TP);
    a.tile_cnt = tile_cnt[b].as<uint32_t>(); a.out = out[b].as<uint8_t>(); a.out_cap = cap;
    const int grid = (int)std::min<uint64_t>(a.n_tiles, (uint64_t)e->n_sm * 8);
    return dispatch(e, e->kw, t.slot_bits, [&](auto KW, auto SB) -> int {
      constexpr int kw = decltype(KW)::value, sb = decltype(SB)::value;
      dump_count_kernel<sb><<<grid, DUMP_NTH, 0, e->cs>>>(a); JF_LAUNCHED();
      dump_scan_kernel<<<1, 1024, 0, e->cs>>>(a.tile_cnt, a.n_tiles); JF_LAUNCHED();
      auto kern = dump_emit_kernel<kw, sb>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, DUMP_NTH, smem, e->cs>>>(a); JF_LAUNCHED();
      CUDA_OK(e, cudaMemcpyAsync(h_total + b, a.tile_cnt + a.n_tiles, 4, cudaMemcpyDeviceToHost, e->cs));
      CUDA_OK(e, cudaEventRecord(ev_emit[b], e->cs));
      return JFGPU_OK;
    });
  };
  uint64_t n_in_buf[2] = { 0, 0 };
  if(!rc && n_seg) rc = launch(0);
  for(uint64_t si = 0; si < n_seg && !rc; ++si) {
    const int b = (int)(si & 1);
    // the segment's size, then its bytes on the copy stream
    cudaError_t c = cudaEventSynchronize(ev_emit[b]);
    if(c != cudaSuccess) { rc = fail(e, JFGPU_ERR_CUDA, std::string("dump: ") + cudaGetErrorString(c)); break; }
    n_in_buf[b] = h_total[b];
    if(n_in_buf[b] > cap) { rc = fail(e, JFGPU_ERR_STATE, "internal: segment overflow in dump"); break; }
    if(n_in_buf[b]) cudaMemcpyAsync(hbuf[b], out[b].p, n_in_buf[b] * rec, cudaMemcpyDeviceToHost, e->hs);
    cudaEventRecord(ev_copy[b], e->hs);
    // the next segment is produced while this one is copied and written (its buffers were released two rounds ago)
    if(si + 1 < n_seg) { rc = launch(si + 1); if(rc) break; }
    c = cudaEventSynchronize(ev_copy[b]);
    if(c != cudaSuccess) { rc = fail(e, JFGPU_ERR_CUDA, std::string("dump: ") + cudaGetErrorString(c)); break; }
    if(n_in_buf[b] && sink(ctx, hbuf[b], n_in_buf[b] * rec) != 0) { rc = fail(e, JFGPU_ERR_SINK, "dump sink failed"); break; }
    total += n_in_buf[b];
  }
  cudaStreamSynchronize(e->cs); cudaStreamSynchronize(e->hs);
  for(int i = 0; i < 2; ++i) {
    tile_cnt[i].free(); out[i].free();
    if(hbuf[i]) cudaFreeHost(hbuf[i]);
    if(ev_emit[i]) cudaEventDestroy(ev_emit[i]);
    if(ev_copy[i]) cudaEventDestroy(ev_copy[i]);
  }
  if(h_total) cudaFreeHost(h_total);
  if(n_records) *n_records = total;
  return rc;
}

int jfgpu_lookup(jfgpu_handle e, const uint64_t* keys, size_t n, uint64_t* vals) {
  if(!e || (n && (!keys || !vals))) return JFGPU_ERR_ARG;
  if(!e->tab.slots.p) return fail(e, JFGPU_ERR_STATE, "this engine holds a Bloom counter, not a hash table");
  if(n == 0) return JFGPU_OK;
  cudaSetDevice(e->device);
  { int rc0 = jfgpu_finish(e, nullptr); if(rc0) return rc0; }
  DevBuf dk, dv;
  CUDA_OK(e, dk.alloc(n * 8 * e->kw));
  CUDA_OK(e, dv.alloc(n * 8));
  CUDA_OK(e, cudaMemcpyAsync(dk.p, keys, n * 8 * e->kw, cudaMemcpyHostToDevice, e->cs));
  TableDev T = table_dev(e, e->tab);
  const size_t smem = (size_t)e->nbytes * 256 * 8;
  const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)e->n_sm * 8);
  int rc = dispatch(e, e->kw, e->tab.slot_bits, [&](auto KW, auto SB) -> int {
    auto kern = lookup_kernel<decltype(KW)::value, decltype(SB)::value>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 256, smem, e->cs>>>(T, e->tab.lut.as<uint64_t>(), e->nbytes, dk.as<uint64_t>(), n, dv.as<uint64_t>(), e->shard_bits);
    return JFGPU_OK;
  });
  if(!rc) { JF_LAUNCHED();
    cudaError_t c = cudaMemcpyAsync(vals, dv.p, n * 8, cudaMemcpyDeviceToHost, e->cs);
    if(c == cudaSuccess) c = cudaStreamSynchronize(e->cs);
    if(c != cudaSuccess) rc = fail(e, JFGPU_ERR_CUDA, std::string("lookup: ") + cudaGetErrorString(c));
  }
  dk.free(); dv.free();
  return rc;
}

int jfgpu_histogram(jfgpu_handle e, uint64_t* hist, uint32_t n_bins) {
  if(!e || !hist || n_bins == 0) return JFGPU_ERR_ARG;
  if(!e->tab.slots.p) return fail(e, JFGPU_ERR_STATE, "this engine holds a Bloom counter, not a hash table");
  cudaSetDevice(e->device);
  int rc = jfgpu_finish(e, nullptr);
  if(rc) return rc;
  DevBuf dh;
  CUDA_OK(e, dh.alloc((size_t)n_bins * 8));
  CUDA_OK(e, cudaMemsetAsync(dh.p, 0, (size_t)n_bins * 8, e->cs));
  TableDev T = table_dev(e, e->tab);
  const uint64_t ns = e->tab.local_size + e->tab.margin;
  const int grid = (int)std::min<uint64_t>((ns + 255) / 256, (uint64_t)e->n_sm * 16);
  switch(e->tab.slot_bits) {
  case 32:  histogram_kernel<32><<<grid, 256, 0, e->cs>>>(T, ns, dh.as<unsigned long long>(), n_bins); break;
  case 64:  histogram_kernel<64><<<grid, 256, 0, e->cs>>>(T, ns, dh.as<unsigned long long>(), n_bins); break;
  default:  histogram_kernel<128><<<grid, 256, 0, e->cs>>>(T, ns, dh.as<unsigned long long>(), n_bins); break;
  }
  JF_LAUNCHED();
  cudaError_t c = cudaMemcpyAsync(hist, dh.p, (size_t)n_bins * 8, cudaMemcpyDeviceToHost, e->cs);
  if(c == cudaSuccess) c = cudaStreamSynchronize(e->cs);
  dh.free();
  if(c != cudaSuccess) return fail(e, JFGPU_ERR_CUDA, std::string("histogram: ") + cudaGetErrorString(c));
  return JFGPU_OK;
}

int jfgpu_bloom_info_get(jfgpu_handle e, jfgpu_bloom_info* info) {
  if(!e || !info) return JFGPU_ERR_ARG;
  cudaSetDevice(e->device);
  memset(info, 0, sizeof(*info));
  BloomState& b = e->bloom;
  info->mode = b.mode;
  if(b.mode == BLOOM_NONE) return JFGPU_OK;
  if(!b.drawn) { int rc = bloom_draw(e); if(rc) return rc; }
  info->nb_hashes = b.k; info->m = b.m;
  info->nb_bytes = b.mode == BLOOM_FILTER ? b.m / 8 + (b.m % 8 != 0) : b.m / 5 + (b.m % 5 != 0);
  info->matrix_r = 64; info->matrix_c = 2 * e->k;
  info->matrix1 = b.cols1.data(); info->matrix2 = b.cols2.data();
  return JFGPU_OK;
}

int jfgpu_bloom_load(jfgpu_handle e, uint64_t m, uint32_t nb_hashes, const uint64_t* c1, const uint64_t* c2, const void* bytes, size_t nbytes) {
  if(!e || !c1 || !c2 || !bytes) return JFGPU_ERR_ARG;
  cudaSetDevice(e->device);
  if(!e->tab.slots.p) return fail(e, JFGPU_ERR_STATE, "this engine holds a Bloom counter, not a hash table");
  if(e->bloom.mode != BLOOM_NONE) return fail(e, JFGPU_ERR_STATE, "a Bloom filter is already attached to this engine");
  if(m == 0 || nb_hashes == 0) return fail(e, JFGPU_ERR_ARG, "empty Bloom counter");
  if(nbytes < m / 5 + (m % 5 != 0)) return fail(e, JFGPU_ERR_ARG, "Bloom filter file is truncated");
  int rc = jfgpu_finish(e, nullptr);          // text fed so far is counted unfiltered
  if(rc) return rc;
  BloomState& b = e->bloom;
  b.m = m; b.k = nb_hashes;
  b.inv = m == 1 ? ~(uint64_t)0 : (uint64_t)(((unsigned __int128)1 << 64) / m);
  b.n_words = (m + 31) / 32;
  DevBuf raw;
  const size_t nb = m / 5 + (m % 5 != 0);
  if(b.bits.alloc((size_t)b.n_words * 4 + 16) != cudaSuccess || raw.alloc(nb + 16) != cudaSuccess) {
    cudaGetLastError(); b.bits.free(); raw.free();
    return fail(e, JFGPU_ERR_NOMEM, "Failed to allocate the Bloom counter in device memory");
  }
  CUDA_OK(e, cudaMemcpyAsync(raw.p, bytes, nb, cudaMemcpyHostToDevice, e->cs));
  const int grid = (int)std::min<uint64_t>((b.n_words + 255) / 256, (uint64_t)e->n_sm * 16);
  bloom_unpack_kernel<<<grid, 256, 0, e->cs>>>(raw.as<uint8_t>(), m, b.n_words, b.bits.as<uint32_t>()); JF_LAUNCHED();
  CUDA_OK(e, cudaStreamSynchronize(e->cs));
  raw.free();
  b.M1 = jfb::gf2_matrix(64, 2 * e->k, c1); b.M2 = jfb::gf2_matrix(64, 2 * e->k, c2);
  b.mode = BLOOM_CHECK;
  return bloom_upload_matrices(e);
}

int jfgpu_bloom_dump(jfgpu_handle e, jfgpu_sink_fn sink, void* ctx) {
  if(!e || !sink) return JFGPU_ERR_ARG;
  cudaSetDevice(e->device);
  BloomState& b = e->bloom;
  if(b.mode != BLOOM_COUNT) return fail(e, JFGPU_ERR_STATE, "no Bloom counter has been built by this engine");
  int rc = jfgpu_finish(e, nullptr);
  if(rc) return rc;
  const uint64_t nb = b.m / 5 + (b.m % 5 != 0);
  const uint64_t piece = (uint64_t)64 << 20;
  DevBuf out; uint8_t* hbuf = nullptr;
  if(out.alloc(piece) != cudaSuccess || cudaHostAlloc((void**)&hbuf, piece, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError(); out.free();
    return fail(e, JFGPU_ERR_NOMEM, "allocation of the Bloom counter staging buffers failed");
  }
  for(uint64_t off = 0; off < nb && !rc; off += piece) {
    const uint64_t len = std::min(piece, nb - off);
    const int grid = (int)std::min<uint64_t>((len + 255) / 256, (uint64_t)e->n_sm * 16);
    // positions 5*off .. : the kernel takes the bit array shifted by whole words (5*off*2 bits; piece is a multiple of 16 bytes)
    bloom_pack_kernel<<<grid, 256, 0, e->cs>>>(b.bits.as<uint32_t>() + (5 * off) / 16, b.m - 5 * off, len, out.as<uint8_t>()); JF_LAUNCHED();
    cudaError_t c = cudaMemcpyAsync(hbuf, out.p, len, cudaMemcpyDeviceToHost, e->cs);
    if(c == cudaSuccess) c = cudaStreamSynchronize(e->cs);
    if(c != cudaSuccess) { rc = fail(e, JFGPU_ERR_CUDA, std::string("bloom dump: ") + cudaGetErrorString(c)); break; }
    if(sink(ctx, hbuf, len) != 0) rc = fail(e, JFGPU_ERR_SINK, "dump sink failed");
  }
  cudaFreeHost(hbuf); out.free();
  return rc;
}

uint64_t jfgpu_synth_fasta_bytes(uint64_t n_bases) {
  return SYNTH_HDR + n_bases + (n_bases + SYNTH_LINE - 1) / SYNTH_LINE;
}

int jfgpu_synth_fasta_device(int device, void* dev_out, uint64_t capacity, uint64_t n_bases, uint64_t seed, uint64_t* n_bytes, void* stream) {
  const uint64_t need = jfgpu_synth_fasta_bytes(n_bases);
  if(!dev_out || capacity < need) return JFGPU_ERR_ARG;
  if(cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return JFGPU_ERR_CUDA; }
  const int grid = (int)std::min<uint64_t>((need + 255) / 256, (uint64_t)148 * 32);
  synth_fasta_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint8_t*)dev_out, need, n_bases, seed);
  JF_LAUNCHED();
  if(cudaGetLastError() != cudaSuccess) return JFGPU_ERR_CUDA;
  if(n_bytes) *n_bytes = need;
  return JFGPU_OK;
}

}  // extern "C"
