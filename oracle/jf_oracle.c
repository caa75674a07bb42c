/* jf_oracle.c -- CPU restatement of the `jellyfish count` hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build or run this
 * program, and only as the checker.  The product (jellyfish_b200/) never links or calls it.
 *
 * Parity is PINNED: tests/test_oracle.py checks this program byte for byte against the
 * unmodified reference built in oracle/_ref (which itself reproduces the reference's golden
 * md5s of tests/parallel_hashing.sh:10-12), and against fixtures in tests/golden/.
 *
 * It is written for clarity, not speed: parse -> collect every canonical k-mer -> sort ->
 * count runs -> hash -> sort by (position, key) -> write.  Each step cites the reference code
 * whose observable behaviour it restates (paths relative to /root/reference).
 *
 *   jf_oracle count -m K -s SIZE [-C] [-c VAL_LEN] [-p REPROBES] [--out-counter-len N]
 *                   [-L LOW] [-U HIGH] [--text] [--if FILE]... [-Q CHAR | --min-quality N
 *                   [--quality-start S]] [--bf-size N [--bf-fp P]] [-o OUT] file...
 *   jf_oracle bc -m K -s N [-f FPR] [-C] -o OUT file...     the Bloom counter file of `jellyfish bc`
 *   (count also takes --bc FILE: keep the k-mers that file has seen at least twice)
 *   jf_oracle matrix R C [SKIP]        print the hash matrix columns the reference would draw
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <math.h>
#include <sys/mman.h>

typedef unsigned __int128 u128;

/* ---- glibc random(), TYPE_3, default seed 1: the stream lib/misc.cc:66-72 consumes ---------- */
static int32_t rnd_r[34];
static int rnd_f, rnd_b;
static uint32_t rnd_step(void) {
  uint32_t v = (uint32_t)rnd_r[rnd_f] + (uint32_t)rnd_r[rnd_b];
  rnd_r[rnd_f] = (int32_t)v;
  if(++rnd_f >= 31) rnd_f = 0;
  if(++rnd_b >= 31) rnd_b = 0;
  return v >> 1;
}
static void rnd_seed(uint32_t seed) {
  rnd_r[0] = (int32_t)seed;
  for(int i = 1; i < 31; ++i) {
    int64_t w = (16807LL * rnd_r[i - 1]) % 2147483647LL;
    if(w < 0) w += 2147483647LL;
    rnd_r[i] = (int32_t)w;
  }
  rnd_f = 3; rnd_b = 0;
  for(int i = 0; i < 310; ++i) rnd_step();
}
/* random_bits(64): lib/misc.cc:66-72 */
static uint64_t random_bits64(void) {
  uint64_t res = 0;
  for(int i = 0; i < 64; i += 30) res ^= (uint64_t)rnd_step() << i;   /* step = floor(log2(RAND_MAX)) = 30: draws overlap by one bit */
  return res;
}

/* ---- GF(2) matrix: include/jellyfish/rectangular_binary_matrix.hpp, lib/...matrix.cc -------- */
typedef struct { unsigned r, c; int identity; uint64_t col[256]; } matrix_t;

/* times(): XOR of columns[c-1-i] over set bits i (rectangular_binary_matrix.hpp:223-261) */
static uint64_t mat_times(const matrix_t* m, u128 v) {
  if(m->identity) return (uint64_t)v & (m->r >= 64 ? ~0ULL : ((1ULL << m->r) - 1));
  uint64_t res = 0;
  for(unsigned i = 0; i < m->c; ++i)
    if((v >> i) & 1) res ^= m->col[m->c - 1 - i];
  return res;
}
/* pseudo_inverse(): lib/rectangular_binary_matrix.cc:160-210.  Returns 0 when singular. */
static int mat_pseudo_inverse(const matrix_t* m, matrix_t* res) {
  uint64_t piv[256];
  memcpy(piv, m->col, sizeof(uint64_t) * m->c);
  res->r = m->r; res->c = m->c; res->identity = 0;
  memset(res->col, 0, sizeof(res->col));
  unsigned srow = m->r < m->c ? m->r : m->c, scol = m->c - srow;
  for(unsigned i = scol; i < m->c; ++i) res->col[i] = 1ULL << (srow - 1 - (i - scol));
  uint64_t mask = 1ULL << (srow - 1);
  for(unsigned i = scol; i < m->c; ++i, mask >>= 1) {
    if(!(piv[i] & mask)) {
      unsigned j = i + 1;
      while(j < m->c && !(piv[j] & mask)) ++j;
      if(j == m->c) return 0;
      piv[i] ^= piv[j]; res->col[i] ^= res->col[j];
    }
    for(unsigned j = i + 1; j < m->c; ++j)
      if(piv[j] & mask) { piv[j] ^= piv[i]; res->col[j] ^= res->col[i]; }
  }
  mask = 1ULL << (srow - 1);
  for(unsigned i = scol; i < m->c; ++i, mask >>= 1)
    for(unsigned j = 0; j < i; ++j)
      if(piv[j] & mask) { piv[j] ^= piv[i]; res->col[j] ^= res->col[i]; }
  return 1;
}
/* randomize_pseudo_inverse(): lib/rectangular_binary_matrix.cc:240-247 */
static void mat_draw(unsigned r, unsigned c, matrix_t* out) {
  matrix_t m;
  m.r = r; m.c = c; m.identity = 0;
  uint64_t cmask = r >= 64 ? ~0ULL : ((1ULL << r) - 1);
  do {
    for(unsigned i = 0; i < c; ++i) m.col[i] = random_bits64() & cmask;
  } while(!mat_pseudo_inverse(&m, out));
}

/* ---- k-mer extraction ---------------------------------------------------------------------- */
static unsigned K;
static int canonical;
static u128* mers; static size_t n_mers, cap_mers;
static int bf_filter(u128 m);
static int bf_on;
static void bc_insert(u128 m); static int bc_check(u128 m);
static int bc_build, bc_on;
static void emit(u128 m) {
  if(bc_build) { bc_insert(m); return; }   /* `jellyfish bc`: bc_main.cc:66-70 */
  if(bf_on && !bf_filter(m)) return;   /* filter_bf: count_main.cc:122-133,157-161 */
  if(bc_on && bc_check(m) <= 1) return;    /* filter_bc: count_main.cc:110-120 */
  if(n_mers == cap_mers) { cap_mers = cap_mers ? cap_mers * 2 : (1 << 20); mers = realloc(mers, cap_mers * sizeof(u128)); if(!mers) { perror("realloc"); exit(1); } }
  mers[n_mers++] = m;
}
/* codes[]: include/jellyfish/mer_dna.hpp:38-55 */
static int code(int c) {
  switch(c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; }
  return -1;
}
static u128 fwd, rev; static unsigned filled;
/* one character of clean sequence: mer_iterator.hpp:67-76; shift_left / shift_right: mer_dna.hpp:322-370 */
static void feed_char(int ch) {
  int c = code(ch);
  if(c < 0) { filled = 0; return; }
  u128 mask = K == 64 ? ~(u128)0 : (((u128)1 << (2 * K)) - 1);
  fwd = ((fwd << 2) | (u128)c) & mask;
  rev = (rev >> 2) | ((u128)(3 - c) << (2 * K - 2));
  if(filled < K) ++filled;
  if(filled >= K) emit(canonical && rev < fwd ? rev : fwd);   /* operator<: mer_dna.hpp:227-250 */
}
/* One file: mer_overlap_sequence_parser.hpp:134-148 (format sniffing), :161-185 read_fasta
 * (headers dropped, an 'N' between records), :260-274 read_sequence ('\n' and line-end '\r'
 * dropped, lines concatenated), :111 (no k-mer across files). */
static int count_file(const char* path) {
  FILE* f = fopen(path, "rb");
  if(!f) { fprintf(stderr, "Can't open file '%s'\n", path); return 0; }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  unsigned char* d = malloc(n + 1);
  if(n && fread(d, 1, n, f) != (size_t)n) { perror("fread"); exit(1); }
  fclose(f);
  filled = 0; fwd = rev = 0;
  if(n == 0) { free(d); return 1; }
  if(d[0] != '>' && d[0] != '@') { fprintf(stderr, "Unsupported format\n"); exit(134); }
  long p = 0;
  if(d[0] == '@') {
    /* FASTQ: mer_overlap_sequence_parser.hpp:187-217 (read_fastq) and :290-307 (skip_quals):
     * header line dropped; sequence lines up to the line starting with '+'; the '+' line dropped;
     * then exactly as many quality characters as there were sequence characters are skipped (BY
     * LENGTH, across lines), after which '@' or EOF must follow; an 'N' separates reads. */
    while(p < n && d[p] != '\n') ++p;                       /* first header */
    if(p < n) ++p;
    for(;;) {
      long seq_len = 0;
      for(;;) {                                              /* sequence lines */
        while(p < n && (d[p] == '\n' || d[p] == '\r')) ++p;
        if(p >= n || d[p] == '+') break;
        long e = p; while(e < n && d[e] != '\n') ++e;
        long t = e; while(t > p && d[t - 1] == '\r') --t;
        for(long i = p; i < t; ++i) feed_char(d[i]);
        seq_len += t - p;
        p = e;
      }
      if(p >= n) break;
      while(p < n && d[p] != '\n') ++p;                      /* the '+' line */
      if(p < n) ++p;
      long quals = 0;
      while(p < n && quals < seq_len) {                      /* qualities, by length */
        while(p < n && (d[p] == '\n' || d[p] == '\r')) ++p;
        long e = p; while(e < n && d[e] != '\n' && quals + (e - p) < seq_len) ++e;
        long t = e; while(t > p && d[t - 1] == '\r' && e < n && d[e] == '\n') --t;
        quals += t - p;
        p = e;
      }
      while(p < n && (d[p] == '\n' || d[p] == '\r')) ++p;
      if(p >= n) break;
      if(quals != seq_len || d[p] != '@') { fprintf(stderr, "Invalid fastq sequence\n"); exit(1); }
      feed_char('N');
      while(p < n && d[p] != '\n') ++p;                      /* next header */
      if(p < n) ++p;
    }
    free(d);
    filled = 0;
    return 1;
  }
  int at_line_start = 1;
  while(p < n) {
    if(at_line_start) {
      while(p < n && (d[p] == '\n' || d[p] == '\r')) ++p;                 /* skip_newlines */
      if(p >= n) break;
      if(d[p] == '>') { feed_char('N'); while(p < n && d[p] != '\n') ++p; if(p < n) ++p; continue; }   /* ignore_line + 'N' */
      at_line_start = 0;
    }
    long e = p; while(e < n && d[e] != '\n') ++e;                          /* istream::get up to '\n' */
    long t = e; while(t > p && d[t - 1] == '\r') --t;                      /* strip line-end '\r' */
    for(long i = p; i < t; ++i) feed_char(d[i]);
    p = e; at_line_start = 1;
  }
  free(d);
  filled = 0;
  return 1;
}

/* ---- --bf-size / --bf-fp: one-pass Bloom prefilter --------------------------------------------
 * count_main.cc:317-321 builds mer_dna_bloom_filter(bf_fp, bf_size) AFTER the table (and after the
 * --if pass), so its two 64 x 2k hash matrices (mer_dna_bloom_counter.hpp:22-36: randomize(), no
 * invertibility test) are the draws that follow the first table matrix.  bloom_common.hpp:62-67:
 * m = n * lrint(-ln fp / ln^2 2) bits, k = lrint(-ln fp / ln 2) functions.  bloom_filter.hpp:40-63
 * insert__: positions (h1 % m + i * (h2 % m)) % m, test-and-set, "present" = all k bits were set.
 * An occurrence reaches the table only when the filter already held the k-mer (or a false
 * positive).  The outcome depends on the ORDER of insertions: this restatement processes the
 * occurrences in input order, which is what the reference does with -t 1 (one thread, one
 * parser buffer at a time); with more threads the reference's own output varies in the false
 * positives. */
static unsigned char* bf_bits; static uint64_t bf_m; static unsigned long bf_k; static matrix_t bf_m1, bf_m2;
static void bf_setup(double fp, uint64_t n) {
  const double LOG2 = 0.6931471805599453, LOG2_SQ = 0.4804530139182014;
  bf_m = n * (uint64_t)lrint(-log(fp) / LOG2_SQ);
  bf_k = lrint(-log(fp) / LOG2);
  bf_bits = calloc(bf_m / 8 + (bf_m % 8 != 0) + 1, 1);
  if(!bf_bits) { perror("calloc"); exit(1); }
  bf_m1.r = bf_m2.r = 64; bf_m1.c = bf_m2.c = 2 * K; bf_m1.identity = bf_m2.identity = 0;
  for(unsigned i = 0; i < 2 * K; ++i) bf_m1.col[i] = random_bits64();
  for(unsigned i = 0; i < 2 * K; ++i) bf_m2.col[i] = random_bits64();
  bf_on = 1;
}
static int bf_filter(u128 m) {
  uint64_t base = mat_times(&bf_m1, m) % bf_m, inc = mat_times(&bf_m2, m) % bf_m;
  int present = 1;
  for(unsigned long i = 0; i < bf_k; ++i) {
    uint64_t pos = (base + i * inc) % bf_m;
    unsigned char mask = (unsigned char)(1u << (pos % 8));
    if(!(bf_bits[pos / 8] & mask)) { present = 0; bf_bits[pos / 8] |= mask; }
  }
  return present;
}

/* ---- `jellyfish bc` and `count --bc FILE`: two-pass Bloom counter ------------------------------
 * bloom_counter2.hpp:34-36,49-108: m positions, each a base-3 digit (0,1,2 = saturated) packed
 * five to a byte; insert increments the k positions (h1 % m + i * (h2 % m)) % m that are below 2.
 * Every position ends at min(2, number of hits): the file does not depend on insertion order or
 * thread count.  check = minimum digit over the k positions; `count --bc` keeps a k-mer
 * occurrence only if check > 1 (count_main.cc:110-120), i.e. all occurrences of k-mers seen at
 * least twice (plus false positives).  bc_main.cc:103-113: the two 64 x 2k matrices are the first
 * two draws of the random stream (randomize(), no invertibility test), m = n * lrint(-ln f / ln^2 2),
 * k = lrint(-ln f / ln 2); header format "bloomcounter" with matrix1, matrix2, size = m, nb_hashes = k. */
static const unsigned bc_pow3[5] = {1, 3, 9, 27, 81};
static void bc_insert(u128 m) {
  uint64_t base = mat_times(&bf_m1, m) % bf_m, inc = mat_times(&bf_m2, m) % bf_m;
  for(unsigned long i = 0; i < bf_k; ++i) {
    uint64_t pos = (base + i * inc) % bf_m;
    unsigned d = pos % 5;
    if(bf_bits[pos / 5] / bc_pow3[d] % 3 < 2) bf_bits[pos / 5] += bc_pow3[d];
  }
}
static int bc_check(u128 m) {
  uint64_t base = mat_times(&bf_m1, m) % bf_m, inc = mat_times(&bf_m2, m) % bf_m;
  unsigned res = 2;
  for(unsigned long i = 0; i < bf_k; ++i) {
    uint64_t pos = (base + i * inc) % bf_m;
    unsigned w = bf_bits[pos / 5] / bc_pow3[pos % 5] % 3;
    if(w < res) res = w;
  }
  return (int)res;
}
static const char* js_find(const char* js, const char* key) {
  const char* p = strstr(js, key);
  if(!p) { fprintf(stderr, "Failed to parse bloom filter file (no %s)\n", key); exit(1); }
  return p + strlen(key);
}
static void js_columns(const char* js, const char* name, matrix_t* m) {
  const char* p = js_find(js_find(js, name), "\"columns\":[");
  unsigned n = 0;
  while(*p && *p != ']') { char* e; m->col[n++] = strtoull(p, &e, 10); p = *e == ',' ? e + 1 : e; if(n > 256) break; }
  m->c = n; m->r = 64; m->identity = 0;
}
/* load_bloom_filter: count_main.cc:191-206 */
static void bc_load(const char* path) {
  FILE* f = fopen(path, "rb");
  if(!f) { fprintf(stderr, "Failed to parse bloom filter file '%s'\n", path); exit(1); }
  char digits[10] = {0};
  if(fread(digits, 1, 9, f) != 9) { fprintf(stderr, "Failed to parse bloom filter file '%s'\n", path); exit(1); }
  size_t hlen = strtoull(digits, 0, 10);
  char* js = calloc(hlen + 1, 1);
  if(fread(js, 1, hlen, f) != hlen) { fprintf(stderr, "Failed to parse bloom filter file '%s'\n", path); exit(1); }
  if(!strstr(js, "\"format\":\"bloomcounter\"")) { fprintf(stderr, "Invalid format. Expected 'bloomcounter'\n"); exit(1); }
  if(strtoul(js_find(js, "\"key_len\":"), 0, 10) != 2 * K) { fprintf(stderr, "Invalid mer length in bloom filter\n"); exit(1); }
  js_columns(js, "\"matrix1\":", &bf_m1); js_columns(js, "\"matrix2\":", &bf_m2);
  bf_m = strtoull(js_find(js, "\"size\":"), 0, 10);
  bf_k = strtoul(js_find(js, "\"nb_hashes\":"), 0, 10);
  size_t nb = bf_m / 5 + (bf_m % 5 != 0);
  bf_bits = calloc(nb + 1, 1);
  if(fread(bf_bits, 1, nb, f) != nb) { fprintf(stderr, "Bloom filter file is truncated\n"); exit(1); }
  fclose(f); free(js);
  bc_on = 1;
}
static uint64_t parse_size(const char* s);
static int count_file(const char* path);
static int bc_main(int argc, char** argv) {
  uint64_t n = 0; double fpr = 0.001; const char* out = 0; int first = argc;   /* bc_main_cmdline.yaggo */
  for(int i = 2; i < argc; ++i) {
    if(!strcmp(argv[i], "-m")) K = atoi(argv[++i]);
    else if(!strcmp(argv[i], "-s")) n = parse_size(argv[++i]);
    else if(!strcmp(argv[i], "-f")) fpr = atof(argv[++i]);
    else if(!strcmp(argv[i], "-C")) canonical = 1;
    else if(!strcmp(argv[i], "-t")) ++i;
    else if(!strcmp(argv[i], "-o")) out = argv[++i];
    else { first = i; break; }
  }
  if(K < 1 || K > 64 || n == 0 || !out) { fprintf(stderr, "usage: jf_oracle bc -m K -s N [-f FPR] [-C] -o OUT file...\n"); return 1; }
  const double LOG2 = 0.6931471805599453, LOG2_SQ = 0.4804530139182014;
  bf_m1.r = bf_m2.r = 64; bf_m1.c = bf_m2.c = 2 * K; bf_m1.identity = bf_m2.identity = 0;
  for(unsigned i = 0; i < 2 * K; ++i) bf_m1.col[i] = random_bits64();
  for(unsigned i = 0; i < 2 * K; ++i) bf_m2.col[i] = random_bits64();
  bf_m = n * (uint64_t)lrint(-log(fpr) / LOG2_SQ);
  bf_k = lrint(-log(fpr) / LOG2);
  size_t nb = bf_m / 5 + (bf_m % 5 != 0);
  bf_bits = calloc(nb + 1, 1);
  bc_build = 1;
  for(int i = first; i < argc; ++i) if(!count_file(argv[i])) return 1;
  FILE* f = fopen(out, "wb");
  if(!f) { fprintf(stderr, "Can't open output file '%s'\n", out); return 1; }
  char* js = malloc(1 << 20); size_t o = 0;
  o += sprintf(js + o, "{\"alignment\":8,\"canonical\":%s,\"cmdline\":[\"jf_oracle\"],\"exe_path\":\"jf_oracle\",\"format\":\"bloomcounter\",\"hostname\":\"hostname\",\"key_len\":%u,",
               canonical ? "true" : "false", 2 * K);
  for(int w = 1; w <= 2; ++w) {
    const matrix_t* M = w == 1 ? &bf_m1 : &bf_m2;
    o += sprintf(js + o, "\"matrix%d\":{\"c\":%u,\"columns\":[", w, M->c);
    for(unsigned i = 0; i < M->c; ++i) o += sprintf(js + o, "%s%llu", i ? "," : "", (unsigned long long)M->col[i]);
    o += sprintf(js + o, "],\"identity\":false,\"r\":64},");
  }
  o += sprintf(js + o, "\"nb_hashes\":%lu,\"pwd\":\".\",\"size\":%llu,\"time\":\"Thu Jan  1 00:00:00 1970\"}", bf_k, (unsigned long long)bf_m);
  size_t hlen = o, pad = (9 + o) % 8;
  if(pad) hlen += 8 - pad;
  fprintf(f, "%09zu", hlen);
  fwrite(js, 1, o, f);
  for(size_t i = o; i < hlen; ++i) fputc(0, f);
  fwrite(bf_bits, 1, nb, f);
  fclose(f);
  return 0;
}

/* ---- -Q / --min-quality: whole reads with their quality strings -----------------------------
 * count_main.cc:326-329 switches to mer_qual_counter = whole_sequence_parser (one record at a
 * time: whole_sequence_parser.hpp:137-152 read_fasta, :154-193 read_fastq) + mer_qual_iterator
 * (mer_qual_iterator.hpp:64-92: a base counts only if it is ACGT/acgt AND its quality character,
 * compared as a (signed) char, is >= min_qual; a FASTA record has no qualities = all pass; every
 * record starts with filled_ = 0). Unlike the default parser, std::getline keeps a line-end '\r'
 * inside the sequence (where it breaks the k-mer) and inside the quality string (where it is
 * counted), FASTQ sequence and quality may span several lines, and a malformed record is an
 * error rather than being skipped. */
static int min_qual = 0, use_qual = 0;
static long ws_getline(const unsigned char* d, long n, long p, long* b, long* e) {   /* std::getline: [b,e) without '\n' */
  *b = p; while(p < n && d[p] != '\n') ++p;
  *e = p; return p < n ? p + 1 : n;
}
static int count_file_qual(const char* path) {
  FILE* f = fopen(path, "rb");
  if(!f) { fprintf(stderr, "Can't open file '%s'\n", path); return 0; }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  unsigned char* d = malloc(n + 1);
  if(n && fread(d, 1, n, f) != (size_t)n) { perror("fread"); exit(1); }
  fclose(f);
  filled = 0; fwd = rev = 0;
  if(n == 0) { free(d); return 1; }
  if(d[0] != '>' && d[0] != '@') { fprintf(stderr, "Unsupported format\n"); exit(134); }
  long p = 0, b, e;
  if(d[0] == '>') {
    while(p < n) {
      ++p;                                                   /* '>' */
      p = ws_getline(d, n, p, &b, &e);                       /* header */
      filled = 0;
      while(p < n && d[p] != '>') {                          /* lines up to the next line starting with '>' */
        p = ws_getline(d, n, p, &b, &e);
        for(long i = b; i < e; ++i) feed_char(d[i]);
      }
    }
    free(d); filled = 0; return 1;
  }
  unsigned char* seq = malloc(n + 1); unsigned char* qual = malloc(n + 1);
  while(p < n) {
    long ns = 0, nq = 0;
    ++p;                                                     /* '@' (whatever the character is) */
    p = ws_getline(d, n, p, &b, &e);                         /* header */
    while(p < n && d[p] != '+') {                            /* sequence lines up to the line starting with '+' */
      p = ws_getline(d, n, p, &b, &e);
      memcpy(seq + ns, d + b, e - b); ns += e - b;
    }
    if(p >= n) { fprintf(stderr, "Truncated fastq file\n"); exit(1); }
    p = ws_getline(d, n, p, &b, &e);                         /* the '+' line */
    /* first quality line unconditionally when the read is not empty (a line starting with '+' is
     * cleared and then appended by the loop: same bytes); an empty read consumes one line unless it starts with '+' */
    if(ns == 0 && p < n && d[p] != '+') { p = ws_getline(d, n, p, &b, &e); memcpy(qual, d + b, e - b); nq = e - b; }
    while(nq < ns && p < n) {
      p = ws_getline(d, n, p, &b, &e);
      memcpy(qual + nq, d + b, e - b); nq += e - b;
    }
    if(nq != ns) { fprintf(stderr, "Invalid fastq file: wrong number of quals\n"); exit(1); }
    if(p < n && d[p] != '@') { fprintf(stderr, "Invalid fastq file: header missing\n"); exit(1); }
    filled = 0;
    for(long i = 0; i < ns; ++i) {
      if((signed char)qual[i] >= (signed char)min_qual) feed_char(seq[i]); else filled = 0;
    }
  }
  free(seq); free(qual); free(d);
  filled = 0;
  return 1;
}

typedef struct { uint64_t pos; u128 key; uint64_t count; uint64_t id; } rec_t;
static unsigned ceil_log2(uint64_t x) { unsigned l = 0; while(l < 64 && (1ULL << l) < x) ++l; return l; }
static uint64_t reprobe_off(unsigned i) { return i == 0 ? 1 : (uint64_t)i * (i + 1) / 2; }   /* lib/storage.cc:13-41 */

/* ---- the table, slot by slot ------------------------------------------------------------------
 * A logical model of large_hash::array (large_hash_array.hpp) and of hash_counter's growth
 * (hash_counter.hpp:91-115,178-238), driven in input order, i.e. what the reference does with one
 * thread.  A slot is EMPTY, the first entry of a KEY (key + the reprobe count it was claimed with +
 * val_len bits of value) or a LARGE continuation entry (the reprobe count back to where its chain
 * step started + lval_len = min(raw_key_len + val_len, 64) bits of value: offsets_key_value.hpp:91).
 *  - claim_key (:509-597): probe pos, pos+1, pos+3, ... up to the reprobe limit for an EMPTY slot
 *    or this key's own first entry; failure = "full".
 *  - add_rec_at (:674-694): add into the value field; a carry goes to a LARGE entry searched from
 *    (id + reprobes[0]) with the same probe sequence (claim_large_key :603-643: EMPTY, or LARGE
 *    with the same reprobe count); its carry continues from there.  A failure leaves what was
 *    stored in place and hands the rest (carry << bits stored) back to the caller.
 *  - full (handle_full_ary/double_size): a new array of twice the size -- or, once size == 4^k,
 *    of the same size with val_len + 1 -- with a new matrix draw (identity when size >= 4^k,
 *    :992-1002) and the OLD array's clipped reprobe limit (ary_->max_reprobe()), clipped again
 *    (reprobe_limit_t :29-39; 0 when key_len <= lsize :160).  Every first entry of the old array
 *    is re-added in slot order with its resolved value (resolve_val_rec :889-937; a failure of
 *    this add is ignored, as in the reference); then the failed operation is retried with the
 *    remaining value.
 * Slots are kept in a sparse map so that a 2^31-slot table costs memory only for what it holds. */
/* zeroed memory for the slot map: anonymous mmap with transparent huge pages requested (random
 * first touches of 4 KB pages are what this program would otherwise spend its time on) */
static void* big_zalloc(size_t bytes) {
  bytes = (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
  void* p = mmap(0, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if(p == MAP_FAILED) { perror("mmap"); exit(1); }
#ifdef MADV_HUGEPAGE
  madvise(p, bytes, MADV_HUGEPAGE);
#endif
  return p;
}
static void big_free(void* p, size_t bytes) { if(p) munmap(p, (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1)); }
enum { SLOT_EMPTY = 0, SLOT_KEY = 1, SLOT_LARGE = 2 };
enum { SIM_ADD, SIM_SET, SIM_UPDATE };
typedef struct { u128 key; uint64_t val; uint64_t id : 48, r : 12, state : 4; } slot_t;   /* 32 bytes: memory touched is what this program costs */
typedef struct {
  slot_t* a; size_t cap, n;
  unsigned lsize, kbits, limit, val_len, lval_len; uint64_t mask; matrix_t M;
} sim_t;
static size_t sim_home(const sim_t* T, uint64_t id) { return (size_t)((id * 0x9E3779B97F4A7C15ULL) >> 20) & (T->cap - 1); }
static slot_t* sim_get(const sim_t* T, uint64_t id) {
  for(size_t h = sim_home(T, id); ; h = (h + 1) & (T->cap - 1)) {
    if(T->a[h].state == SLOT_EMPTY) return 0;
    if(T->a[h].id == id) return &T->a[h];
  }
}
static slot_t* sim_put(sim_t* T, uint64_t id) {
  if((T->n + 1) * 10 > T->cap * 7) {
    slot_t* old = T->a; size_t ocap = T->cap;
    T->cap *= 2; T->a = big_zalloc(T->cap * sizeof(slot_t));
    for(size_t i = 0; i < ocap; ++i) if(old[i].state != SLOT_EMPTY) {
      size_t h = sim_home(T, old[i].id);
      while(T->a[h].state != SLOT_EMPTY) h = (h + 1) & (T->cap - 1);
      T->a[h] = old[i];
    }
    big_free(old, ocap * sizeof(slot_t));
  }
  size_t h = sim_home(T, id);
  while(T->a[h].state != SLOT_EMPTY) h = (h + 1) & (T->cap - 1);
  T->a[h].id = id; ++T->n;
  return &T->a[h];
}
static unsigned ceil_log2(uint64_t x);
static uint64_t reprobe_off(unsigned i);
static void sim_geometry(sim_t* T, unsigned lsize, unsigned limit_in) {
  T->lsize = lsize; T->mask = (1ULL << lsize) - 1;
  unsigned limit = T->kbits > lsize ? limit_in : 0;                                  /* :160 */
  while(limit >= 1 && reprobe_off(limit) >= (1ULL << lsize)) --limit;                /* :29-39 */
  T->limit = limit;
  unsigned raw = T->kbits > lsize ? T->kbits - lsize : 0;
  T->lval_len = raw + T->val_len < 64 ? raw + T->val_len : 64;
}
static void sim_init(sim_t* T, unsigned lsize, unsigned kbits, unsigned val_len, unsigned reprobes, const matrix_t* M) {
  memset(T, 0, sizeof(*T));
  T->cap = 1 << 16; T->a = big_zalloc(T->cap * sizeof(slot_t));
  T->kbits = kbits; T->val_len = val_len; T->M = *M;
  sim_geometry(T, lsize, reprobes);
}
/* add `v` to the field of `bits` bits of slot s; returns the carry */
static uint64_t sim_add_field(slot_t* s, uint64_t v, unsigned bits) {
  u128 nval = (u128)s->val + v;
  if(bits >= 64) { s->val = (uint64_t)nval; return (uint64_t)(nval >> 64); }
  s->val = (uint64_t)nval & ((1ULL << bits) - 1);
  return (uint64_t)(nval >> bits);
}
/* returns 1 when done, 0 when full (*rem = what is left to add), -1 when UPDATE found no key */
static int sim_try(sim_t* T, u128 key, uint64_t v, int op, uint64_t* rem) {
  const uint64_t pos = mat_times(&T->M, key) & T->mask;
  uint64_t cid = pos; unsigned r = 0; slot_t* s;
  for(;;) {                                                  /* claim_key / get_key_id */
    s = sim_get(T, cid);
    if(!s) {
      if(op == SIM_UPDATE) return -1;
      s = sim_put(T, cid); s->state = SLOT_KEY; s->key = key; s->r = r; s->val = 0;
      break;
    }
    if(s->state == SLOT_KEY && s->key == key) break;
    if(++r > T->limit) { if(op == SIM_UPDATE) return -1; *rem = v; return 0; }
    cid = (pos + reprobe_off(r)) & T->mask;
  }
  if(op == SIM_SET) return 1;
  uint64_t carry = sim_add_field(s, v, T->val_len);
  unsigned stored = T->val_len;
  while(carry) {                                             /* claim_large_key from (id + reprobes[0]) */
    const uint64_t start = (cid + reprobe_off(0)) & T->mask;
    uint64_t c = start; r = 0;
    for(;;) {
      s = sim_get(T, c);                                     /* (sim_put may move entries: re-fetched each time) */
      if(!s) { s = sim_put(T, c); s->state = SLOT_LARGE; s->r = r; s->val = 0; break; }
      if(s->state == SLOT_LARGE && s->r == r) break;
      if(++r > T->limit) { *rem = stored >= 64 ? 0 : carry << stored; return 0; }
      c = (start + reprobe_off(r)) & T->mask;
    }
    carry = sim_add_field(s, carry, T->lval_len);
    stored += T->lval_len;
    cid = c;
  }
  return 1;
}
/* value of the key whose first entry is at id: get_val_at_id + resolve_val_rec (:868-937) */
static uint64_t sim_resolve(const sim_t* T, uint64_t id) {
  const slot_t* s = sim_get(T, id);
  uint64_t val = s->val; unsigned shift = T->val_len;
  uint64_t start = (id + reprobe_off(0)) & T->mask;
  for(;;) {
    unsigned r = 0; uint64_t c = start; int found = 0;
    while(r <= T->limit) {
      s = sim_get(T, c);
      if(s && s->state == SLOT_LARGE) { if(s->r == r) { found = 1; break; } }
      else if(!s) break;
      c = (start + reprobe_off(++r)) & T->mask;
    }
    if(!found) return val;
    if(shift < 64) val += s->val << shift;
    shift += T->lval_len;
    start = (c + reprobe_off(0)) & T->mask;
  }
}
static int cmp_slot_id(const void* a, const void* b) { const slot_t* x = a; const slot_t* y = b; return x->id < y->id ? -1 : x->id > y->id; }
static void mat_draw(unsigned r, unsigned c, matrix_t* out);
static void sim_grow(sim_t* T) {                             /* hash_counter.hpp:200-238 */
  sim_t N = *T;
  N.cap = T->cap; N.a = big_zalloc(N.cap * sizeof(slot_t)); N.n = 0;
  const int can_double = T->kbits >= 64 || T->lsize < T->kbits;
  unsigned nl = T->lsize;
  if(can_double) ++nl; else ++N.val_len;
  if(T->kbits < 64 && nl >= T->kbits) { N.M.identity = 1; N.M.r = N.M.c = T->kbits; }   /* :992-1002: size >= 4^k */
  else mat_draw(nl, T->kbits, &N.M);
  sim_geometry(&N, nl, T->limit);
  /* re-add every first entry in slot order with its resolved value (eager_slice(0, 1)) */
  size_t nk = 0;
  slot_t* keys = malloc((T->n ? T->n : 1) * sizeof(slot_t));
  for(size_t i = 0; i < T->cap; ++i) if(T->a[i].state == SLOT_KEY) keys[nk++] = T->a[i];
  qsort(keys, nk, sizeof(slot_t), cmp_slot_id);
  for(size_t i = 0; i < nk; ++i) {
    uint64_t rem, v = sim_resolve(T, keys[i].id);
    if(v == 0) sim_try(&N, keys[i].key, 0, SIM_ADD, &rem);   /* add(key, 0): the key is claimed, nothing to add */
    else sim_try(&N, keys[i].key, v, SIM_ADD, &rem);         /* a failure here is ignored by the reference too */
  }
  free(keys); big_free(T->a, T->cap * sizeof(slot_t));
  *T = N;
}
static void sim_op(sim_t* T, u128 key, uint64_t v, int op) {
  uint64_t rem = 0;
  for(unsigned guard = 0; ; ++guard) {
    int rc = sim_try(T, key, v, op, &rem);
    if(rc != 0) return;
    if(guard > 200) { fprintf(stderr, "Hash full\n"); exit(1); }
    /* hash_counter::update_add (hash_counter.hpp:155-165) takes "carry_shift == v" for "key not
     * present" and gives up: a retried remainder that fails again with the same remainder (one
     * carry unit that still finds no continuation slot after the growth) is dropped -- the
     * reference loses those occurrences, and so does this model */
    if(op == SIM_UPDATE && rem == v) return;
    sim_grow(T);
    if(op == SIM_SET) continue;
    v = rem;                                                 /* hash_counter::add / update_add retry with carry_shift */
    if(v == 0) return;
  }
}

/* ---- order of the records in the dump ---------------------------------------------------------
 * sorted_dumper.hpp:57-101: the table is cut into chunks of B slots, B = the smallest multiple of
 * the packing block length (offsets_key_value.hpp:236-262) holding 5 * max(1, reprobes[limit])
 * records; a chunk yields the keys whose ORIGINAL position lies in it, read in slot order from its
 * start up to reprobes[limit] slots past its end (large_hash_iterator.hpp:150-198, wrapping), and
 * passes them through a min-heap on (position, key) of capacity reprobes[limit]
 * (mer_heap.hpp:26-30,57-100): fill, then pop one / push one.  With the default limit (capacity
 * 8001) that is a sort by (position, key); with a clipped limit of 1 or 2 (tables of a few
 * slots grown from -s < 8) the heap is too small to reorder equal positions and the output keeps
 * slot order there.  Chunk boundaries do not depend on the number of dumper threads. */
static unsigned bitsize(uint64_t x) { unsigned b = 0; while(x) { ++b; x >>= 1; } return b ? b : 1; }
static unsigned packing_block_len(unsigned key_field, unsigned val_len) {
  unsigned cword = 0, cboff = 0, n = 0;
  do {
    unsigned add = key_field + 1;                            /* + large bit */
    if(cboff + add <= 64) { cboff = (cboff + add) % 64; cword += cboff == 0; }
    else { add -= 63 - cboff; cword += 1 + add / 63; cboff = add % 63; cboff += cboff > 0; }
    cboff += val_len; cword += cboff / 64; cboff %= 64;
    ++n;
  } while(cboff != 0 && cboff < 62);
  (void)cword;
  return n;
}
static int cmp_rec_id(const void* a, const void* b) { const rec_t* x = a; const rec_t* y = b; return x->id < y->id ? -1 : x->id > y->id; }
static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
static int rec_less(const rec_t* x, const rec_t* y) { return x->pos != y->pos ? x->pos < y->pos : x->key < y->key; }
static void heap_push(const rec_t** h, size_t* n, const rec_t* r) {
  size_t i = (*n)++; h[i] = r;
  while(i && rec_less(h[i], h[(i - 1) / 2])) { const rec_t* t = h[i]; h[i] = h[(i - 1) / 2]; h[(i - 1) / 2] = t; i = (i - 1) / 2; }
}
static const rec_t* heap_pop(const rec_t** h, size_t* n) {
  const rec_t* top = h[0]; h[0] = h[--*n];
  for(size_t i = 0; ; ) {
    size_t l = 2 * i + 1, r = l + 1, m = i;
    if(l < *n && rec_less(h[l], h[m])) m = l;
    if(r < *n && rec_less(h[r], h[m])) m = r;
    if(m == i) break;
    const rec_t* t = h[i]; h[i] = h[m]; h[m] = t; i = m;
  }
  return top;
}
static rec_t* dump_order(rec_t* recs, size_t n, uint64_t size, unsigned limit, unsigned key_field, unsigned val_len) {
  const uint64_t max_off = reprobe_off(limit), cap = max_off;
  const unsigned blen = packing_block_len(key_field, val_len);
  const uint64_t want = 5 * (max_off > 1 ? max_off : 1);
  const uint64_t B = (want / blen + (want % blen != 0)) * blen;
  qsort(recs, n, sizeof(rec_t), cmp_rec_id);
  uint64_t* chunks = malloc((n ? n : 1) * sizeof(uint64_t));
  for(size_t i = 0; i < n; ++i) chunks[i] = recs[i].pos / B;
  qsort(chunks, n, sizeof(uint64_t), cmp_u64);
  rec_t* out = malloc((n ? n : 1) * sizeof(rec_t)); size_t no = 0;
  const rec_t** heap = malloc((cap + 1) * sizeof(rec_t*));
  const rec_t** stream = malloc((n ? n : 1) * sizeof(rec_t*));
  for(size_t ci = 0; ci < n; ++ci) {
    if(ci && chunks[ci] == chunks[ci - 1]) continue;
    const uint64_t start = chunks[ci] * B, end = start + B < size ? start + B : size;
    uint64_t mid = end - start + max_off; if(mid > size) mid = size;
    /* slots start .. start + mid - 1 (mod size) in that order, keys whose position is in [start, end) */
    size_t ns = 0, lo = 0, hi = n;
    while(lo < hi) { size_t m = (lo + hi) / 2; if(recs[m].id < start) lo = m + 1; else hi = m; }
    const uint64_t lim1 = start + mid < size ? start + mid : size;
    for(size_t i = lo; i < n && recs[i].id < lim1; ++i) if(recs[i].pos >= start && recs[i].pos < end) stream[ns++] = &recs[i];
    if(start + mid > size) {
      const uint64_t lim2 = start + mid - size;
      for(size_t i = 0; i < n && recs[i].id < lim2 && recs[i].id < start; ++i) if(recs[i].pos >= start && recs[i].pos < end) stream[ns++] = &recs[i];
    }
    size_t hn = 0, next = 0;
    while(next < ns && hn < cap) heap_push(heap, &hn, stream[next++]);
    while(hn) {
      out[no++] = *heap_pop(heap, &hn);
      if(next < ns) heap_push(heap, &hn, stream[next++]);
    }
  }
  free(heap); free(stream); free(chunks); free(recs);
  return out;
}

static uint64_t parse_size(const char* s) {
  char* end; double v = strtod(s, &end); uint64_t u = strtoull(s, &end, 0); (void)v;
  switch(*end) { case 'k': u *= 1000ULL; break; case 'M': u *= 1000000ULL; break; case 'G': u *= 1000000000ULL; break; case 'T': u *= 1000000000000ULL; break; }
  return u;
}

int main(int argc, char** argv) {
  rnd_seed(1);
  if(argc >= 4 && !strcmp(argv[1], "matrix")) {
    unsigned r = atoi(argv[2]), c = atoi(argv[3]), skip = argc > 4 ? atoi(argv[4]) : 0;
    matrix_t m;
    for(unsigned i = 0; i <= skip; ++i) mat_draw(r, c, &m);
    for(unsigned i = 0; i < c; ++i) printf("%llu\n", (unsigned long long)m.col[i]);
    return 0;
  }
  if(argc >= 2 && !strcmp(argv[1], "bc")) return bc_main(argc, argv);
  if(argc < 2 || strcmp(argv[1], "count")) { fprintf(stderr, "usage: jf_oracle count ... | matrix R C [SKIP]\n"); return 1; }
  uint64_t size = 0, low = 0, high = ~0ULL; unsigned val_len = 7, reprobes = 126, ocl = 4; const char* out = "mer_counts.jf";
  const char* if_files[64]; int n_if = 0;   /* --if: count only the k-mers of these files (count_main.cc:288-295) */
  int text = 0;   /* --text: text_dumper.hpp ("MER count" lines, format "text/sorted", no counter_len) */
  int first_file = argc;
  const char* bc_path = 0;
  uint64_t bf_size = 0; double bf_fp = 0.01;   /* count_main_cmdline.yaggo:44-49 */
  int min_quality = 0, quality_start = 64;   /* count_main_cmdline.yaggo:56-61 */
  for(int i = 2; i < argc; ++i) {
    if(!strcmp(argv[i], "-m")) K = atoi(argv[++i]);
    else if(!strcmp(argv[i], "-s")) size = parse_size(argv[++i]);
    else if(!strcmp(argv[i], "-C")) canonical = 1;
    else if(!strcmp(argv[i], "-c")) val_len = atoi(argv[++i]);
    else if(!strcmp(argv[i], "-p")) reprobes = atoi(argv[++i]);
    else if(!strcmp(argv[i], "--out-counter-len")) ocl = atoi(argv[++i]);
    else if(!strcmp(argv[i], "-L")) low = strtoull(argv[++i], 0, 0);
    else if(!strcmp(argv[i], "-U")) high = strtoull(argv[++i], 0, 0);
    else if(!strcmp(argv[i], "-o")) out = argv[++i];
    else if(!strcmp(argv[i], "-t")) ++i;
    else if(!strcmp(argv[i], "--text")) text = 1;
    else if(!strcmp(argv[i], "-Q") || !strcmp(argv[i], "--min-qual-char")) {   /* count_main.cc:234-244 */
      const char* a = argv[++i];
      if(strlen(a) != 1 || a[0] < '!' || a[0] > '~') { fprintf(stderr, "[-Q, --min-qual-char] must be one printable character\n"); return 1; }
      min_qual = a[0]; use_qual = 1;
    }
    else if(!strcmp(argv[i], "--bf-size")) bf_size = parse_size(argv[++i]);
    else if(!strcmp(argv[i], "--bf-fp")) bf_fp = atof(argv[++i]);
    else if(!strcmp(argv[i], "--bc")) bc_path = argv[++i];
    else if(!strcmp(argv[i], "--min-quality")) { min_quality = atoi(argv[++i]); use_qual = 2; }
    else if(!strcmp(argv[i], "--quality-start")) quality_start = atoi(argv[++i]);
    else if(!strcmp(argv[i], "--if")) { if(n_if < 64) if_files[n_if++] = argv[++i]; else ++i; }
    else { first_file = i; break; }
  }
  if(use_qual == 2) {   /* count_main.cc:245-256 */
    min_qual = quality_start + min_quality;
    if(quality_start < '!' || quality_start > '~' || min_qual < '!' || min_qual > '~') { fprintf(stderr, "Quality out of range\n"); return 1; }
  }
  if(K < 1 || K > 64 || size == 0) { fprintf(stderr, "need -m (1..64) and -s\n"); return 1; }
  const unsigned kbits = 2 * K;

  /* table + first matrix are created BEFORE the input is read: count_main.cc:275 */
  uint64_t key_space = kbits >= 64 ? ~0ULL / 2 : (1ULL << kbits);
  unsigned lsize = ceil_log2(size < key_space ? size : key_space);
  matrix_t M;
  if(lsize == 0) { fprintf(stderr, "Invalid matrix size\n"); return 134; }
  if(lsize > 47) { fprintf(stderr, "jf_oracle: tables of more than 2^47 slots are not modelled\n"); return 1; }   /* RectangularBinaryMatrix(0, c) throws: the reference aborts */
  if(size < key_space) mat_draw(lsize, kbits, &M); else { M.identity = 1; M.r = M.c = kbits; }

  /* The table is simulated slot by slot in input order (= a reference run with -t 1): see sim_*
   * above.  --if: first pass PRIME over the --if files (array::set: the key enters with count 0),
   * second pass UPDATE (update_add: only keys already present are incremented),
   * count_main.cc:288-295,152-184. */
  sim_t T;
  sim_init(&T, lsize, kbits, val_len, reprobes, &M);
  if(n_if) {
    for(int i = 0; i < n_if; ++i) if(!count_file(if_files[i])) return 1;
    for(size_t i = 0; i < n_mers; ++i) sim_op(&T, mers[i], 0, SIM_SET);
    n_mers = 0;
  }
  if(bc_path) bc_load(bc_path);
  if(bf_size) bf_setup(bf_fp, bf_size);
  for(int i = first_file; i < argc; ++i) if(!(use_qual ? count_file_qual(argv[i]) : count_file(argv[i]))) return 1;
  for(size_t i = 0; i < n_mers; ++i) sim_op(&T, mers[i], 1, n_if ? SIM_UPDATE : SIM_ADD);

  /* what the dumper sees: every first entry of a key with the sum of its chain */
  size_t n_rec = 0;
  rec_t* recs = malloc((T.n ? T.n : 1) * sizeof(rec_t));
  for(size_t i = 0; i < T.cap; ++i) {
    if(T.a[i].state != SLOT_KEY) continue;
    recs[n_rec].key = T.a[i].key; recs[n_rec].count = sim_resolve(&T, T.a[i].id);
    recs[n_rec].pos = mat_times(&T.M, T.a[i].key) & T.mask; recs[n_rec].id = T.a[i].id; ++n_rec;
  }
  M = T.M; lsize = T.lsize; val_len = T.val_len;
  const unsigned limit = T.limit;
  const uint64_t tsize = 1ULL << lsize;
  recs = dump_order(recs, n_rec, tsize, limit, (kbits > lsize ? kbits - lsize : 0) + bitsize(limit + 1), val_len);

  /* header: generic_file_header.hpp:88-111, file_header.hpp:26-108 (keys sorted, terse JSON) */
  FILE* f = fopen(out, "wb");
  if(!f) { fprintf(stderr, "Can't open output file '%s'\n", out); return 1; }
  char* js = malloc(1 << 20); size_t o = 0;
  o += sprintf(js + o, "{\"alignment\":8,\"canonical\":%s,\"cmdline\":[\"jf_oracle\"],", canonical ? "true" : "false");
  if(!text) o += sprintf(js + o, "\"counter_len\":%u,", ocl);
  o += sprintf(js + o, "\"exe_path\":\"jf_oracle\",\"format\":\"%s\",\"hostname\":\"hostname\",\"key_len\":%u,\"matrix1\":{\"c\":%u,",
               text ? "text/sorted" : "binary/sorted", kbits, M.c);
  if(!M.identity) {
    o += sprintf(js + o, "\"columns\":[");
    for(unsigned i = 0; i < M.c; ++i) o += sprintf(js + o, "%s%llu", i ? "," : "", (unsigned long long)M.col[i]);
    o += sprintf(js + o, "],\"identity\":false,");
  } else o += sprintf(js + o, "\"identity\":true,");
  o += sprintf(js + o, "\"r\":%u},\"max_reprobe\":%u,\"pwd\":\".\",\"reprobes\":[", M.r, limit);
  for(unsigned i = 0; i <= limit; ++i) o += sprintf(js + o, "%s%llu", i ? "," : "", (unsigned long long)reprobe_off(i));
  o += sprintf(js + o, "],\"size\":%llu,\"time\":\"Thu Jan  1 00:00:00 1970\",\"val_len\":%u}", (unsigned long long)tsize, val_len);
  size_t hlen = o, pad = (9 + o) % 8;
  if(pad) hlen += 8 - pad;
  fprintf(f, "%09zu", hlen);
  fwrite(js, 1, o, f);
  for(size_t i = o; i < hlen; ++i) fputc(0, f);
  /* records: binary_dumper.hpp:36-40 */
  unsigned kb = (kbits + 7) / 8;
  uint64_t maxv = ocl >= 8 ? ~0ULL : ((1ULL << (8 * ocl)) - 1);
  for(size_t i = 0; i < n_rec; ++i) {
    if(recs[i].count < low || recs[i].count > high) continue;
    if(text) {
      char mer[80];
      for(unsigned j = 0; j < K; ++j) mer[j] = "ACGT"[(unsigned)(recs[i].key >> (2 * (K - 1 - j))) & 3];
      mer[K] = 0;
      fprintf(f, "%s %llu\n", mer, (unsigned long long)recs[i].count);
      continue;
    }
    fwrite(&recs[i].key, 1, kb, f);
    uint64_t v = recs[i].count < maxv ? recs[i].count : maxv;
    fwrite(&v, 1, ocl, f);
  }
  fclose(f);
  fprintf(stderr, "jf_oracle: %zu k-mers, %zu distinct, size %llu\n", n_mers, n_rec, (unsigned long long)tsize);
  return 0;
}
