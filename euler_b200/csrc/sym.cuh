// Symmetric-region layout shared by the exchange kernels (p2p.cu) and the bucket kernels that push into it (shard.cu).
#pragma once
#include "common.cuh"

struct eu_ctx;

namespace eu {

static constexpr int kSymMaxRanks = 16;

struct SymHeader {
  unsigned int flagA[kSymMaxRanks];   // [src]   epoch of the last inbox segment pushed by src
  unsigned int flagB[kSymMaxRanks];   // [owner] epoch of the last reply written by owner
  int in_cnt[kSymMaxRanks];           // [src]   seeds in src's segment
  int in_total[kSymMaxRanks];         // [src]   ids src bucketed in this exchange (before any drop): every rank must issue the
                                      //         same exchange shape, the owner verifies it against its own
  unsigned int epoch;                 // local exchange counter
  unsigned int done;                  // last-block ticket
  int error;                          // 1 = a wait timed out
  int pad;
};

struct SymLayout {
  int64_t cap;          // inbox slots per source
  int64_t max_out;      // rows * count slots of the sample outputs
  int64_t max_rows_f;   // rows of the feature output
  int32_t max_dim;
  long long timeout_cycles;   // bound of every flag wait (EU_SYM_TIMEOUT_S, default 30 s)
  int64_t off_inbox_ids, off_inbox_src, off_eng, off_ids, off_w, off_t, off_rows, off_flags, bytes;
};

struct SymPeers {
  char* base[kSymMaxRanks];
};

__device__ __forceinline__ SymHeader* hdr_of(char* base) { return reinterpret_cast<SymHeader*>(base); }

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// shard.cu: stable bucket by owner whose placement pass writes straight into the owners' inboxes and raises flagA
// pb_tab: DEVICE array of the N peer base pointers (indexing a by-value table with a runtime index costs every thread a
// local-memory copy of it)
int bucket_push(struct ::eu_ctx* c, const int64_t* ids, int64_t rows, int P, int N, int self, bool drop_placeholders, int64_t* counts,
                int64_t* offsets, char* const* pb_tab, const SymLayout& lay, const char* label);

__device__ __forceinline__ int ld_volatile_i32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// source rank of position p of a batch's compact owner input: bo[0..N] are the batch's per-source offsets
__device__ __forceinline__ int src_of(const int32_t* __restrict__ bo /* [N+1] */, int N, int32_t p) {
  int s = 0;
  while (s + 1 < N && p >= bo[s + 1]) ++s;
  return s;
}

// Owner-side sampleNB whose results go STRAIGHT into the requesters' output arrays (k_prepare / k_sample of sample.cu write
// over NVLink; no reply pass, no intermediate result arrays).  Row p of batch g of the owner's compact input came from source
// rank s = src_of(...) and sits at position src[...] of that rank's request: its `count` results land at [position][0..count).
struct SymRedirect {
  int on, me, N, nb, want_packed, pad_;
  char* const* pb_tab;          // device table of the peers' region bases
  int64_t cap, off_inbox_src, off_eng, off_ids, off_w, off_t;
  const int32_t* seg_lo;        // [N][nb+1] start of batch g inside source s's inbox segment
  const int32_t* boff;          // [nb][N+1] per-source offsets of batch g's compact input
};
struct RowOut {
  unsigned long long* eng;      // engine ids (next frontier), may be null
  long long* ids;               // TF-packed ids / weights / types, may be null
  float* w;
  int32_t* t;
  int64_t ob;                   // index of the row's first slot
};
__device__ __forceinline__ RowOut sym_row_out(const SymRedirect& rd, int g, int32_t p, int32_t count) {
  const int32_t* bo = rd.boff + g * (rd.N + 1);
  const int s = src_of(bo, rd.N, p);
  const int32_t* src = reinterpret_cast<const int32_t*>(rd.pb_tab[rd.me] + rd.off_inbox_src);
  const int64_t pos = src[(int64_t)s * rd.cap + rd.seg_lo[s * (rd.nb + 1) + g] + (p - bo[s])];
  char* pb = rd.pb_tab[s];
  RowOut o;
  o.eng = reinterpret_cast<unsigned long long*>(pb + rd.off_eng);
  o.ids = rd.want_packed ? reinterpret_cast<long long*>(pb + rd.off_ids) : nullptr;
  o.w = reinterpret_cast<float*>(pb + rd.off_w);
  o.t = reinterpret_cast<int32_t*>(pb + rd.off_t);
  o.ob = pos * count;
  return o;
}

}  // namespace eu
