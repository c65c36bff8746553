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

}  // namespace eu
