/* TEST INFRASTRUCTURE -- CPU restatement of the SYNTHETIC GRAPH GENERATOR of euler_b200
 * (euler_b200/csrc/graph.cu: k_rmat_edges / k_rmat_fill / k_build_cum / k_fill_feat), so that
 *   - bench.py's reference arm builds its input graph on the host without loading libeuler_b200.so,
 *   - bench.py's parity gate gets the feature rows of the nodes a batch touched from an independent source,
 *   - tests/ can check the device generator against a second implementation.
 * This is not reference code: alibaba/euler has no graph generator; the inputs are SURVEY.md section 8(d)'s
 * "G-RMAT" (a,b,c,d = 0.57,0.19,0.19,0.05; ids 1..n; weight = 1 + (hash(src,dst) % 100) / 10; features ~ U(-1,1)).
 * Same integer hashes and the same IEEE operations as the device code, hence bit-identical outputs:
 *   u = (splitmix >> 11) * 2^-53 is exact, the quadrant compares are plain double compares,
 *   weight = 1.0f + (float)(h % 100) / 10.0f (two correctly rounded f32 ops), the per-node cumulative weights are a
 *   left-to-right f32 sum (Node::Init, euler/core/graph/node.cc:46-70), feature = (double)(h >> 11) * 2^-52 - 1.0 (the
 *   product is exact, so an FMA on the device rounds the same as mul + sub here).
 * Built with -ffp-contract=off like the rest of the oracle.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}
static inline uint64_t splitmix(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

typedef struct {
  int64_t n_nodes, n_edges;
  int scale, T;
  double a, b, c;
  uint64_t seed;
} rmat_par;

/* edge e -> (0-based src, 0-based dst, edge type): graph.cu k_rmat_edges */
static inline void rmat_edge(const rmat_par* p, int64_t e, uint64_t* src_o, uint64_t* dst_o, uint64_t* et_o) {
  uint64_t s = mix64(p->seed ^ (uint64_t)e * 0xD6E8FEB86659FD93ULL);
  uint64_t src = 0, dst = 0;
  const double ab = p->a + p->b, abc = p->a + p->b + p->c;
  for (int l = 0; l < p->scale; ++l) {
    double u = (double)(splitmix(&s) >> 11) * (1.0 / 9007199254740992.0);
    int q = u < p->a ? 0 : (u < ab ? 1 : (u < abc ? 2 : 3));
    src = (src << 1) | (uint64_t)(q >> 1);
    dst = (dst << 1) | (uint64_t)(q & 1);
  }
  *src_o = mix64(src + 0x51ED27) % (uint64_t)p->n_nodes;
  *dst_o = mix64(dst + 0x51ED27) % (uint64_t)p->n_nodes;
  *et_o = p->T > 1 ? mix64(p->seed * 0x2545F4914F6CDD1DULL + (uint64_t)e) % (uint64_t)p->T : 0ull;
}

typedef struct {
  const rmat_par* p;
  int tid, nthreads, phase;
  int64_t* deg;          /* [n*T+1] counts, then cursors */
  uint64_t* nbr;         /* [E] */
  const int64_t* ptr;    /* [n*T+1] */
  float* cum_w; float* grp_cum;
} rmat_job;

static int cmp_u64(const void* x, const void* y) {
  uint64_t a = *(const uint64_t*)x, b = *(const uint64_t*)y;
  return a < b ? -1 : (a > b ? 1 : 0);
}

static void* rmat_worker(void* arg) {
  rmat_job* j = (rmat_job*)arg;
  const rmat_par* p = j->p;
  if (j->phase == 0 || j->phase == 1) {
    const int64_t per = (p->n_edges + j->nthreads - 1) / j->nthreads;
    const int64_t b = (int64_t)j->tid * per, e = b + per < p->n_edges ? b + per : p->n_edges;
    for (int64_t k = b; k < e; ++k) {
      uint64_t src, dst, et;
      rmat_edge(p, k, &src, &dst, &et);
      const int64_t g = (int64_t)(src * (uint64_t)p->T + et);
      if (j->phase == 0) {
        __atomic_fetch_add(&j->deg[g], 1, __ATOMIC_RELAXED);
      } else {
        const int64_t pos = __atomic_fetch_add(&j->deg[g], 1, __ATOMIC_RELAXED);
        j->nbr[pos] = dst + 1;   /* ids are 1..n */
      }
    }
  } else {
    /* per row: sort every group by neighbor id (the device sorts (group, dst) keys), weights, cumulative sums */
    const int64_t n = p->n_nodes, T = p->T;
    const int64_t per = (n + j->nthreads - 1) / j->nthreads;
    const int64_t b = (int64_t)j->tid * per, e = b + per < n ? b + per : n;
    for (int64_t r = b; r < e; ++r) {
      float sum_weight = 0.f, cwc = 0.f;
      for (int64_t t = 0; t < T; ++t) {
        const int64_t lo = j->ptr[r * T + t], hi = j->ptr[r * T + t + 1];
        if (hi - lo > 1) qsort(j->nbr + lo, (size_t)(hi - lo), sizeof(uint64_t), cmp_u64);
        float type_weight = 0.f;
        for (int64_t k = lo; k < hi; ++k) {
          const uint64_t dst = j->nbr[k] - 1;
          const uint64_t h = mix64((uint64_t)r * 0x9E3779B97F4A7C15ULL ^ dst);
          const float w = 1.0f + (float)(h % 100ull) / 10.0f;
          sum_weight = sum_weight + w;
          type_weight = type_weight + w;
          j->cum_w[k] = sum_weight;
        }
        cwc = cwc + type_weight;
        if (j->grp_cum) j->grp_cum[r * T + t] = cwc;
      }
    }
  }
  return NULL;
}

static void run_phase(rmat_job* jobs, int nthreads, int phase) {
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) { jobs[t].phase = phase; pthread_create(&th[t], NULL, rmat_worker, &jobs[t]); }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
}

/* The whole (unsharded) graph of eu_graph_create_rmat / _hetero on the host.
 * Caller-allocated outputs: ids u64[n] (1..n), node_type i32[n] (id % NT), node_w f32[n] (1.0), grp_ptr i64[n*T+1],
 * nbr u64[E], cum_w f32[E], grp_cum f32[n*T] (may be NULL when T == 1).  Returns 0, or -1 on a bad argument / OOM. */
int eo_rmat_csr(int64_t n_nodes, int64_t n_edges, double a, double b, double c, uint64_t seed, int32_t T, int32_t NT,
                int32_t nthreads, uint64_t* ids, int32_t* node_type, float* node_w, int64_t* grp_ptr, uint64_t* nbr,
                float* cum_w, float* grp_cum) {
  if (n_nodes <= 0 || n_edges < 0 || T < 1 || NT < 1 || nthreads < 1 || !grp_ptr) return -1;
  rmat_par p;
  p.n_nodes = n_nodes; p.n_edges = n_edges; p.T = T; p.a = a; p.b = b; p.c = c; p.seed = seed;
  p.scale = 1;
  while (((int64_t)1 << p.scale) < n_nodes) ++p.scale;
  const int64_t ng = n_nodes * T;
  int64_t* deg = (int64_t*)calloc((size_t)ng + 1, sizeof(int64_t));
  rmat_job* jobs = (rmat_job*)calloc((size_t)nthreads, sizeof(rmat_job));
  if (!deg || !jobs) { free(deg); free(jobs); return -1; }
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].p = &p; jobs[t].tid = t; jobs[t].nthreads = nthreads; jobs[t].deg = deg; jobs[t].nbr = nbr;
    jobs[t].ptr = grp_ptr; jobs[t].cum_w = cum_w; jobs[t].grp_cum = grp_cum;
  }
  run_phase(jobs, nthreads, 0);
  int64_t run = 0;
  for (int64_t g = 0; g < ng; ++g) { grp_ptr[g] = run; run += deg[g]; deg[g] = grp_ptr[g]; }
  grp_ptr[ng] = run;
  run_phase(jobs, nthreads, 1);
  run_phase(jobs, nthreads, 2);
  for (int64_t r = 0; r < n_nodes; ++r) {
    if (ids) ids[r] = (uint64_t)r + 1;
    if (node_type) node_type[r] = (int32_t)(((uint64_t)r + 1) % (uint64_t)NT);
    if (node_w) node_w[r] = 1.0f;
  }
  free(deg); free(jobs);
  return 0;
}

/* Dense feature row of node `id` (graph.cu k_fill_feat): out[d] for d < dim; zeros for ids outside 1..n_nodes
 * (tf_euler/kernels/get_dense_feature_op.cc:66-75 zero-fills missing nodes). */
void eo_rmat_feat_rows(const int64_t* ids, int64_t M, int64_t n_nodes, int32_t dim, uint64_t feat_seed, float* out) {
  for (int64_t i = 0; i < M; ++i) {
    const uint64_t id = (uint64_t)ids[i];
    float* o = out + i * (int64_t)dim;
    if (id < 1 || id > (uint64_t)n_nodes) { memset(o, 0, sizeof(float) * (size_t)dim); continue; }
    for (int32_t d = 0; d < dim; ++d) {
      const uint64_t gi = (id - 1) * (uint64_t)dim + (uint64_t)d;
      const uint64_t h = mix64(feat_seed ^ (gi * 0x9E3779B97F4A7C15ULL));
      o[d] = (float)((double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0);
    }
  }
}

typedef struct { int64_t b, e, n_nodes; int32_t dim; uint64_t seed; float* out; } feat_job;
static void* feat_worker(void* arg) {
  feat_job* j = (feat_job*)arg;
  for (int64_t r = j->b; r < j->e; ++r) {
    const int64_t id = r + 1;
    eo_rmat_feat_rows(&id, 1, j->n_nodes, j->dim, j->seed, j->out + r * (int64_t)j->dim);
  }
  return NULL;
}

/* the whole feature table f32[n_nodes, dim], multi-threaded */
void eo_rmat_feat_full(int64_t n_nodes, int32_t dim, uint64_t feat_seed, int32_t nthreads, float* out) {
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  feat_job* jobs = (feat_job*)calloc((size_t)nthreads, sizeof(feat_job));
  const int64_t per = (n_nodes + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].b = (int64_t)t * per; jobs[t].e = jobs[t].b + per < n_nodes ? jobs[t].b + per : n_nodes;
    if (jobs[t].b > n_nodes) jobs[t].b = n_nodes;
    jobs[t].n_nodes = n_nodes; jobs[t].dim = dim; jobs[t].seed = feat_seed; jobs[t].out = out;
    pthread_create(&th[t], NULL, feat_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
