/*
 * Plain-C restatement of the retrieval hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * (Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.)
 *
 * Same arithmetic as oracle/oracle_np.py, written as strictly sequential float64 loops so that
 * the result does not depend on a BLAS library's blocking:
 *   oracle_cosine_pair     cosine_similarity            33_run_all_experiments.py:76-77
 *   oracle_cosine_topk     the same formula for every (query, row) pair + (score desc, id asc) top-k
 *   oracle_l2_normalize    normalise-on-upsert of a COSINE collection (third party, parity unpinned)
 *   oracle_segment_mean    compute_average per class     32_create_delegate_vector.py:9-10
 *                          + the stored (normalised, float32) form of the mean  :41-42
 *   oracle_distance_topk   the other distances of the collection menu (util/qdrant_manager.py:61-66): DOT, EUCLID,
 *                          MANHATTAN on the stored values, ordered by (key desc, id asc), key = q.g / -d^2 / -d
 *                          (third-party scoring semantics, parity unpinned)
 * Pinned against tests/golden/ (vectors produced by the reference's own functions).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

double oracle_cosine_pair(const double* a, const double* b, int64_t dim) {
  double dot = 0.0, aa = 0.0, bb = 0.0;
  for (int64_t i = 0; i < dim; ++i) {
    dot += a[i] * b[i];
    aa += a[i] * a[i];
    bb += b[i] * b[i];
  }
  return dot / (sqrt(aa) * sqrt(bb));
}

void oracle_l2_normalize(const float* x, int64_t n, int64_t dim, float* out, float* norms) {
  for (int64_t r = 0; r < n; ++r) {
    double s = 0.0;
    for (int64_t i = 0; i < dim; ++i) s += (double)x[r * dim + i] * (double)x[r * dim + i];
    const double inv = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
    for (int64_t i = 0; i < dim; ++i) out[r * dim + i] = (float)((double)x[r * dim + i] * inv);
    if (norms) norms[r] = (float)sqrt(s);
  }
}

/* stored rows (float32, already in stored form) -> per-class normalised mean */
void oracle_segment_mean(const float* stored, int64_t dim, const int64_t* row_idx, const int64_t* offsets,
                         int64_t n_classes, float* out) {
  double* acc = (double*)malloc(sizeof(double) * (size_t)dim);
  float* mean = (float*)malloc(sizeof(float) * (size_t)dim);
  for (int64_t c = 0; c < n_classes; ++c) {
    const int64_t a = offsets[c], b = offsets[c + 1];
    memset(acc, 0, sizeof(double) * (size_t)dim);
    for (int64_t i = a; i < b; ++i) {
      const int64_t r = row_idx ? row_idx[i] : i;
      for (int64_t d = 0; d < dim; ++d) acc[d] += (double)stored[r * dim + d];
    }
    for (int64_t d = 0; d < dim; ++d) mean[d] = b > a ? (float)(acc[d] / (double)(b - a)) : 0.0f;
    oracle_l2_normalize(mean, 1, dim, out + c * dim, NULL);
  }
  free(acc);
  free(mean);
}

/* insertion into a (score desc, id asc) ordered list of length k */
static void topk_insert(double* s, int64_t* ids, int64_t k, double score, int64_t id) {
  if (!(score > s[k - 1] || (score == s[k - 1] && (ids[k - 1] < 0 || id < ids[k - 1])))) return;
  int64_t p = k - 1;
  while (p > 0 && (score > s[p - 1] || (score == s[p - 1] && (ids[p - 1] < 0 || id < ids[p - 1])))) {
    s[p] = s[p - 1];
    ids[p] = ids[p - 1];
    --p;
  }
  s[p] = score;
  ids[p] = id;
}

/* row_allowed: optional byte mask (1 = row may be returned) */
void oracle_cosine_topk(const float* q, int64_t Q, const float* g, int64_t N, int64_t dim, int64_t k,
                        const uint8_t* row_allowed, double* out_scores, int64_t* out_ids) {
  double* gn = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  for (int64_t r = 0; r < N; ++r) {
    double s = 0.0;
    for (int64_t d = 0; d < dim; ++d) s += (double)g[r * dim + d] * (double)g[r * dim + d];
    gn[r] = sqrt(s);
  }
  for (int64_t i = 0; i < Q; ++i) {
    double* s = out_scores + i * k;
    int64_t* ids = out_ids + i * k;
    for (int64_t j = 0; j < k; ++j) { s[j] = -INFINITY; ids[j] = -1; }
    double qq = 0.0;
    for (int64_t d = 0; d < dim; ++d) qq += (double)q[i * dim + d] * (double)q[i * dim + d];
    const double qn = sqrt(qq);
    for (int64_t r = 0; r < N; ++r) {
      if (row_allowed && !row_allowed[r]) continue;
      double dot = 0.0;
      for (int64_t d = 0; d < dim; ++d) dot += (double)q[i * dim + d] * (double)g[r * dim + d];
      const double den = qn * gn[r];
      const double sc = den > 0.0 ? dot / den : 0.0;
      topk_insert(s, ids, k, sc, r);
    }
  }
  free(gn);
}

/* metric: 1 = DOT (key = q.g), 2 = EUCLID (key = -sum (q-g)^2), 3 = MANHATTAN (key = -sum |q-g|).
 * out_keys: the ordering keys, descending; out_ids: rows, -1 where fewer than k rows are allowed. */
void oracle_distance_topk(const float* q, int64_t Q, const float* g, int64_t N, int64_t dim, int64_t k, int metric,
                          const uint8_t* row_allowed, double* out_keys, int64_t* out_ids) {
  for (int64_t i = 0; i < Q; ++i) {
    double* s = out_keys + i * k;
    int64_t* ids = out_ids + i * k;
    for (int64_t j = 0; j < k; ++j) { s[j] = -INFINITY; ids[j] = -1; }
    for (int64_t r = 0; r < N; ++r) {
      if (row_allowed && !row_allowed[r]) continue;
      double acc = 0.0;
      for (int64_t d = 0; d < dim; ++d) {
        const double a = (double)q[i * dim + d], b = (double)g[r * dim + d];
        if (metric == 1) acc += a * b;
        else if (metric == 2) acc += (a - b) * (a - b);
        else acc += fabs(a - b);
      }
      topk_insert(s, ids, k, metric == 1 ? acc : -acc, r);
    }
  }
}
