// TEST INFRASTRUCTURE, not product code: compiles the __host__ __device__ operator arithmetic of
// clip_glass_b200/csrc/ga_ops.cuh with g++ so that the CPU suite can check the logic the CUDA kernels wrap (ga.cu)
// against clip_glass_b200/ga.py without a GPU.  Each function loops exactly like the kernel of the same name.
// Built on the fly by tests/test_ga_native.py; nothing in clip_glass_b200/ loads it.
#include <cstring>
#include <vector>

#include "../../clip_glass_b200/csrc/ga_ops.cuh"

using namespace glass_ga;

extern "C" {

struct host_ga_params {
  double sbx_eta, sbx_prob, sbx_prob_var, pm_eta, pm_prob;
  int32_t n_var, integer;
};

void ga_host_philox(uint32_t* ctr, uint32_t k0, uint32_t k1) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, k0, k1);
  memcpy(ctr, c, sizeof c);
}

void ga_host_uniform(uint64_t seed, uint64_t offset, double* out, int64_t n) {
  for (int64_t q = 0; q < (n + 1) / 2; ++q) {
    double u0, u1;
    uniform_pair(seed, offset + (uint64_t)q, u0, u1);
    out[2 * q] = u0;
    if (2 * q + 1 < n) out[2 * q + 1] = u1;
  }
}

void ga_host_permutations(const double* keys, int n, int n_perm, int32_t* out) {
  for (int p = 0; p < n_perm; ++p)
    for (int j = 0; j < n; ++j) out[(size_t)p * n + stable_rank(keys + (size_t)p * n, n, j)] = j;
}

void ga_host_tournament(const int32_t* pairs, const int32_t* rank, const double* crowd, int n_select, int32_t* sel) {
  for (int t = 0; t < n_select; ++t) sel[t] = tournament_element(pairs, rank, crowd, t);
}

void ga_host_offspring(const host_ga_params* g, const double* X, const int32_t* parents, const double* bounds,
                       const double* rnd, int M, double* out) {
  OpParams p;
  p.sbx_eta = g->sbx_eta; p.sbx_prob = g->sbx_prob; p.sbx_prob_var = g->sbx_prob_var;
  p.pm_eta = g->pm_eta; p.pm_prob = g->pm_prob; p.n_var = g->n_var; p.integer = g->integer;
  for (int m = 0; m < M; ++m)
    for (int v = 0; v < p.n_var; ++v) offspring_element(p, X, parents, bounds, rnd, M, m, v, out);
}

// ga_dup_flags_kernel + ga_dup_dest_kernel + ga_append_kernel
void ga_host_dedup_append(const double* cand, int n_c, const double* X, int n_x, double* off, int n_off,
                          int32_t* n_have, int V, double eps, int eliminate, float* z32, int32_t* flags,
                          int32_t* dest) {
  const int have = *n_have;
  for (int j = 0; j < n_c; ++j) {
    int dup = 0;
    const int rows = n_x + have + j;
    for (int r = 0; r < rows && !dup && eliminate; ++r) {
      const double* row = r < n_x ? X + (size_t)r * V
                                  : (r < n_x + have ? off + (size_t)(r - n_x) * V : cand + (size_t)(r - n_x - have) * V);
      dup = rows_equal(row, cand + (size_t)j * V, V, eps) ? 1 : 0;
    }
    flags[j] = dup;
  }
  int kept = 0;
  for (int j = 0; j < n_c; ++j) {
    int d = -1;
    if (!flags[j]) {
      int before = 0;
      for (int i = 0; i < j; ++i) before += flags[i] ? 0 : 1;
      if (have + before < n_off) { d = have + before; ++kept; }
    }
    dest[j] = d;
  }
  *n_have = have + kept;
  for (int j = 0; j < n_c; ++j) {
    if (dest[j] < 0) continue;
    for (int v = 0; v < V; ++v) {
      off[(size_t)dest[j] * V + v] = cand[(size_t)j * V + v];
      if (z32) z32[(size_t)dest[j] * V + v] = (float)cand[(size_t)j * V + v];
    }
  }
}

void ga_host_survive(const float* F, int ld, int n, int n_obj, int n_survive, int nsga2, int32_t* idx, int32_t* rank,
                     double* crowd) {
  std::vector<int> ndom(n), front(n), cur(n), pos(n);
  std::vector<double> cd(n), sorted(n);
  int fsize = 0;
  SurviveState s;
  s.F = F; s.ld = ld; s.n = n; s.n_obj = n_obj; s.n_survive = n_survive; s.nsga2 = nsga2;
  s.ndom = ndom.data(); s.front = front.data(); s.cur = cur.data(); s.pos = pos.data();
  s.cd = cd.data(); s.sorted = sorted.data(); s.fsize = &fsize;
  s.out_idx = idx; s.out_rank = rank; s.out_crowd = crowd;
  survive_body(s);
}

// ga_pad_kernel, ga_cast_kernel, ga_gather_kernel
void ga_host_pad(double* off, int n_off, const int32_t* n_have, int V, float* z32) {
  const int have = *n_have;
  if (have <= 0 || have >= n_off) return;
  for (size_t e = 0; e < (size_t)(n_off - have) * V; ++e) {
    const double x = off[(size_t)(have - 1) * V + e % V];
    off[(size_t)have * V + e] = x;
    if (z32) z32[(size_t)have * V + e] = (float)x;
  }
}

void ga_host_cast(const double* x, float* z, int64_t n) {
  for (int64_t i = 0; i < n; ++i) z[i] = (float)x[i];
}

void ga_host_gather(const double* X_all, const float* F_all, int ld_in, const int32_t* idx, int n_out, int V,
                    int n_obj, double* X_out, float* F_out, int ld_out) {
  for (size_t e = 0; e < (size_t)n_out * V; ++e) {
    const int r = (int)(e / V), v = (int)(e % V);
    X_out[e] = X_all[(size_t)idx[r] * V + v];
    if (v < n_obj) F_out[(size_t)v * ld_out + r] = F_all[(size_t)v * ld_in + idx[r]];
  }
}

int64_t ga_host_rand_count(int M, int V) { return (int64_t)rand_count(M, V); }

}  // extern "C"
