// GPU-resident genetic operators: the per-generation GA work of the reference's search loop
// (run.py:59-76 -> pymoo GA / NSGA-II with the operators of operators.py:66-78) on device buffers, so that the
// population never leaves the GPU between generations: tournament -> SBX -> polynomial mutation -> duplicate
// elimination -> [fitness: glass_evaluate_device] -> rank + crowding survival -> gather.  The arithmetic lives in
// ga_ops.cuh (shared with the CPU test harness); this file is the kernels and the C ABI (glass_ga_*).
//
// These are microsecond kernels on kilobytes (the population is 256 KB at P = 64): the point is residency — no
// per-generation H2D of latents, no D2H of fitnesses, no host synchronisation inside a generation — not bandwidth.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/clipglass_b200.h"
#include "ga_ops.cuh"

using namespace glass_ga;

namespace {

thread_local std::string g_ga_error;

int ga_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_ga_error = buf;
  return code;
}

#define GA_LAUNCH_OK(what)                                                                          \
  do {                                                                                              \
    cudaError_t err__ = cudaGetLastError();                                                         \
    if (err__ != cudaSuccess) return ga_fail(GLASS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err__)); \
  } while (0)

constexpr int kThreads = 256;
constexpr int kMaxGroup = 4096;     // candidates one cooperating block handles (survival, permutations)

inline int blocks_for(size_t n) {
  size_t b = (n + kThreads - 1) / kThreads;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

OpParams to_params(const glass_ga_params* g) {
  OpParams p;
  p.sbx_eta = g->sbx_eta; p.sbx_prob = g->sbx_prob; p.sbx_prob_var = g->sbx_prob_var;
  p.pm_eta = g->pm_eta; p.pm_prob = g->pm_prob;
  p.n_var = g->n_var; p.integer = g->integer;
  return p;
}

__global__ void ga_uniform_kernel(double* out, size_t n, uint64_t seed, uint64_t offset) {
  const size_t pairs = (n + 1) / 2;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += (size_t)gridDim.x * blockDim.x) {
    double u0, u1;
    uniform_pair(seed, offset + q, u0, u1);
    out[2 * q] = u0;
    if (2 * q + 1 < n) out[2 * q + 1] = u1;
  }
}

// One block per permutation: out[p][stable_rank(key[p][j])] = j  ==  argsort(key[p], stable).
__global__ void ga_perm_kernel(const double* keys, int n, int32_t* out) {
  extern __shared__ double s_key[];
  const double* k = keys + (size_t)blockIdx.x * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) s_key[j] = k[j];
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) out[(size_t)blockIdx.x * n + stable_rank(s_key, n, j)] = j;
}

__global__ void ga_tournament_kernel(const int32_t* pairs, const int32_t* rank, const double* crowd, int n_select,
                                     int32_t* sel) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_select; t += gridDim.x * blockDim.x)
    sel[t] = tournament_element(pairs, rank, crowd, t);
}

__global__ void ga_offspring_kernel(OpParams p, const double* X, const int32_t* parents, const double* bounds,
                                    const double* rnd, int M, double* out) {
  const size_t n = (size_t)M * p.n_var;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
    offspring_element(p, X, parents, bounds, rnd, M, (int)(e / p.n_var), (int)(e % p.n_var), out);
}

// Block j decides whether candidate j duplicates a row of the population, an accepted offspring or an earlier
// candidate.  Rows almost always differ in their first variable, so a thread per row with an early exit.
__global__ void ga_dup_flags_kernel(const double* cand, const double* X, int n_x, const double* off,
                                    const int32_t* n_have, int V, double eps, int32_t* flags) {
  const int j = blockIdx.x;
  const int have = *n_have;
  const int rows = n_x + have + j;
  const double* c = cand + (size_t)j * V;
  int dup = 0;
  for (int r = threadIdx.x; r < rows && !dup; r += blockDim.x) {
    const double* row = r < n_x ? X + (size_t)r * V
                                : (r < n_x + have ? off + (size_t)(r - n_x) * V : cand + (size_t)(r - n_x - have) * V);
    dup = rows_equal(row, c, V, eps) ? 1 : 0;
  }
  dup = __syncthreads_or(dup);
  if (threadIdx.x == 0) flags[j] = dup;
}

// One block: destination row of every kept candidate (kept candidates fill off[n_have ...] in candidate order until
// n_off rows exist), then the new n_have.
__global__ void ga_dup_dest_kernel(const int32_t* flags, int n_c, int n_off, int32_t* n_have, int32_t* dest) {
  const int have = *n_have;
  __shared__ int s_kept;
  if (threadIdx.x == 0) s_kept = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < n_c; j += blockDim.x) {
    int d = -1;
    if (!flags[j]) {
      int before = 0;
      for (int i = 0; i < j; ++i) before += flags[i] ? 0 : 1;
      if (have + before < n_off) {
        d = have + before;
        atomicAdd(&s_kept, 1);
      }
    }
    dest[j] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) *n_have = have + s_kept;
}

__global__ void ga_append_kernel(const double* cand, const int32_t* dest, int n_c, int V, double* off, float* z32) {
  const size_t n = (size_t)n_c * V;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(e / V), v = (int)(e % V);
    const int d = dest[j];
    if (d < 0) continue;
    const double x = cand[e];
    off[(size_t)d * V + v] = x;
    if (z32) z32[(size_t)d * V + v] = (float)x;        // latent.py:38: the f64 population cast to f32
  }
}

// Rows [n_have, n_off) repeat the last accepted row (ga.Algorithm._mate pads with off[-1]).
__global__ void ga_pad_kernel(const int32_t* n_have, int n_off, int V, double* off, float* z32) {
  const int have = *n_have;
  if (have <= 0 || have >= n_off) return;
  const size_t n = (size_t)(n_off - have) * V;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(e % V);
    const double x = off[(size_t)(have - 1) * V + v];
    off[(size_t)have * V + e] = x;
    if (z32) z32[(size_t)have * V + e] = (float)x;
  }
}

__global__ void ga_cast_kernel(const double* x, float* z, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    z[i] = (float)x[i];
}

__global__ void __launch_bounds__(1024) ga_survive_kernel(SurviveState s) {
  __shared__ int s_fsize;
  s.fsize = &s_fsize;
  survive_body(s);
}

__global__ void ga_gather_kernel(const double* X_all, const float* F_all, int ld_in, const int32_t* idx, int n_out,
                                 int V, int n_obj, double* X_out, float* F_out, int ld_out) {
  const size_t n = (size_t)n_out * V;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / V), v = (int)(e % V);
    X_out[e] = X_all[(size_t)idx[r] * V + v];
    if (v < n_obj) F_out[(size_t)v * ld_out + r] = F_all[(size_t)v * ld_in + idx[r]];
  }
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace

extern "C" {

const char* glass_ga_last_error(void) { return g_ga_error.c_str(); }

int glass_ga_uniform(uint64_t seed, uint64_t offset, double* out_dev, int64_t n, void* stream) {
  if (!out_dev || n < 0) return ga_fail(GLASS_ERR_ARG, "glass_ga_uniform: bad argument");
  if (n == 0) return GLASS_OK;
  ga_uniform_kernel<<<blocks_for((size_t)(n + 1) / 2), kThreads, 0, (cudaStream_t)stream>>>(out_dev, (size_t)n, seed,
                                                                                             offset);
  GA_LAUNCH_OK("ga_uniform_kernel");
  return GLASS_OK;
}

int glass_ga_permutations(const double* keys_dev, int32_t n, int32_t n_perm, int32_t* out_dev, void* stream) {
  if (!keys_dev || !out_dev || n <= 0 || n_perm <= 0 || n > kMaxGroup)
    return ga_fail(GLASS_ERR_ARG, "glass_ga_permutations: need 0 < n <= %d, n_perm > 0", kMaxGroup);
  ga_perm_kernel<<<n_perm, kThreads, (size_t)n * sizeof(double), (cudaStream_t)stream>>>(keys_dev, n, out_dev);
  GA_LAUNCH_OK("ga_perm_kernel");
  return GLASS_OK;
}

int glass_ga_tournament(const int32_t* pairs_dev, const int32_t* rank_dev, const double* crowd_dev, int32_t n_select,
                        int32_t* selected_dev, void* stream) {
  if (!pairs_dev || !rank_dev || !crowd_dev || !selected_dev || n_select <= 0)
    return ga_fail(GLASS_ERR_ARG, "glass_ga_tournament: bad argument");
  ga_tournament_kernel<<<blocks_for((size_t)n_select), kThreads, 0, (cudaStream_t)stream>>>(pairs_dev, rank_dev,
                                                                                            crowd_dev, n_select,
                                                                                            selected_dev);
  GA_LAUNCH_OK("ga_tournament_kernel");
  return GLASS_OK;
}

int64_t glass_ga_rand_count(int32_t n_matings, int32_t n_var) {
  if (n_matings < 0 || n_var < 0) return GLASS_ERR_ARG;
  return (int64_t)rand_count(n_matings, n_var);
}

int glass_ga_offspring(const glass_ga_params* params, const double* x_dev, const int32_t* parents_dev,
                       const double* bounds_dev, const double* rand_dev, int32_t n_matings, double* out_dev,
                       void* stream) {
  if (!params || !x_dev || !parents_dev || !bounds_dev || !rand_dev || !out_dev || n_matings <= 0 ||
      params->n_var <= 0)
    return ga_fail(GLASS_ERR_ARG, "glass_ga_offspring: bad argument");
  const OpParams p = to_params(params);
  ga_offspring_kernel<<<blocks_for((size_t)n_matings * p.n_var), kThreads, 0, (cudaStream_t)stream>>>(
      p, x_dev, parents_dev, bounds_dev, rand_dev, n_matings, out_dev);
  GA_LAUNCH_OK("ga_offspring_kernel");
  return GLASS_OK;
}

int64_t glass_ga_dedup_workspace(int32_t n_cand) { return n_cand < 0 ? GLASS_ERR_ARG : (int64_t)n_cand * 8; }

int glass_ga_dedup_append(const double* cand_dev, int32_t n_cand, const double* x_dev, int32_t n_x, double* off_dev,
                          int32_t n_off, int32_t* n_have_dev, int32_t n_var, double eps, int32_t eliminate,
                          float* z32_dev, void* workspace_dev, void* stream) {
  if (!cand_dev || !off_dev || !n_have_dev || !workspace_dev || n_cand <= 0 || n_off <= 0 || n_var <= 0 || n_x < 0 ||
      (n_x > 0 && !x_dev))
    return ga_fail(GLASS_ERR_ARG, "glass_ga_dedup_append: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* flags = (int32_t*)workspace_dev;
  int32_t* dest = flags + n_cand;
  if (eliminate) {
    ga_dup_flags_kernel<<<n_cand, 128, 0, s>>>(cand_dev, x_dev, n_x, off_dev, n_have_dev, n_var, eps, flags);
    GA_LAUNCH_OK("ga_dup_flags_kernel");
  } else {
    cudaError_t err = cudaMemsetAsync(flags, 0, (size_t)n_cand * 4, s);
    if (err != cudaSuccess) return ga_fail(GLASS_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(err));
  }
  ga_dup_dest_kernel<<<1, 1024, 0, s>>>(flags, n_cand, n_off, n_have_dev, dest);
  GA_LAUNCH_OK("ga_dup_dest_kernel");
  ga_append_kernel<<<blocks_for((size_t)n_cand * n_var), kThreads, 0, s>>>(cand_dev, dest, n_cand, n_var, off_dev,
                                                                           z32_dev);
  GA_LAUNCH_OK("ga_append_kernel");
  return GLASS_OK;
}

int glass_ga_pad(double* off_dev, int32_t n_off, const int32_t* n_have_dev, int32_t n_var, float* z32_dev,
                 void* stream) {
  if (!off_dev || !n_have_dev || n_off <= 0 || n_var <= 0) return ga_fail(GLASS_ERR_ARG, "glass_ga_pad: bad argument");
  ga_pad_kernel<<<blocks_for((size_t)n_off * n_var), kThreads, 0, (cudaStream_t)stream>>>(n_have_dev, n_off, n_var,
                                                                                         off_dev, z32_dev);
  GA_LAUNCH_OK("ga_pad_kernel");
  return GLASS_OK;
}

int glass_ga_cast_f32(const double* x_dev, float* z_dev, int64_t n, void* stream) {
  if (!x_dev || !z_dev || n < 0) return ga_fail(GLASS_ERR_ARG, "glass_ga_cast_f32: bad argument");
  if (n == 0) return GLASS_OK;
  ga_cast_kernel<<<blocks_for((size_t)n), kThreads, 0, (cudaStream_t)stream>>>(x_dev, z_dev, (size_t)n);
  GA_LAUNCH_OK("ga_cast_kernel");
  return GLASS_OK;
}

int64_t glass_ga_survive_workspace(int32_t n) {
  if (n < 0) return GLASS_ERR_ARG;
  return (int64_t)(4 * align16((size_t)n * 4) + 2 * align16((size_t)n * 8));
}

int glass_ga_survive(const float* f_dev, int32_t ld, int32_t n, int32_t n_obj, int32_t n_survive, int32_t nsga2,
                     int32_t* idx_dev, int32_t* rank_dev, double* crowd_dev, void* workspace_dev, void* stream) {
  if (!f_dev || !idx_dev || !rank_dev || !crowd_dev || !workspace_dev || n <= 0 || n > kMaxGroup || n_obj <= 0 ||
      n_obj > 8 || n_survive <= 0 || n_survive > n || ld < n)
    return ga_fail(GLASS_ERR_ARG, "glass_ga_survive: need 0 < n_survive <= n <= %d, 0 < n_obj <= 8, ld >= n",
                   kMaxGroup);
  SurviveState s;
  s.F = f_dev; s.ld = ld; s.n = n; s.n_obj = n_obj; s.n_survive = n_survive; s.nsga2 = nsga2 ? 1 : 0;
  char* w = (char*)workspace_dev;
  const size_t ni = align16((size_t)n * 4), nd = align16((size_t)n * 8);
  s.ndom = (int*)w; w += ni;
  s.front = (int*)w; w += ni;
  s.cur = (int*)w; w += ni;
  s.pos = (int*)w; w += ni;
  s.cd = (double*)w; w += nd;
  s.sorted = (double*)w;
  s.fsize = nullptr;
  s.out_idx = idx_dev; s.out_rank = rank_dev; s.out_crowd = crowd_dev;
  ga_survive_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(s);
  GA_LAUNCH_OK("ga_survive_kernel");
  return GLASS_OK;
}

int glass_ga_gather(const double* x_all_dev, const float* f_all_dev, int32_t ld_in, const int32_t* idx_dev,
                    int32_t n_out, int32_t n_var, int32_t n_obj, double* x_out_dev, float* f_out_dev, int32_t ld_out,
                    void* stream) {
  if (!x_all_dev || !f_all_dev || !idx_dev || !x_out_dev || !f_out_dev || n_out <= 0 || n_var <= 0 || n_obj <= 0 ||
      n_obj > n_var)
    return ga_fail(GLASS_ERR_ARG, "glass_ga_gather: bad argument");
  ga_gather_kernel<<<blocks_for((size_t)n_out * n_var), kThreads, 0, (cudaStream_t)stream>>>(
      x_all_dev, f_all_dev, ld_in, idx_dev, n_out, n_var, n_obj, x_out_dev, f_out_dev, ld_out);
  GA_LAUNCH_OK("ga_gather_kernel");
  return GLASS_OK;
}

}  // extern "C"
